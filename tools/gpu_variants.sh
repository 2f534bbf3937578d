#!/bin/bash
# Other workloads / layouts on one B200 for the tables in DESIGN.md and profiles/r02_ab_table.md (not the bench value).
mkdir -p gpurun_out
run() {
echo "== $*"
timeout 200 python bench.py --no-cpu-baseline --no-ref-cuda --steps 10 --warmup 3 "$@" 2> gpurun_out/variants.err | python -c "
import sys, json
r = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('fps %.1f' % r['value'], 'e2e %.1f' % r['e2e']['value'], r['config'].get('parameters'), r['config']['derivative_planes_rank0'], r['stages_ms_per_frame'], r['kernel_ms_per_frame'])
"
tail -2 gpurun_out/variants.err
}
run --pose-only
run --mode dcsfd
run --mode csfd --dirs 6 --res 256
run --mode csfd --dirs 6
