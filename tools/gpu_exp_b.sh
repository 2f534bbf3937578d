#!/bin/bash
TAG=${1:-r02m}
mkdir -p gpurun_out
timeout 250 python -m pytest tests/test_gpu_hessian.py tests/test_gpu_second_order_oracle.py tests/test_gpu_intrinsics.py -m gpu -q -x > gpurun_out/test_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_$TAG.log
rm -f gpurun_out/*.npz
tail -3 gpurun_out/test_$TAG.log | cut -c1-600
run() {
echo "== $* $EXTRA"
env "$@" timeout 120 python bench.py --no-cpu-baseline --no-ref-cuda --steps 10 --warmup 3 $EXTRA 2> gpurun_out/exp_$TAG.err | python -c "
import sys, json
r = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('fps %.1f' % r['value'], r['config']['derivative_planes_rank0'], r['stages_ms_per_frame'], r['kernel_ms_per_frame'], r['roofline']['kernel_ms'])
" | tee -a gpurun_out/exp_$TAG.txt
tail -3 gpurun_out/exp_$TAG.err
}
EXTRA=""
run XS_X=0
run XS_ICP_NO_TAIL=1
EXTRA="--emulate-share 3/8"
run XS_X=0
