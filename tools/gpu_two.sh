#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
bash tools/gpu_scale.sh 2 quick 2>&1 | tee gpurun_out/scale2_quick.txt
B="python bench.py --no-cpu-baseline --no-ref-cuda --steps 10 --warmup 3"
for v in "XS_X=0" "XS_ICP_H_RED_HP=2" "XS_ICP_H_FULL=1"; do
  echo "== $v"
  env $v timeout 300 $B 2>> gpurun_out/exp_two.err | python -c "
import sys, json
r = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('fps %.1f' % r['value'], r['stages_ms_per_frame'], r['kernel_ms_per_frame'])
"
done 2>&1 | tee gpurun_out/exp_two.txt
