#!/bin/bash
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hessian.py tests/test_gpu_bench_config.py -m gpu -q > gpurun_out/test_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_$TAG.log
rm -f gpurun_out/*.npz
tail -8 gpurun_out/test_$TAG.log
B="python bench.py --no-cpu-baseline --no-ref-cuda --steps 10 --warmup 3"
for v in "XS_X=0" "XS_ICP_H_FULL=1"; do
  echo "== $v"
  env $v timeout 300 $B 2>> gpurun_out/exp_$TAG.err | python -c "
import sys, json
r = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('fps %.1f' % r['value'], r['stages_ms_per_frame'], r['kernel_ms_per_frame'])
"
done 2>&1 | tee gpurun_out/exp_$TAG.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'icp_deriv_h' -s 31 -c 1 \
    -f -o gpurun_out/prof_${TAG}_icp python bench.py --steps 1 --warmup 1 --frames-per-step 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_${TAG}_icp.log 2>&1
echo "ncu icp rc=$?"; tail -2 gpurun_out/ncu_${TAG}_icp.log
