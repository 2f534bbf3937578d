#!/bin/bash
# multi-GPU runs on an N-GPU box: hardware checks of the N-rank path, bench.py at 1..N ranks (as the driver launches it),
# configs[2] (300-frame sequence) and configs[4] (1024^3, pose-set Hessian) as written.  usage: tools/gpu_scale.sh N [quick]
N=${1:-2}
QUICK=${2:-}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_multigpu_nccl.py -m gpu -q > gpurun_out/test_nccl_n$N.log 2>&1; tail -3 gpurun_out/test_nccl_n$N.log
timeout 600 bash tools/check_driver_ranks.sh gpurun_out/driver_ranks 2>&1 | tail -3
if [ $N -ge 8 ]; then XS_CHECK_RES=256 timeout 600 $TR --nproc-per-node 8 --master-port 29534 tools/check_gather.py > gpurun_out/gather_check_n8.log 2>&1; tail -1 gpurun_out/gather_check_n8.log | cut -c1-300; fi
for n in 1 2 4 8; do
  [ $n -gt $N ] && break
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --no-cpu-baseline --no-ref-cuda > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    timeout 600 $TR --nproc-per-node $n --master-port 29511 bench.py --gpus $n > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  tail -1 gpurun_out/scale_n$n.json | python -c "
import sys, json
try:
    r = json.loads(sys.stdin.read())
    print('n_gpus', r['n_gpus'], 'fps %.1f' % r['value'], 'ms/frame %.3f' % r['ms_per_frame'], 'e2e %.1f' % r['e2e']['value'], r['stages_ms_per_frame'], r['config']['derivative_planes_rank0'])
except Exception as e:
    print('failed', e)
"
  tail -2 gpurun_out/scale_n$n.err
done
if [ -n "$QUICK" ]; then export XS_FRAMES=60; fi
if [ $N -eq 1 ]; then timeout 900 python tools/run_config3.py 2>&1 | tail -1 | cut -c1-600
else timeout 900 $TR --nproc-per-node $N --master-port 29512 tools/run_config3.py 2>&1 | grep '^{' | tail -1 | cut -c1-600; fi
if [ -n "$QUICK" ]; then export XS_RES=256 XS_DIRS=16 XS_FRAMES=8; else unset XS_FRAMES; fi
timeout 1200 $TR --nproc-per-node $N --master-port 29513 tools/run_config5.py > gpurun_out/config5_n$N.log 2>&1; grep '^{' gpurun_out/config5_n$N.log | tail -1 | cut -c1-700; tail -2 gpurun_out/config5_n$N.log | cut -c1-300
