#!/bin/bash
# bench.py at 2 / 4 / 8 ranks of one 8-GPU box (as the driver launches it) + the 8-rank gather check.  usage: tools/gpu_scale_bench.sh
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
XS_CHECK_RES=256 timeout 300 $TR --nproc-per-node 8 --master-port 29534 tools/check_gather.py > gpurun_out/gather_check_n8.log 2>&1; grep '^{' gpurun_out/gather_check_n8.log | tail -1 | cut -c1-400
for n in 8 4 2; do
  timeout 300 $TR --nproc-per-node $n --master-port 29511 bench.py --gpus $n > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  tail -1 gpurun_out/scale_n$n.json | python -c "
import sys, json
try:
    r = json.loads(sys.stdin.read())
    print('n_gpus', r['n_gpus'], 'fps %.1f' % r['value'], 'ms/frame %.3f' % r['ms_per_frame'], 'e2e %.1f' % r['e2e']['value'], r['stages_ms_per_frame'], r['config']['derivative_planes_rank0'])
except Exception as e:
    print('failed', e)
"
  tail -2 gpurun_out/scale_n$n.err | cut -c1-300
done
