#!/bin/bash
# scaling runs on an N-GPU box: bench.py at 1..N ranks (N = $1), as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
for n in 1 2 4 8; do
  [ $n -gt $N ] && break
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  tail -1 gpurun_out/scale_n$n.json | python -c "
import sys, json
try:
    r = json.loads(sys.stdin.read())
    print('n_gpus', r['n_gpus'], 'fps %.1f' % r['value'], 'ms %.3f' % r['ms_per_step'], 'e2e %.1f' % r['e2e']['value'], r['stages_ms_per_step'])
except Exception as e:
    print('failed', e)
"
  tail -3 gpurun_out/scale_n$n.err
done
