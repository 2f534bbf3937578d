#!/bin/bash
# 2-GPU hardware check of the multi-GPU layer: NCCL test, C++ driver ranks, gathered record == 1-rank record, bench at 2 ranks
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_multigpu_nccl.py -m gpu -q > gpurun_out/test_nccl_n2.log 2>&1; tail -3 gpurun_out/test_nccl_n2.log
timeout 600 $TR --nproc-per-node 2 --master-port 29534 tools/check_gather.py > gpurun_out/gather_check_n2.log 2>&1; tail -1 gpurun_out/gather_check_n2.log | cut -c1-400
timeout 600 $TR --nproc-per-node 2 --master-port 29511 bench.py --gpus 2 > gpurun_out/scale_n2.json 2> gpurun_out/scale_n2.err
tail -1 gpurun_out/scale_n2.json | python -c "
import sys, json
r = json.loads(sys.stdin.read())
print('n_gpus', r['n_gpus'], 'fps %.1f' % r['value'], 'e2e %.1f' % r['e2e']['value'], r['stages_ms_per_frame'], r['config']['derivative_planes_rank0'])
"
tail -2 gpurun_out/scale_n2.err
