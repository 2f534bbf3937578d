#!/bin/bash
# Round capture on one B200: parity tests, the default bench line of both arms, the ncu launch list of the bench command and
# --set full captures of the three hot kernels.  usage: tools/gpu_round.sh TAG
TAG=${1:-r02}
mkdir -p gpurun_out
bash tools/gpu_phase_a.sh $TAG
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --frames-per-step 4 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_list_$TAG.log 2>&1
wc -l gpurun_out/launches_$TAG.csv
bash tools/gpu_ncu.sh $TAG icp raycast integrate
