#!/bin/bash
TAG=${1:-r02u}
mkdir -p gpurun_out
XS_ICP_H_TILE=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${TAG}_share8.csv \
    python bench.py --steps 1 --warmup 1 --frames-per-step 4 --no-cpu-baseline --no-ref-cuda --emulate-share 3/8 > gpurun_out/ncu_list_$TAG.log 2>&1
wc -l gpurun_out/launches_${TAG}_share8.csv
