#!/bin/bash
TAG=${1:-r02u}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'icp_deriv_tile_kernel' -s 12 -c 1 -f -o gpurun_out/prof_${TAG}_icp_l2 \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_${TAG}_icp_l2.log 2>&1
echo rc=$?; tail -2 gpurun_out/ncu_${TAG}_icp_l2.log
