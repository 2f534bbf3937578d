#!/bin/bash
# ncu --set full captures of the hot kernels of the default bench workload (frame 2 of the run).  usage: tools/gpu_ncu.sh TAG [kernels...]
TAG=${1:-r02}
shift
KERNELS=${@:-"icp raycast integrate"}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ref-cuda"
for k in $KERNELS; do
  case $k in
    icp) RE='icp_deriv_tile_kernel|icp_deriv_h_kernel'; SKIP=19;;   # frame 1 has 12 launches; level 0 of frame 2 starts at its 8th
    raycast) RE='raycast_hit_kernel'; SKIP=2;;
    integrate) RE='integrate_kernel'; SKIP=2;;
    assoc) RE='icp_assoc_kernel'; SKIP=19;;
    march) RE='raycast_march_kernel'; SKIP=2;;
  esac
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c 1 -f -o gpurun_out/prof_${TAG}_$k $B > gpurun_out/ncu_${TAG}_$k.log 2>&1
  echo "ncu $k rc=$?"; tail -2 gpurun_out/ncu_${TAG}_$k.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep | tail
