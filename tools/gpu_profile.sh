#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests, bench, ncu launch list and one full capture of the hot kernels.
# usage: tools/gpu_profile.sh <tag> [dirs-for-ncu]
TAG=${1:-r1}
ND=${2:-21}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.csv
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/test_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_$TAG.log
tail -5 gpurun_out/test_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_$TAG.log 2>&1
wc -l gpurun_out/launches_$TAG.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'raycast_hit_kernel|raycast_march_kernel|integrate_kernel' -s 6 -c 3 \
    -f -o gpurun_out/prof_${TAG}_vol python bench.py --steps 1 --warmup 3 --no-cpu-baseline --dirs $ND > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full (volume kernels) rc=$?"; tail -3 gpurun_out/ncu_full_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'icp_' -s 60 -c 3 \
    -f -o gpurun_out/prof_${TAG}_icp python bench.py --steps 1 --warmup 3 --no-cpu-baseline --dirs $ND > gpurun_out/ncu_full_${TAG}_icp.log 2>&1
echo "ncu full (icp) rc=$?"; tail -3 gpurun_out/ncu_full_${TAG}_icp.log; ls -la gpurun_out/
