#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests, bench, ncu launch list and full captures of the hot kernels.
# usage: tools/gpu_profile.sh <tag> [dirs-for-ncu-full]
TAG=${1:-r1}
ND=${2:-55}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.csv
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/test_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_$TAG.log
rm -f gpurun_out/*.npz
tail -3 gpurun_out/test_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_ref_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_$TAG.log 2>&1
wc -l gpurun_out/launches_$TAG.csv
# frame 2 of the run (kernel instances 0.. of frame 0 and 1 skipped): integrate, march, hit, resize x4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'raycast_hit_kernel|raycast_march_kernel|integrate_kernel|resize_map' -s 14 -c 7 \
    -f -o gpurun_out/prof_${TAG}_vol python bench.py --steps 1 --warmup 3 --no-cpu-baseline --dirs $ND > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full (volume kernels) rc=$?"; tail -1 gpurun_out/ncu_full_$TAG.log
# level-0 iteration of frame 2: assoc, deriv (two launches per iteration; frame 1 has 24, levels 2 and 1 of frame 2 have 14)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'icp_' -s 38 -c 2 \
    -f -o gpurun_out/prof_${TAG}_icp python bench.py --steps 1 --warmup 3 --no-cpu-baseline --dirs $ND > gpurun_out/ncu_full_${TAG}_icp.log 2>&1
echo "ncu full (icp) rc=$?"; tail -1 gpurun_out/ncu_full_${TAG}_icp.log; ls -la gpurun_out/ | tail -12
