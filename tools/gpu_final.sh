#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests, the default bench line and the ncu launch list (no full captures).
# usage: tools/gpu_final.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/test_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_$TAG.log
rm -f gpurun_out/*.npz
tail -3 gpurun_out/test_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_$TAG.log 2>&1
wc -l gpurun_out/launches_$TAG.csv
