#!/bin/bash
# GPU check of a development step: parity tests, default bench line of both arms.
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.csv
nproc > gpurun_out/nproc_$TAG.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/test_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_$TAG.log
rm -f gpurun_out/*.npz
tail -15 gpurun_out/test_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_ref_$TAG.json
