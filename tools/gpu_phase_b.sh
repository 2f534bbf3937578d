#!/bin/bash
# GPU check of a development step: parity tests, the default bench line (Hessian batch), the DCSFD-list layout for comparison.
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/test_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_$TAG.log
rm -f gpurun_out/*.npz
tail -25 gpurun_out/test_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --mode dcsfd --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_${TAG}_dcsfd.json 2>> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_${TAG}_dcsfd.json
