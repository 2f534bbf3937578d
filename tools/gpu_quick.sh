#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests + one bench line.  usage: tools/gpu_quick.sh <tag> [bench args]
TAG=${1:-q}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/test_$TAG.log 2>&1; rm -f gpurun_out/*.npz; echo "pytest rc=$?" >> gpurun_out/test_$TAG.log
tail -15 gpurun_out/test_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 3 "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
