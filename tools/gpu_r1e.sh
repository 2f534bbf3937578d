#!/bin/bash
# validate HEAD: parity tests, bench at 55 / 7 / 1 directions (per-rank loads of the 1- and 8-GPU shards)
TAG=r1e
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/test_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_$TAG.log; rm -f gpurun_out/*.npz
tail -8 gpurun_out/test_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
for d in 7 1; do
timeout 300 python bench.py --steps 20 --warmup 3 --dirs $d --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_${TAG}_d$d.json 2>> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_${TAG}_d$d.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_${TAG}_d7.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --dirs 7 > gpurun_out/ncu_list_$TAG.log 2>&1
wc -l gpurun_out/launches_${TAG}_d7.csv
