#!/bin/bash
# experiment runner on the GPU box: parity tests, then bench lines for a list of "ENV=.. ENV=.. -- bench args" variants
TAG=${1:-exp}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/test_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_$TAG.log; rm -f gpurun_out/*.npz
tail -6 gpurun_out/test_$TAG.log
i=0
for v in "$@"; do
  i=$((i+1))
  echo "=== variant $i: $v"
  ( eval "env $v" ) > gpurun_out/bench_${TAG}_$i.json 2> gpurun_out/bench_${TAG}_$i.err
  python - <<PY
import json
try:
    r=json.loads(open("gpurun_out/bench_${TAG}_$i.json").read().strip().splitlines()[-1])
    print("fps %.1f ms %.3f e2e %.1f  stages %s  roofline %.3f (%.4f ms)  integrate %.4f ms" % (r["value"], r["ms_per_step"], r["e2e"]["value"], {k: round(v,3) for k,v in r["stages_ms_per_step"].items()}, r["roofline"]["frac"], r["roofline"]["kernel_ms"], r["roofline_integrate"]["kernel_ms"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_${TAG}_$i.err").read()[-1500:])
PY
done
