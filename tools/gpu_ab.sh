#!/bin/bash
# parity tests, then the bench under a list of environment settings (A/B runs)
# usage: tools/gpu_ab.sh "VAR=val VAR2=val" "VAR=val" ...
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_ab.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_ab.log
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_ab$i.json 2> gpurun_out/bench_ab$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_ab$i.json"))
    print("$cfg", "value", round(d["value"],4), "ms", round(d["ms_per_step"],1), {k:round(v,1) for k,v in d["kernels_ms_per_step"].items()}, "e2e", round(d["e2e"]["value"],4), d["config"]["reads_selected"])
except Exception as e:
    print("$cfg", "FAILED", e); print(open("gpurun_out/bench_ab$i.err").read()[-1500:])
PY
done
