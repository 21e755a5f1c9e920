#!/bin/bash
# bench under a list of environment settings (A/B runs), short form, no tests
# usage: tools/gpu_ab2.sh "VAR=val VAR2=val" "VAR=val" ...
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg GRB_BENCH_SKIP_CPU=1 timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_ab$i.json 2> gpurun_out/bench_ab$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_ab$i.json"))
    print("$cfg", "value", round(d["value"],4), "ms", round(d["ms_per_step"],1), {k:round(v,1) for k,v in d["kernels_ms_per_step"].items()}, "e2e", round(d["e2e"]["value"],4), d["config"]["reads_selected"], d["commit_profile_last_step"]["plans_changed"])
except Exception as e:
    print("$cfg", "FAILED", e); print(open("gpurun_out/bench_ab$i.err").read()[-1500:])
PY
done
