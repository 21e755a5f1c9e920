#!/bin/bash
# round 2, GPU call J: cfg3 (1 Gbp genome, 30x) on one B200
mkdir -p gpurun_out
timeout 1500 python bench.py --workload cfg3 --steps 1 --warmup 1 > gpurun_out/bench_j_cfg3_n1.json 2> gpurun_out/bench_j_cfg3_n1.err; echo "cfg3 rc=$?"
tail -5 gpurun_out/bench_j_cfg3_n1.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_j_cfg3_n1.json"))
    print("ms/step", round(d["ms_per_step"],1), "value", round(d["value"],3), "e2e", d["e2e"], "parity", d["parity"])
    print("  kernels", {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()})
    print("  config", d["config"]); print("  roofline", d["roofline"]); print(" cpu", d["cpu_baseline"])
except Exception as e:
    print("failed", e)
PY
nvidia-smi --query-gpu=memory.used --format=csv
