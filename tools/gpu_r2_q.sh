#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_q.log
export GRB_BENCH_SKIP_CPU=1
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 > gpurun_out/bench_q_cfg2.json 2> gpurun_out/bench_q_cfg2.err; echo "cfg2 rc=$?"
timeout 600 python bench.py --workload cfg1 --steps 5 --warmup 3 > gpurun_out/bench_q_cfg1.json 2> gpurun_out/bench_q_cfg1.err; echo "cfg1 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_q_*.json")):
    try:
        d=json.load(open(f))
        print(f, "ms/step", round(d["ms_per_step"],1), "e2e_s", round(d["e2e"]["s_per_step"],3), "parity", d["parity_digest_ok"], "launches", d["gpu_launches"])
        print("  kernels", {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()}, "sum", round(sum(v for v in d["kernels_ms_per_step"].values()),1))
    except Exception as e:
        print(f, "failed", e)
PY
GRB_TIMING=1 timeout 300 python tools/run_once.py cfg2 3 > gpurun_out/run_once_q.log 2>&1; grep -E "grb timing|call" gpurun_out/run_once_q.log | tail -12
