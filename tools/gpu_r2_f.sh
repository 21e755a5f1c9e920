#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_f.log
tail -3 gpurun_out/pytest_f.log
for agg in 1 0; do
GRB_FIX_AGG=$agg GRB_BENCH_SKIP_CPU=1 timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 > gpurun_out/bench_f_cfg2_agg$agg.json 2> gpurun_out/bench_f_cfg2_agg$agg.err; echo "cfg2 agg=$agg rc=$?"
done
GRB_BENCH_SKIP_CPU=1 timeout 600 python bench.py --workload cfg1 --steps 5 --warmup 3 > gpurun_out/bench_f_cfg1.json 2> gpurun_out/bench_f_cfg1.err; echo "cfg1 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_f_*.json")):
    try:
        d=json.load(open(f))
        print(f, "ms/step", round(d["ms_per_step"],1), "e2e_s", round(d["e2e"]["s_per_step"],3), "parity", d["parity_digest_ok"], "launches", d["gpu_launches"])
        print("  kernels", {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()}, d["commit_profile_last_step"])
    except Exception as e:
        print(f, "failed", e)
PY
