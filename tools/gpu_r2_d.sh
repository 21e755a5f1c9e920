#!/bin/bash
# round 2, GPU call D: tests + cfg2 / cfg1 bench after the commit-kernel changes (A/B: GRB_FIX_BS)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_e.log
tail -4 gpurun_out/pytest_e.log
for bs in 512; do
GRB_FIX_BS=$bs GRB_BENCH_SKIP_CPU=1 timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 > gpurun_out/bench_e_cfg2_bs$bs.json 2> gpurun_out/bench_e_cfg2_bs$bs.err; echo "cfg2 bs=$bs rc=$?"
done
GRB_BENCH_SKIP_CPU=1 timeout 600 python bench.py --workload cfg1 --steps 5 --warmup 3 > gpurun_out/bench_e_cfg1.json 2> gpurun_out/bench_e_cfg1.err; echo "cfg1 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_e_*.json")):
    try:
        d=json.load(open(f))
        print(f, "ms/step", round(d["ms_per_step"],1), "e2e_s", round(d["e2e"]["s_per_step"],3), "parity", d["parity_digest_ok"], "launches", d["gpu_launches"])
        print("  kernels", {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()}, d["commit_profile_last_step"])
    except Exception as e:
        print(f, "failed", e)
PY
