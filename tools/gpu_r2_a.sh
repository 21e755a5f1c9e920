#!/bin/bash
# round 2, GPU call A: box facts, the GPU test suite, cfg1 / cfg2 bench lines with the digest check
mkdir -p gpurun_out
{ nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total --format=csv; df -h /dev/shm /tmp | tail -2; } > gpurun_out/box.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_a.log
tail -5 gpurun_out/pytest_a.log
GRB_BENCH_SKIP_CPU=1 timeout 600 python bench.py --workload cfg1 --steps 3 --warmup 3 > gpurun_out/bench_a_cfg1.json 2> gpurun_out/bench_a_cfg1.err; echo "cfg1 rc=$?"
GRB_BENCH_SKIP_CPU=1 timeout 900 python bench.py --workload cfg2 --steps 5 --warmup 3 > gpurun_out/bench_a_cfg2.json 2> gpurun_out/bench_a_cfg2.err; echo "cfg2 rc=$?"
cat gpurun_out/box.txt
python - <<'PY'
import json
for w in ("cfg1","cfg2"):
    try:
        d=json.load(open(f"gpurun_out/bench_a_{w}.json"))
        print(w, "ms/step", round(d["ms_per_step"],1), "value", round(d["value"],3), "e2e", round(d["e2e"]["value"],3), "parity", d["parity_digest_ok"], d["parity"])
        print("  kernels", {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()})
    except Exception as e:
        print(w, "failed", e)
PY
