#!/bin/bash
# round 2, GPU call M (2 GPUs): whole GPU suite, cfg2 at N=1, the two-stage path on a tenth of cfg4 at N=2 with 4 cores per rank
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_m.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_m.log
export GRB_BENCH_SKIP_CPU=1
timeout 600 python bench.py --workload cfg2 --steps 3 --warmup 3 > gpurun_out/bench_m_cfg2_n1.json 2> gpurun_out/bench_m_cfg2_n1.err; echo "cfg2 rc=$?"
export GRB_TIMING=1
taskset -c 0-7 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --workload cfg4s --steps 1 --warmup 1 > gpurun_out/bench_m_cfg4s_n2.json 2> gpurun_out/bench_m_cfg4s_n2.err; echo "cfg4s rc=$?"
grep "grb timing" gpurun_out/bench_m_cfg4s_n2.err | tail -14
python - <<'PY'
import json
for f in ("gpurun_out/bench_m_cfg2_n1.json","gpurun_out/bench_m_cfg4s_n2.json"):
    try:
        d=json.load(open(f))
        print(f, "ms/step", round(d["ms_per_step"],1), "value", round(d["value"],3), "parity", d["parity_digest_ok"], d["parity"])
        print("  e2e", d["e2e"])
    except Exception as e:
        print(f, "failed", e)
PY
