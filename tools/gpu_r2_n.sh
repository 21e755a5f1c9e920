#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "two_stage or sharded_run or full_size" > gpurun_out/pytest_n.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_n.log
export GRB_BENCH_SKIP_CPU=1 GRB_TIMING=1
taskset -c 0-7 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --workload cfg4s --steps 1 --warmup 1 > gpurun_out/bench_n_cfg4s_n2.json 2> gpurun_out/bench_n_cfg4s_n2.err; echo "cfg4s rc=$?"
grep "grb timing" gpurun_out/bench_n_cfg4s_n2.err | tail -14
python - <<'PY'
import json
for f in ("gpurun_out/bench_n_cfg4s_n2.json",):
    try:
        d=json.load(open(f))
        print(f, "ms/step", round(d["ms_per_step"],1), "value", round(d["value"],3), "parity", d["parity"])
        print("  e2e", d["e2e"])
    except Exception as e:
        print(f, "failed", e)
PY
