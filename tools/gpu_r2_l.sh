#!/bin/bash
# round 2, GPU call L (2 GPUs): multi-GPU tests again + the two-stage path on a tenth of cfg4 with 4 host cores per
# rank (what the 8-GPU box gives), host phase clocks on stderr
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_l.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_l.log
export GRB_BENCH_SKIP_CPU=1 GRB_TIMING=1
taskset -c 0-7 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --workload cfg4s --steps 1 --warmup 1 > gpurun_out/bench_l_cfg4s_n2.json 2> gpurun_out/bench_l_cfg4s_n2.err; echo "cfg4s rc=$?"
grep "grb timing" gpurun_out/bench_l_cfg4s_n2.err | tail -24
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_l_cfg4s_n2.json"))
    print("ms/step", round(d["ms_per_step"],1), "value", round(d["value"],3), "parity", d["parity"])
    print("  kernels", {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()})
    print("  e2e", d["e2e"])
except Exception as e:
    print("failed", e)
PY
