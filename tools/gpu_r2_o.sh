#!/bin/bash
# round 2, GPU calls O: cfg4 (human scale, silver + golden) at N GPUs, host phase clocks on stderr.  usage: gpu_r2_o.sh N
N=${1:-8}; TAG=${2:-}
mkdir -p gpurun_out
{ nproc; free -g | head -2; } > gpurun_out/box$N.txt 2>&1
export GRB_BENCH_SKIP_CPU=1 GRB_TIMING=1 ${GRB_EXTRA_ENV}
timeout 1100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N bench.py --gpus $N --workload cfg4 --steps 1 --warmup 1 > gpurun_out/bench_o_cfg4_n$N$TAG.json 2> gpurun_out/bench_o_cfg4_n$N$TAG.err; echo "cfg4 N=$N rc=$?"
grep "grb timing" gpurun_out/bench_o_cfg4_n$N$TAG.err | tail -$((7*N)) | sort | uniq -c | sort -k4 | awk '{print $1, $4, $5, $6, $7}' | tail -30
tail -2 gpurun_out/bench_o_cfg4_n$N$TAG.err | cut -c1-300
cat gpurun_out/box$N.txt
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_o_cfg4_n$N$TAG.json"))
    print("ms/step", round(d["ms_per_step"],1), "value", round(d["value"],3), "parity", d["parity"])
    print("  kernels", {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()})
    print("  e2e", d["e2e"])
except Exception as e:
    print("failed", e)
PY
