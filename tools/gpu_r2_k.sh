#!/bin/bash
# round 2, GPU call K (8 GPUs): cfg2, cfg3 and cfg4 (silver + golden) at N = 8
mkdir -p gpurun_out
{ nproc; free -g | head -2; nvidia-smi --query-gpu=index,name,memory.total --format=csv; } > gpurun_out/box8.txt 2>&1
run() { # name workload steps warmup timeout port
  timeout $5 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $6 bench.py --gpus 8 --workload $2 --steps $3 --warmup $4 > gpurun_out/bench_k_$1.json 2> gpurun_out/bench_k_$1.err
  echo "$1 rc=$?"; tail -3 gpurun_out/bench_k_$1.err | cut -c1-300
}
export GRB_BENCH_SKIP_CPU=1
run cfg2_n8 cfg2 3 2 300 29521
run cfg3_n8 cfg3 1 1 600 29522
run cfg4_n8 cfg4 1 1 900 29523
cat gpurun_out/box8.txt | head -6
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_k_*.json")):
    try:
        d=json.load(open(f))
        print(f, "ms/step", round(d["ms_per_step"],1), "value", round(d["value"],3), "e2e", round(d["e2e"]["value"],3), "s", round(d["e2e"]["s_per_step"],2), "parity", d["parity_digest_ok"], d["parity"])
        print("  kernels", {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()})
        print("  e2e", d["e2e"])
    except Exception as e:
        print(f, "failed", e)
PY
