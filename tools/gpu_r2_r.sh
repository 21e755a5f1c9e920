#!/bin/bash
# round 2: cfg3 at N GPUs with the final code.  usage: gpu_r2_r.sh N
N=${1:-1}
mkdir -p gpurun_out
export GRB_BENCH_SKIP_CPU=1
if [ "$N" = 1 ]; then
  timeout 900 python bench.py --workload cfg3 --steps 1 --warmup 1 > gpurun_out/bench_r02f_cfg3_n1.json 2> gpurun_out/bench_r02f_cfg3_n1.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2958$N bench.py --gpus $N --workload cfg3 --steps 1 --warmup 1 > gpurun_out/bench_r02f_cfg3_n$N.json 2> gpurun_out/bench_r02f_cfg3_n$N.err
fi
echo "cfg3 N=$N rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r02f_cfg3_n$N.json"))
print("ms/step", round(d["ms_per_step"],1), "value", round(d["value"],3), "e2e_s", round(d["e2e"]["s_per_step"],2), d["e2e"]["phases_ms"], "parity", d["parity"]["out_digest"], d["parity"]["records"])
print("  kernels", {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()})
PY
