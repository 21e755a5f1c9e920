#!/bin/bash
# round 2, GPU call H (2 GPUs): multi-GPU parity tests (sharded ingest, slice mode, two-stage) + cfg2 at N=2
mkdir -p gpurun_out
nproc > gpurun_out/box2.txt; free -g | head -2 >> gpurun_out/box2.txt; nvidia-smi topo -m >> gpurun_out/box2.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_h.log
tail -15 gpurun_out/pytest_h.log
GRB_BENCH_SKIP_CPU=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload cfg2 --steps 3 --warmup 2 > gpurun_out/bench_h_cfg2_n2.json 2> gpurun_out/bench_h_cfg2_n2.err; echo "n2 rc=$?"
tail -5 gpurun_out/bench_h_cfg2_n2.err
GRB_SHARD_QUERY=1 GRB_BENCH_SKIP_CPU=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload cfg2 --steps 3 --warmup 2 > gpurun_out/bench_h_cfg2_n2_sq.json 2> gpurun_out/bench_h_cfg2_n2_sq.err; echo "n2 sq rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_h_*.json")):
    try:
        d=json.load(open(f))
        print(f, "ms/step", round(d["ms_per_step"],1), "e2e_s", round(d["e2e"]["s_per_step"],3), d["e2e"]["phases_ms"], "parity", d["parity_digest_ok"])
        print("  kernels", {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
cat gpurun_out/box2.txt | head -8
