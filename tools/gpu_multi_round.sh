#!/bin/bash
# One multi-GPU box round (gpurun --gpus N): parity tests (single- and multi-GPU), the bench line
# at 1 GPU and at N GPUs (pass-2 query replicated = default, and sharded with GRB_SHARD_QUERY=1).
# usage: tools/gpu_multi_round.sh <tag> <N> [notest] [non1]
set -u
TAG=${1:-r01m}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus_$TAG.txt 2>&1
if [[ " $* " != *" notest "* ]]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1
  echo "pytest exit $?" >> $OUT/pytest_$TAG.log
  tail -6 $OUT/pytest_$TAG.log
fi
if [[ " $* " != *" non1 "* ]]; then
  timeout 900 python bench.py > $OUT/bench_${TAG}_n1.json 2> $OUT/bench_${TAG}_n1.err
  echo "bench n1 exit $?"; tail -5 $OUT/bench_${TAG}_n1.err
fi
for SQ in 0 1; do
  GRB_SHARD_QUERY=$SQ timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 2954$SQ bench.py --gpus $N > $OUT/bench_${TAG}_n${N}_sq$SQ.json 2> $OUT/bench_${TAG}_n${N}_sq$SQ.err
  echo "bench n$N shard_query=$SQ exit $?"; tail -4 $OUT/bench_${TAG}_n${N}_sq$SQ.err | cut -c1-300
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_${TAG}_n*.json")):
    try:
        d=json.load(open(f))
        print(f.split("/")[-1], "value", round(d["value"],4), "ms", round(d["ms_per_step"],1), {k:round(v,1) for k,v in d["kernels_ms_per_step"].items()}, "e2e", round(d["e2e"]["value"],4), d["e2e"]["phases_ms"])
    except Exception as e:
        print(f, "FAILED", e)
PY
