#!/bin/bash
# One multi-GPU box round (gpurun --gpus N): parity tests (single- and multi-GPU), the fill A/B,
# the bench line at 1 GPU and at N GPUs.  usage: tools/gpu_multi_round.sh <tag> <N>
set -u
TAG=${1:-r01m}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_$TAG.log
tail -6 $OUT/pytest_$TAG.log
timeout 600 python tools/fill_ab.py cfg2 > $OUT/fill_ab_$TAG.json 2> $OUT/fill_ab_$TAG.err
echo "fill_ab exit $?"; cat $OUT/fill_ab_$TAG.json; tail -3 $OUT/fill_ab_$TAG.err
timeout 900 python bench.py > $OUT/bench_${TAG}_n1.json 2> $OUT/bench_${TAG}_n1.err
echo "bench n1 exit $?"; tail -c 2500 $OUT/bench_${TAG}_n1.json; tail -5 $OUT/bench_${TAG}_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
  --master-port 29541 bench.py --gpus $N > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err
echo "bench n$N exit $?"; tail -c 2500 $OUT/bench_${TAG}_n$N.json; tail -8 $OUT/bench_${TAG}_n$N.err
