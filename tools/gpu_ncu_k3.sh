#!/bin/bash
# ncu --set full of the batch-index / commit kernels of one mid-run batch
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
GRB_BENCH_SKIP_CPU=1 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'k3_(mark|members|frames|conf|bulk|fix)' -s 120 -c 6 \
  -o $OUT/ncu_k3_$TAG -f python bench.py --steps 1 --warmup 0 > $OUT/ncu_k3_$TAG.log 2>&1
echo "ncu exit $?"; tail -2 $OUT/ncu_k3_$TAG.log
