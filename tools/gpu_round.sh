#!/bin/bash
# One GPU-box round: parity tests, the bench line, the ncu launch list and one full capture.
# Usage (from the repo root, on the GPU box): tools/gpu_round.sh [tag] [steps...]
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
if [[ " $* " != *" notest "* ]]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1
  echo "pytest exit $?" >> $OUT/pytest_$TAG.log
  tail -5 $OUT/pytest_$TAG.log
fi
if [[ " $* " != *" nobench "* ]]; then
  GRB_TIMING=1 timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
  echo "bench exit $?"; tail -c 3000 $OUT/bench_$TAG.json; tail -12 $OUT/bench_$TAG.err
fi
if [[ " $* " == *" ncu "* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 1 --warmup 1 > $OUT/ncu_bench_$TAG.log 2>&1
  echo "ncu launches exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-k2_query} -s ${NCU_SKIP:-40} -c 1 \
    -o $OUT/ncu_query_$TAG -f python bench.py --steps 1 --warmup 1 > $OUT/ncu_full_$TAG.log 2>&1
  echo "ncu full exit $?"
fi
