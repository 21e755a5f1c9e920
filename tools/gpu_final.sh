#!/bin/bash
# Round-end check on one GPU: smoke(), the GPU tests, the bench line (ours + reference arm), the ncu
# launch list and the full capture of the dominant kernel.  usage: tools/gpu_final.sh <tag>
set -u
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
tools/gpu_round.sh $TAG ncu
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_${TAG}_reference.json 2> $OUT/bench_${TAG}_reference.err
echo "reference arm exit $?"; cut -c1-400 $OUT/bench_${TAG}_reference.json
