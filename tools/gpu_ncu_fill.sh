#!/bin/bash
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
FILL_AB_ONLY=${2:-part_pshift26_bs512} timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fill_part -s 2 -c 1 \
  -o $OUT/ncu_fill_full_$TAG -f python tools/fill_ab.py cfg2 > $OUT/ncu_fill_full_$TAG.log 2>&1
echo "ncu exit $?"; tail -3 $OUT/ncu_fill_full_$TAG.log
ls -la $OUT/*.ncu-rep
