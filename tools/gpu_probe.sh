#!/bin/bash
# sector roofline + fetch-granularity A/B of the bench
mkdir -p gpurun_out
timeout 600 build/sector-roofline 0.0625 1 4 21 64 > gpurun_out/sector_roofline.json 2> gpurun_out/sector_roofline.err
cat gpurun_out/sector_roofline.json
for g in 32 128; do
  GRB_L2_FETCH=$g timeout 600 python bench.py --steps 1 --warmup 1 > gpurun_out/bench_fetch$g.json 2> gpurun_out/bench_fetch$g.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_fetch$g.json"))
print($g, d["value"], d["kernels_ms_per_step"], d["e2e"]["value"])
PY
done
