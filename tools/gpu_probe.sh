#!/bin/bash
# sector roofline, fetch-granularity A/B of the bench, one full ncu capture of the dominant kernel
mkdir -p gpurun_out
timeout 600 build/sector-roofline 0.03125 0.125 0.5 4 21 64 > gpurun_out/sector_roofline.json 2> gpurun_out/sector_roofline.err
cat gpurun_out/sector_roofline.json
for g in 32 128; do
  GRB_L2_FETCH=$g timeout 600 python bench.py --steps 1 --warmup 1 > gpurun_out/bench_fetch$g.json 2> gpurun_out/bench_fetch$g.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_fetch$g.json"))
print($g, d["value"], d["kernels_ms_per_step"], d["e2e"]["value"])
PY
done
for k in k2_query k3_index k_fill_bits k3_bulk k3_fix; do
  skip=20; [ $k = k_fill_bits ] && skip=1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 \
    -o gpurun_out/ncu_$k -f python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_full_$k.log 2>&1
  echo "ncu $k exit $?"
done
