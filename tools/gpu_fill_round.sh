#!/bin/bash
# fill A/B sweep + per-kernel times of the partitioned fill under ncu (times only, no replay sets)
set -u
TAG=${1:-r01f}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -x -q -k "partitioned_fill or bitvector" > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?"; tail -3 $OUT/pytest_$TAG.log
timeout 600 python tools/fill_ab.py cfg2 > $OUT/fill_ab_$TAG.json 2> $OUT/fill_ab_$TAG.err
echo "fill_ab exit $?"; python - <<PY
import json
d=json.load(open("$OUT/fill_ab_$TAG.json"))
print("identical", d["identical"])
for k,v in d["variants"].items(): print(k, round(v["best_ms"],1), "ms", round(v["gprobes_per_s"],1), "Gprobe/s")
PY
for ps in 26 26_bs512; do
FILL_AB_ONLY=part_pshift$ps timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_fill -c 12 --csv \
  --log-file $OUT/ncu_fill_ps${ps}_$TAG.csv python tools/fill_ab.py cfg2 > /dev/null 2> $OUT/ncu_fill_$TAG.err
grep -v "^==" $OUT/ncu_fill_ps${ps}_$TAG.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    print('ps$ps', r['Kernel Name'][:14], r['Metric Value'], r['Metric Unit'])
" | head -8
done
