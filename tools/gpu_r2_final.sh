#!/bin/bash
# round 2, final 1-GPU call: smoke(), GPU tests, the bench line as the driver runs it (ours + reference arm),
# the ncu launch list of a short bench run and the full capture of the dominant kernel
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_final.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_final_reference.json 2> gpurun_out/bench_r02_final_reference.err; echo "reference rc=$?"
GRB_BENCH_SKIP_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r02_cfg2.csv python bench.py --steps 1 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_query -s 200 -c 2 -o gpurun_out/ncu_r02_k2_query python tools/run_once.py cfg2 1 > gpurun_out/ncu_q.log 2>&1; echo "ncu k2_query rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r02_final.json"))
print("ms/step", round(d["ms_per_step"],1), "value", round(d["value"],3), "e2e", d["e2e"]["value"], d["e2e"]["s_per_step"], "parity", d["parity_digest_ok"], "launches", d["gpu_launches"])
print("roofline", d["roofline"]); print("cpu", d["cpu_baseline"]); print("same sample", d["e2e_on_reference_sample"]); print("clocks", d["clocks"])
r=json.load(open("gpurun_out/bench_r02_final_reference.json")); print("reference", r["value"], r["ms_per_step"])
PY
