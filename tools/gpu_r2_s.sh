#!/bin/bash
# round 2, closing check of the committed tree: smoke(), GPU tests, the default bench line
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_t.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_t.log
timeout 150 python bench.py > gpurun_out/bench_r02_close.json 2> gpurun_out/bench_r02_close.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r02_close.json"))
print("ms/step", round(d["ms_per_step"],1), "value", round(d["value"],3), "e2e", d["e2e"]["value"], "parity", d["parity_digest_ok"], "launches", d["gpu_launches"], "frac", d["roofline"]["frac"])
PY
