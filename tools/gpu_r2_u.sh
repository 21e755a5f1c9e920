#!/bin/bash
# round 2: (f4) waves pipelined (a wave's Bloom filters copied back under the next wave's kernel)
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_polish.py tests/test_polish_inputs.py -q -m gpu 2>&1 | tail -3
timeout 100 python tools/polish_bench.py 1024 > gpurun_out/pb_d.json 2> gpurun_out/pb_d.err; echo "rc=$?"; cut -c1-600 gpurun_out/pb_d.json
timeout 100 python tools/polish_bench.py 4096 > gpurun_out/pb_e.json 2> gpurun_out/pb_e.err; echo "rc=$?"; cut -c1-600 gpurun_out/pb_e.json
