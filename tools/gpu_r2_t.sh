#!/bin/bash
# round 2: ncu --set full of the (f4) warp kernel at the bench's size (1024 batches, 3072 jobs, 32 GB of filters)
mkdir -p gpurun_out
timeout 220 ncu --set full --clock-control none --import-source on -k regex:k_polish_fill_warp -s 1 -c 1 -f -o gpurun_out/ncu_r02_polish_warp python tools/polish_bench.py 1024 > gpurun_out/ncu_pw.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_pw.log | cut -c1-300
ls -la gpurun_out/ncu_r02_polish_warp.ncu-rep
