#!/bin/bash
# round 2, GPU call C: what bounds a random probe (microbenchmark + counters), where the ordered
# commit's time goes (per-read records of its re-validation phase), ncu of one whole batch
mkdir -p gpurun_out
build/sector-roofline 0.0625 1 4 22 64 > gpurun_out/random_access_r02.jsonl 2> gpurun_out/random_access_r02.err; echo "roofline rc=$?"
M=gpu__time_duration.sum,dram__sectors_read.sum,dram__sectors_write.sum,dram__bytes_read.sum,lts__t_requests_srcunit_tex.sum,lts__t_sectors_srcunit_tex.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,lts__t_sector_hit_rate.pct,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__m_xbar2l1tex_read_sectors.sum,dram__cycles_active.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/ncu_r02_random_access_22gib.csv build/sector-roofline 22 > /dev/null 2> gpurun_out/ncu_ra.err; echo "ncu roofline rc=$?"
GRB_FIX_DEBUG=gpurun_out/fix_debug_cfg2.bin timeout 300 python tools/run_once.py cfg2 1 > gpurun_out/run_once_dbg.log 2>&1; echo "dbg rc=$?"; cat gpurun_out/run_once_dbg.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k3_|k2_query" -s 780 -c 13 -o gpurun_out/ncu_r02_batch python tools/run_once.py cfg2 1 > gpurun_out/ncu_batch.log 2>&1; echo "ncu batch rc=$?"; tail -3 gpurun_out/ncu_batch.log
ls -la gpurun_out | tail -12
