#!/bin/bash
# round 2, GPU call I: cfg5 in full (1-150 GB x h=1..5), random-access rates up to 150 GiB, what the
# MMU counters say about the > 64 GB cliff, the new long-read test
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hundreds_of_tiles or full_size" > gpurun_out/pytest_i.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_i.log
timeout 1200 python tools/probe_bench.py > gpurun_out/probe_bench_r02.jsonl 2> gpurun_out/probe_bench_r02.err; echo "probe rc=$?"
build/sector-roofline 32 96 128 150 > gpurun_out/random_access_r02_big.jsonl 2>> gpurun_out/random_access_r02.err; echo "roofline big rc=$?"
ncu --query-metrics 2>/dev/null | grep -i -E "tlb|mmu|pte|translat" | head -40 > gpurun_out/ncu_mmu_metrics.txt
echo "mmu metrics: $(wc -l < gpurun_out/ncu_mmu_metrics.txt)"
M=gpu__time_duration.sum,dram__sectors_read.sum,dram__bytes_read.sum,lts__t_requests_srcunit_tex.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,lts__t_sector_hit_rate.pct,dram__cycles_active.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
EXTRA=$(awk '{print $1}' gpurun_out/ncu_mmu_metrics.txt | grep -E "^[a-z_0-9.]+$" | head -12 | tr '\n' ',' | sed 's/,$//')
[ -n "$EXTRA" ] && M="$M,$EXTRA"
for gb in 32 96; do
timeout 600 ncu --metrics $M --clock-control none --csv -k regex:k_probe_query --log-file gpurun_out/ncu_r02_probe_query_${gb}gb.csv python tools/probe_bench.py --footprints $gb --h 3 > /dev/null 2> gpurun_out/ncu_probe_${gb}.err; echo "ncu probe $gb rc=$?"
done
