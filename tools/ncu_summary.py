#!/usr/bin/env python
"""Condenses an .ncu-rep (ncu --set full) into the handful of metrics the roofline discussion uses.
usage: tools/ncu_summary.py report.ncu-rep > profiles/ncu_<round>_<kernel>.txt"""
import csv
import subprocess
import sys

WANT = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread
launch__shared_mem_per_block_dynamic launch__occupancy_limit_registers launch__occupancy_limit_shared_mem
dram__bytes_read.sum dram__bytes_write.sum dram__sectors_read.sum dram__sectors_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed lts__t_sector_hit_rate.pct
lts__t_sectors_srcunit_tex_op_read.sum lts__t_sectors_srcunit_tex_op_write.sum lts__t_sectors_op_atom.sum
lts__t_sectors_op_red.sum l1tex__m_xbar2l1tex_read_bytes.sum l1tex__m_l1tex2xbar_write_bytes.sum
l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum
lts__throughput.avg.pct_of_peak_sustained_elapsed l1tex__throughput.avg.pct_of_peak_sustained_elapsed
sm__warps_active.avg.pct_of_peak_sustained_active sm__throughput.avg.pct_of_peak_sustained_elapsed
smsp__issue_active.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio""".split()


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("# ncu --set full --clock-control none;", sys.argv[2] if len(sys.argv) > 2 else rep)
        print("Kernel Name\t" + vals[hdr.index("Kernel Name")])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"{w}\t{vals[i]}\t{units[i]}")


if __name__ == "__main__":
    main()
