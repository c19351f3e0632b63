#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of counters the walk kernel is judged on.
usage: python tools/ncu_summary.py file.ncu-rep [title] > profiles/xxx.txt"""
import csv
import subprocess
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__sectors_read.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_read.sum.per_second',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__warps_eligible.avg.per_cycle_active']


def main():
    rep = sys.argv[1]
    if len(sys.argv) > 2:
        print("#", sys.argv[2])
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        for i, h in enumerate(hdr):
            if h in WANT or ('issue_stalled' in h and 'per_issue_active' in h) or \
               ('inst_executed_pipe_' in h and h.endswith('.avg.pct_of_peak_sustained_active')):
                try:
                    if ('issue_stalled' in h and float(vals[i]) < 0.15) or ('_pipe_' in h and float(vals[i]) < 1):
                        continue
                except ValueError:
                    pass
                print(f"{h:96s} {vals[i]} {units[i]}")
        print()


main()
