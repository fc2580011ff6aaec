#!/usr/bin/env python
"""Summarise an ncu --set full capture (raw page CSV) into the handful of numbers DESIGN.md / profiles/ quote.
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv ; python scripts/ncu_summary.py raw.csv [pattern]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
                 r"^(Kernel Name|gpu__time_duration.sum|dram__bytes_(read|write).sum$|gpu__dram_throughput.avg.pct|"
                 r"dram__throughput.avg.pct|l1tex__data_bank_conflicts_pipe_lsu_mem_shared|"
                 r"l1tex__data_pipe_lsu_wavefronts(_mem_shared)?(_op_ld)?.sum$|smsp__inst_executed.sum$|"
                 r"sm__inst_executed.avg.per_cycle_(active|elapsed)|smsp__issue_active.avg.pct|sm__warps_active.avg.pct|"
                 r"launch__registers_per_thread|launch__occupancy_limit|sm__cycles_elapsed.avg$|"
                 r"(sm|l1tex|lts)__throughput.avg.pct_of_peak_sustained_elapsed|smsp__inst_executed_op_shared|"
                 r"smsp__average_warps_issue_stalled_.*_per_issue_active|sm__pipe_(alu|fma|lsu).*pct|"
                 r"smsp__inst_executed_pipe_lsu|lts__t_sectors_srcunit_tex_op_read.sum$|lts__t_sector_hit_rate|"
                 r"l1tex__t_sector_hit_rate|smsp__cycles_active.avg$|sm__inst_executed_pipe_.*pct)")
for r in rows[2:]:
    for i, h in enumerate(hdr):
        if pat.search(h):
            print(f"{h} [{units[i]}] = {r[i]}")
    print("-" * 60)
