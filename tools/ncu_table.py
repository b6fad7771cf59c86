#!/usr/bin/env python
"""Markdown table of the tracked ncu metrics, one row per launch: python tools/ncu_table.py raw.csv [raw2.csv ...]"""
import sys
from ncu_summary import launches, _num

COLS = [("µs", 'gpu__time_duration.sum', 1.0, 1), ("DRAM rd MB", 'dram__bytes_read.sum', 1e-6, 1), ("DRAM wr MB", 'dram__bytes_write.sum', 1e-6, 1),
        ("DRAM %", 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 1.0, 1), ("warp-instr (M)", 'smsp__inst_executed.sum', 1e-6, 2),
        ("issue %", 'smsp__issue_active.avg.pct_of_peak_sustained_active', 1.0, 1), ("lanes/instr", 'smsp__thread_inst_executed_per_inst_executed.ratio', 1.0, 1),
        ("regs", 'launch__registers_per_thread', 1.0, 0), ("warps active %", 'sm__warps_active.avg.pct_of_peak_sustained_active', 1.0, 1),
        ("long_sb", 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 1.0, 2),
        ("wait", 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 1.0, 2),
        ("no_inst", 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 1.0, 2),
        ("I$ hit %", 'sm__icc_request_hit_rate.pct', 1.0, 1)]
print("| kernel | " + " | ".join(c[0] for c in COLS) + " |")
print("|---" * (len(COLS) + 1) + "|")
for f in sys.argv[1:]:
    for name, m in launches(f):
        cells = []
        for _, key, scale, nd in COLS:
            try:
                v = (_num(m, key) if key in ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum') else float(m[key][0])) * scale
                cells.append(f"{v:.{nd}f}")
            except (KeyError, ValueError):
                cells.append("-")
        print(f"| {name.split('(')[0].replace('void ', '')} | " + " | ".join(cells) + " |")
