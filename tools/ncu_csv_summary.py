#!/usr/bin/env python
"""Roofline-relevant raw metrics out of `ncu -i X.ncu-rep --page raw --csv` dumps (one block per profiled launch)."""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_not_selected",
        "smsp__pcsamp_warps_issue_stalled_membar", "smsp__pcsamp_warps_issue_stalled_short_scoreboard"]
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    rows = [r for r in rows if len(r) > 5]
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"== {path}: {r[hdr.index('Kernel Name')][:160]}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"   {w} = {r[i]} {units[i]}")
