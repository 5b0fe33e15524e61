#!/usr/bin/env python
"""Pick the roofline-relevant counters out of `ncu -i rep --page raw --csv` output.

usage: ncu -i X.ncu-rep --page raw --csv > raw.csv ; python profiles/ncu_raw_pick.py raw.csv
"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
        "smsp__inst_executed.sum"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
name_i = hdr.index("Kernel Name")
for r in rows[2:]:
    print(f"== {r[name_i][:110]}")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"   {k:66s} {r[i]:>16s} {units[i]}")
