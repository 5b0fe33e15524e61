#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.

usage: python profiles/ncu_launch_summary.py launches.csv [N]
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.
"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{sum(v[0] for v in agg.values())} launches, {tot / 1e3:.3f} ms total")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:n_top]:
    print(f"{v[1] / 1e3:9.3f} ms {100 * v[1] / tot:5.1f}% x{v[0]:5d}  {k[:110]}")
