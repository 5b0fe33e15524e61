#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel from `ncu -i rep --page source --csv` output.

usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME --launch-count 1 > src.csv
       python profiles/ncu_source_top.py src.csv [N]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
tot = sum(int(r[iS]) for r in data)
print(f"kernel: {rows[0][1] if rows[0] else '?'}\ntotal samples {tot}, {len(data)} SASS instructions")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
for r in data:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
print("stall reasons:", ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for idx, r in sorted(enumerate(data), key=lambda ir: -int(ir[1][iS]))[:n_top]:
    st = sorted(((int(r[i]), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"{idx:5d} {int(r[iS]):7d} {100 * int(r[iS]) / max(tot, 1):5.1f}% ex={r[iEx]:>9s} {r[iSrc].strip()[:64]:64s} {st}")
