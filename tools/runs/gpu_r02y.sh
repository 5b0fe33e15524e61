#!/bin/bash
O=gpurun_out/r02y; mkdir -p $O; rm -f $O/loop.txt
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  (FLOWS=0 KG_ROOT=$PWD timeout 100 python /root/repo/tools/diag_smoke.py 2>&1 | grep -E "CHECK|worst|kernel" | tr '\n' ' '; echo) >> $O/loop.txt
done
for i in 1 2 3 4 5 6; do
  (python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1) >> $O/loop.txt
done
grep -c EXCURSION $O/loop.txt; cut -c1-260 $O/loop.txt | tail -20
