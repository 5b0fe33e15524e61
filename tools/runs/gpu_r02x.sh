#!/bin/bash
O=gpurun_out/r02x; mkdir -p $O; rm -f $O/diag.txt
for i in 1 2; do
  (FLOWS=0 KG_ROOT=$PWD timeout 200 python /root/repo/tools/diag_smoke.py 2>&1 | tail -5) >> $O/diag.txt
done
(cd .bisect/pre && FLOWS=0 KG_ROOT=$PWD timeout 200 python /root/repo/tools/diag_smoke.py 2>&1 | tail -5) >> $O/diag.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 >> $O/diag.txt
(cd .bisect/pre && python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1) >> $O/diag.txt
cat $O/diag.txt | cut -c1-330
