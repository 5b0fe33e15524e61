#!/bin/bash
O=gpurun_out/r03f; mkdir -p $O; rm -f $O/loop.txt
for i in 1 2 3 4 5; do python -m pytest tests -m gpu -q -x 2>&1 | tail -1 >> $O/loop.txt; done
for i in 1 2 3 4; do python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "^smoke" | cut -c1-200 >> $O/loop.txt; done
cat $O/loop.txt
