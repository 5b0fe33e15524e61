#!/bin/bash
O=gpurun_out/r03j; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q -x -k "distmult or two_pass" > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.txt | cut -c1-200
timeout 900 python bench.py --workload wikikg2-part --steps 3 --warmup 3 > $O/wk1.json 2> $O/wk1.err; echo "wk rc=$?"
grep "wikikg2-part x1" $O/wk1.err | head -8
