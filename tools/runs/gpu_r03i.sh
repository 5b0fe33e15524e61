#!/bin/bash
O=gpurun_out/r03i; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest.txt | cut -c1-200
timeout 900 python bench.py --workload wikikg2-part --steps 3 --warmup 3 > $O/wk1.json 2> $O/wk1.err; echo "wk rc=$?"
grep "wikikg2-part x1" $O/wk1.err | head -14
