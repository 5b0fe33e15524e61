#!/bin/bash
O=gpurun_out/r03p; mkdir -p $O
timeout 120 python -m pytest tests/test_gpu_ops.py -q -x -k "graph" > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -1 $O/pytest.txt | cut -c1-200
timeout 200 python bench.py --workload wikikg2-part --steps 2 --warmup 3 > $O/wk1.json 2> $O/wk1.err; echo "wk rc=$?"
grep "wikikg2-part x1" $O/wk1.err | grep -E "graph_index|single|triplet_index" 
python -c "
import json; d=json.loads(open('gpurun_out/r03p/wk1.json').read().strip().splitlines()[-1]); print(d.get('ms_per_step'), {k:v for k,v in d.get('all_ops_ms',{}).items() if 'index' in k})"
