#!/bin/bash
O=gpurun_out/r03o; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -1 $O/pytest.txt | cut -c1-200
timeout 600 python bench.py --no-streaming --no-cpu-baseline --no-partitioned > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
grep -h "graph_index\|triplet_index" $O/bench.err | cut -c1-120
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r03o/bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["step_mode"]["eager_ms_per_step"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-streaming --no-cpu-baseline --no-partitioned > $O/ncu_bench.log 2>&1; echo "ncu list rc=$?"
