#!/bin/bash
O=gpurun_out/r03d; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "kl or triplet or distmult or golden or bench_configuration or loss or step" > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.txt
timeout 600 python bench.py --no-streaming --no-cpu-baseline --no-partitioned > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
grep -h "kl_mog\|triplet_index\|graph_index" $O/bench.err | cut -c1-200
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r03d/bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["step_mode"]["eager_ms_per_step"])
PY
