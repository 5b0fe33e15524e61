#!/bin/bash
O=gpurun_out/r03c; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --no-streaming --no-cpu-baseline --no-partitioned > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --workload wn18-full --n-flows 3 --steps 5 --warmup 3 --no-streaming --no-cpu-baseline --no-partitioned > $O/wn18.json 2> $O/wn18.err; echo "wn18 rc=$?"
grep -h "kl_mog\|1x500x500\|e2e losses" $O/bench.err $O/wn18.err | cut -c1-200
python - <<'PY'
import json
for f in ("bench", "wn18"):
    try:
        d = json.loads(open(f"gpurun_out/r03c/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["step_mode"]["eager_ms_per_step"])
    except Exception as e:
        print(f, "failed", e)
PY
