#!/bin/bash
O=gpurun_out/r03e; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.txt
timeout 1200 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --workload fb15k237-step --no-streaming --no-cpu-baseline --no-partitioned > $O/step.json 2> $O/step.err; echo "step rc=$?"
timeout 600 python bench.py --workload am-entity --steps 5 --warmup 3 > $O/am.json 2> $O/am.err; echo "am rc=$?"
timeout 600 python bench.py --workload wn18-full --n-flows 3 --steps 5 --warmup 3 --no-streaming --no-cpu-baseline --no-partitioned > $O/wn18.json 2> $O/wn18.err; echo "wn18 rc=$?"
python - <<'PY'
import json
for f in ("bench", "step", "wn18", "am"):
    try:
        d = json.loads(open(f"gpurun_out/r03e/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"] if "ms_per_step" in d.get("e2e", {}) else d.get("e2e"), d.get("step_mode", {}).get("eager_ms_per_step"))
    except Exception as e:
        print(f, "failed", e)
PY
