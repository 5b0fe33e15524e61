#!/bin/bash
O=gpurun_out/r03l; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.txt | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.txt | cut -c1-220
timeout 1200 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
grep "wikikg2-part x1" $O/bench.err | head -16
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r03l/bench.json").read().strip().splitlines()[-1])
w = d["e2e"]["with_sampler"]
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], w["device_sampler_ms"], w["device_sampler_captured_ms"])
r = d["roofline"]
print({k: (r[k]["kernel"], round(r[k]["frac"], 3)) for k in ("hbm", "l2_reduction", "tensor") if k in r and r[k]}, round(r["frac"], 3))
print(r["partitioned"]["ms_per_step"])
PY
