#!/bin/bash
O=gpurun_out/r03g; mkdir -p $O; rm -f $O/loop.txt
timeout 600 python -m pytest tests/test_gpu_driver.py -q -x > $O/pytest_driver.txt 2>&1; echo "driver rc=$?"; tail -4 $O/pytest_driver.txt
timeout 600 python bench.py --no-streaming --no-cpu-baseline --no-partitioned > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
grep -h "one replay\|e2e losses" $O/bench.err | cut -c1-240
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r03g/bench.json").read().strip().splitlines()[-1])
w = d["e2e"]["with_sampler"]
print(d["ms_per_step"], d["e2e"]["ms_per_step"], w["device_sampler_ms"], w["device_sampler_captured_ms"])
PY
for i in 1 2 3 4; do python -m pytest tests -m gpu -q -x 2>&1 | tail -1 >> $O/loop.txt; done
for i in 1 2 3; do python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "^smoke" | cut -c1-200 >> $O/loop.txt; done
cat $O/loop.txt
