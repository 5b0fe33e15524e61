#!/bin/bash
mkdir -p gpurun_out/r02t
timeout 600 python -m pytest tests/test_gpu_driver.py -x -q -k captured > gpurun_out/r02t/pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/r02t/pytest.txt
tail -30 gpurun_out/r02t/pytest.txt
timeout 900 python bench.py --no-streaming --no-cpu-baseline --no-partitioned > gpurun_out/r02t/bench.json 2> gpurun_out/r02t/bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r02t/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02t/bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["step_mode"])
PY
