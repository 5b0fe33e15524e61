#!/bin/bash
# full GPU suite + default bench + WN18/IAF bench after the prepared-operand GEMM
mkdir -p gpurun_out/r02r
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02r/pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/r02r/pytest.txt
tail -5 gpurun_out/r02r/pytest.txt
timeout 900 python bench.py > gpurun_out/r02r/bench.json 2> gpurun_out/r02r/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --workload wn18-full --n-flows 3 --steps 5 --warmup 3 > gpurun_out/r02r/wn18.json 2> gpurun_out/r02r/wn18.err; echo "wn18 rc=$?"
grep -v "^\[cpu\|streaming\|wikikg2" gpurun_out/r02r/bench.err | tail -22
tail -16 gpurun_out/r02r/wn18.err
python - <<'PY'
import json
for f in ("bench", "wn18"):
    try:
        d = json.loads(open(f"gpurun_out/r02r/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["e2e"]["value"])
    except Exception as e:
        print(f, "failed", e)
PY
