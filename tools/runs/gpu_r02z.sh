#!/bin/bash
O=gpurun_out/r02z; mkdir -p $O
F="--workload wn18-full --n-flows 3 --steps 5 --warmup 3 --no-streaming --no-cpu-baseline --no-partitioned"
timeout 600 python bench.py $F > $O/wn18_captured.json 2> $O/wn18_captured.err; echo "captured rc=$?"
timeout 600 python bench.py $F --eager > $O/wn18_eager.json 2> $O/wn18_eager.err; echo "eager rc=$?"
grep -h "e2e losses" $O/*.err
python - <<'PY'
import json
for f in ("wn18_captured", "wn18_eager"):
    try:
        d = json.loads(open(f"gpurun_out/r02z/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["step_mode"]["eager_ms_per_step"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 $O/wn18_captured.err | cut -c1-300
