#!/bin/bash
O=gpurun_out/r02w; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
tail -3 $O/pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?"
timeout 1200 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --workload wn18-full --n-flows 3 --steps 5 --warmup 3 --no-streaming --no-cpu-baseline --no-partitioned > $O/wn18.json 2> $O/wn18.err; echo "wn18 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --eager --no-streaming --no-cpu-baseline --no-partitioned > $O/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|split_kernel|amax_kernel" -c 24 -o $O/gemm_prepared python tools/ncu_gemm.py > $O/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
python - <<'PY'
import json
for f in ("bench", "wn18"):
    try:
        d = json.loads(open(f"gpurun_out/r02w/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["step_mode"]["eager_ms_per_step"])
    except Exception as e:
        print(f, "failed", e)
PY
ls -la $O
