#!/bin/bash
O=gpurun_out/r03a; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_partitioned.py -q > $O/pytest_part.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_part.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 500 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench rc=$?"
grep -E "parity|wikikg2-part x2, |e2e losses" $O/bench_2gpu.err | cut -c1-250
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r03a/bench_2gpu.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["step_mode"]["eager_ms_per_step"])
p = d["partitioned"]; print({k: p[k] for k in p if k not in ("modes",)})
PY
