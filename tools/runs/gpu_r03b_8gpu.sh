#!/bin/bash
O=gpurun_out/r03b; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 420 $TR bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "bench rc=$?"
grep -E "parity|wikikg2-part x8, |e2e losses" $O/bench_8gpu.err | cut -c1-250
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r03b/bench_8gpu.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["step_mode"]["eager_ms_per_step"], d["e2e"]["with_sampler"]["device_sampler_ms"])
p = d["partitioned"]; print({k: p[k] for k in p if k not in ("modes", "what")}); print(p["modes"]["allgather"]["comm_ms"], p["modes"]["allgather"]["top_ops_ms"])
print(d["roofline"]["eval"])
PY
