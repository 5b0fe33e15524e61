#!/bin/bash
mkdir -p gpurun_out/r02u
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-partitioned > gpurun_out/r02u/bench2.json 2> gpurun_out/r02u/bench2.err; echo "bench2 rc=$?"
tail -5 gpurun_out/r02u/bench2.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02u/bench2.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["step_mode"])
PY
