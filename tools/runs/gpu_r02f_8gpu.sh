#!/bin/bash
O=gpurun_out/r02f; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "rc=$?" >> $O/bench_8gpu.err
timeout 300 $TR bench.py --gpus 8 --steps 5 --warmup 3 --workload am-entity > $O/am_8gpu.json 2> $O/am_8gpu.err; echo "rc=$?" >> $O/am_8gpu.err
