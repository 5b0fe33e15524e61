#!/bin/bash
O=gpurun_out/r02e; mkdir -p $O
python -m pytest tests/test_gpu_partitioned.py -q -k entity > $O/pytest_part.txt 2>&1; echo "rc=$?" >> $O/pytest_part.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 1200 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "rc=$?" >> $O/bench_2gpu.err
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > $O/ref_2gpu.json 2> $O/ref_2gpu.err; echo "rc=$?" >> $O/ref_2gpu.err
