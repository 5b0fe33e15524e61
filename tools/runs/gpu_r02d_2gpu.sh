#!/bin/bash
# round-2: 2-GPU validation of every multi-GPU path
O=gpurun_out/r02d; mkdir -p $O
python -m pytest tests/test_gpu_partitioned.py -q > $O/pytest_part.txt 2>&1; echo "rc=$?" >> $O/pytest_part.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "rc=$?" >> $O/bench_2gpu.err
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --workload am-entity > $O/am_2gpu.json 2> $O/am_2gpu.err; echo "rc=$?" >> $O/am_2gpu.err
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --workload wn18-full --n-flows 3 --no-partitioned > $O/wn18_2gpu.json 2> $O/wn18_2gpu.err; echo "rc=$?" >> $O/wn18_2gpu.err
timeout 600 python bench.py --steps 5 --warmup 3 --workload wn18-full --n-flows 3 --no-partitioned --no-streaming --no-cpu-baseline > $O/wn18_1gpu.json 2> $O/wn18_1gpu.err
timeout 600 python bench.py --steps 5 --warmup 3 --workload am-entity > $O/am_1gpu.json 2> $O/am_1gpu.err
