#!/bin/bash
O=gpurun_out/r02n; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 8 --steps 4 --warmup 3 --workload wikikg2-part --col-chunks 1 > $O/part_c1.json 2> $O/part_c1.err; echo "rc=$?" >> $O/part_c1.err
timeout 600 $TR bench.py --gpus 8 --steps 4 --warmup 3 --workload wikikg2-part --col-chunks 2 > $O/part_c2.json 2> $O/part_c2.err; echo "rc=$?" >> $O/part_c2.err
