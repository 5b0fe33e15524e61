#!/bin/bash
O=gpurun_out/r02p; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_golden.py -m gpu -q -k "kl or golden or train_step or fb15k" > $O/pytest_kl.txt 2>&1; echo "rc=$?" >> $O/pytest_kl.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-streaming --no-cpu-baseline --no-partitioned > $O/bench.json 2> $O/bench.err
grep "kg_kl\|kg_distmult_bce" $O/bench.err
