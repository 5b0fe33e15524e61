#!/bin/bash
O=gpurun_out/r02j; mkdir -p $O
KG_GEMM_MC=0 timeout 300 python tools/gemm_check.py > $O/gemm_old.txt 2>&1
KG_GEMM_MC=1 timeout 300 python tools/gemm_check.py > $O/gemm_mc.txt 2>&1
KG_GEMM_MC=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-streaming --no-cpu-baseline --no-partitioned > $O/bench_old.json 2> $O/bench_old.err
KG_GEMM_MC=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-streaming --no-cpu-baseline --no-partitioned > $O/bench_mc.json 2> $O/bench_mc.err
KG_GEMM_MC=0 timeout 300 python bench.py --steps 5 --warmup 3 --workload wn18-full --n-flows 3 --no-partitioned --no-streaming --no-cpu-baseline > $O/wn18_old.json 2> $O/wn18_old.err
KG_GEMM_MC=1 timeout 300 python bench.py --steps 5 --warmup 3 --workload wn18-full --n-flows 3 --no-partitioned --no-streaming --no-cpu-baseline > $O/wn18_mc.json 2> $O/wn18_mc.err
timeout 300 python bench.py --steps 5 --warmup 3 --workload am-entity > $O/am_1gpu.json 2> $O/am_1gpu.err
