#!/bin/bash
O=gpurun_out/r02h; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest1.txt 2>&1; echo "pytest rc=$?" >> $O/pytest1.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-streaming --no-cpu-baseline --no-partitioned > $O/ncu_bench.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --workload wn18-full --n-flows 3 --no-partitioned --no-streaming --no-cpu-baseline > $O/wn18_1gpu.json 2> $O/wn18_1gpu.err
for i in 1 2 3 4; do python -m pytest tests -m gpu -q 2>&1 | tail -1 >> $O/pytest_loop.txt; done
