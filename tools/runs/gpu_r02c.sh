#!/bin/bash
# round-2 third GPU session: full suite with the new features, bench, launch list
O=gpurun_out/r02c; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest1.txt 2>&1; echo "pytest rc=$?" >> $O/pytest1.txt
python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-streaming --no-cpu-baseline --no-partitioned > $O/ncu_bench.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
