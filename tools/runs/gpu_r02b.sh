#!/bin/bash
# round-2 second GPU session: new tests, restructured bench, L2 counters of the DistMult pass
O=gpurun_out/r02b; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest1.txt 2>&1; echo "pytest rc=$?" >> $O/pytest1.txt
python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
for i in 1 2 3 4 5 6 7 8; do python -m pytest tests -m gpu -q 2>&1 | tail -2 >> $O/pytest_loop.txt; done
timeout 600 python tools/poison_repro.py 4 ff > $O/poison_ff.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_read.sum,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__m_l1tex2xbar_write_bytes.sum --clock-control none -k regex:"distmult_rs|l2_probe" -c 12 --csv --log-file $O/ncu_distmult_l2.csv python bench.py --steps 2 --warmup 1 --no-streaming --no-cpu-baseline --no-partitioned > $O/ncu_bench.log 2>&1
