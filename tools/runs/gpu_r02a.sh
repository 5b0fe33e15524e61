#!/bin/bash
# round-2 first GPU session: baseline tests, poison hunt, sanitizers, bench
mkdir -p gpurun_out/r02a
O=gpurun_out/r02a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
python -m pytest tests -m gpu -x -q > $O/pytest1.txt 2>&1; echo "pytest rc=$?" >> $O/pytest1.txt
timeout 900 python tools/poison_repro.py 6 ff > $O/poison_ff.txt 2>&1; echo "rc=$?" >> $O/poison_ff.txt
timeout 600 python tools/poison_repro.py 4 lo > $O/poison_lo.txt 2>&1; echo "rc=$?" >> $O/poison_lo.txt
T=tests/test_gpu_golden.py::test_fb15k_step_shape_against_oracle
timeout 900 compute-sanitizer --tool racecheck --print-limit 30 python -m pytest $T -x -q > $O/racecheck.txt 2>&1; echo "rc=$?" >> $O/racecheck.txt
timeout 900 compute-sanitizer --tool synccheck --print-limit 30 python -m pytest $T -x -q > $O/synccheck.txt 2>&1; echo "rc=$?" >> $O/synccheck.txt
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 900 compute-sanitizer --tool initcheck --print-limit 40 python -m pytest $T -x -q > $O/initcheck.txt 2>&1; echo "rc=$?" >> $O/initcheck.txt
timeout 900 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest $T -x -q > $O/memcheck.txt 2>&1; echo "rc=$?" >> $O/memcheck.txt
for i in 1 2 3 4 5 6; do python -m pytest tests -m gpu -q 2>&1 | tail -3 >> $O/pytest_loop.txt; done
timeout 900 python tools/stress_determinism.py 100 > $O/stress.txt 2>&1
python bench.py > $O/bench.json 2> $O/bench.err
