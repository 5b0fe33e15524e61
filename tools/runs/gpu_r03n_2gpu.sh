#!/bin/bash
O=gpurun_out/r03n; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_partitioned.py -q > $O/pytest_part.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_part.txt | cut -c1-200
timeout 300 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_partitioned.py > $O/pytest_all.txt 2>&1; echo "pytest all rc=$?"; tail -1 $O/pytest_all.txt | cut -c1-200
