#!/bin/bash
O=gpurun_out/r02l; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "bdd" > $O/pytest_bdd.txt 2>&1; echo "rc=$?" >> $O/pytest_bdd.txt
timeout 900 python -m pytest tests -m gpu -q > $O/pytest1.txt 2>&1; echo "rc=$?" >> $O/pytest1.txt
