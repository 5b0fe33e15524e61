#!/bin/bash
O=gpurun_out/r02i; mkdir -p $O
timeout 300 python tools/gemm_check.py > $O/gemm_check.txt 2>&1; echo "rc=$?" >> $O/gemm_check.txt
timeout 600 python -m pytest tests/test_gpu_determinism.py tests/test_gpu_ops.py -m gpu -q -x > $O/pytest_gemm.txt 2>&1; echo "rc=$?" >> $O/pytest_gemm.txt
