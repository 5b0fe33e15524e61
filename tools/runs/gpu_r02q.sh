#!/bin/bash
# prepared-operand GEMM bring-up: shapes in all four orientations, then the GEMM tests
mkdir -p gpurun_out/r02q
timeout 600 python tools/gemm_check.py > gpurun_out/r02q/gemm_check.txt 2>&1; echo "rc=$?" >> gpurun_out/r02q/gemm_check.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "gemm or linear or made or golden or fb15k" > gpurun_out/r02q/pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/r02q/pytest.txt
tail -30 gpurun_out/r02q/gemm_check.txt; tail -5 gpurun_out/r02q/pytest.txt
