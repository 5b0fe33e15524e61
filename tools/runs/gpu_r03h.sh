#!/bin/bash
O=gpurun_out/r03h; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_driver.py -q -x > $O/pytest_driver.txt 2>&1; echo "driver rc=$?"; tail -15 $O/pytest_driver.txt | cut -c1-220
