#!/bin/bash
mkdir -p gpurun_out/r02s
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02s/pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/r02s/pytest.txt
tail -8 gpurun_out/r02s/pytest.txt
