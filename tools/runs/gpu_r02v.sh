#!/bin/bash
mkdir -p gpurun_out/r02v
timeout 300 python tools/distmult_probe.py > gpurun_out/r02v/probe.txt 2>&1; echo "rc=$?" >> gpurun_out/r02v/probe.txt
tail -12 gpurun_out/r02v/probe.txt
