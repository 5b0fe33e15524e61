#!/bin/bash
O=gpurun_out/r02o; mkdir -p $O
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"distmult_rs|kl_mog_bwd|epilogue_only_vec4|topk_tc|topk_finish|fwd_kernel|l2_probe" -s 6 -c 13 -o $O/r02_kernels python tools/ncu_targets.py > $O/ncu.log 2>&1; echo "rc=$?" >> $O/ncu.log
ls -la $O
