#!/bin/bash
O=gpurun_out/r02o; mkdir -p $O
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"distmult_rs|kl_mog_bwd|epilogue_only_vec4|topk_tc|topk_finish|fwd_kernel|l2_probe" -s 40 -c 14 -o $O/r02_kernels python tools/ncu_targets.py > $O/ncu.log 2>&1; echo "rc=$?" >> $O/ncu.log
timeout 600 ncu --set full --clock-control none -k regex:"basis_id_src_bwd|basis_small_bwd" -s 4 -c 2 -o $O/r02_am_kernels python bench.py --steps 1 --warmup 3 --workload am-entity > $O/ncu_am.log 2>&1; echo "rc=$?" >> $O/ncu_am.log
ls -la $O
