#!/bin/bash
O=gpurun_out/r03k; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"distmult_rs_kernel|distmult_dz_kernel" -s 2 -c 2 -o $O/decoder_stream python bench.py --workload wikikg2-part --steps 1 --warmup 1 > $O/ncu.log 2>&1; echo "ncu rc=$?"
ls -la $O
