#!/bin/bash
# bulk-reduction depth sweep for the streaming message-passing launches (KG_EXP digits: units = 5x5 fwd, tens = 5x10 fwd, hundreds = 5x10 bwd)
O=gpurun_out/r02g; mkdir -p $O
for e in 0 11 22 132 243 302; do
  KG_EXP=$e python bench.py --streaming-only --steps 3 > $O/stream_$e.json 2> $O/stream_$e.err
  echo "== KG_EXP=$e" >> $O/summary.txt; grep "streaming/uniform" $O/stream_$e.err >> $O/summary.txt
done
python -m pytest tests -m gpu -q > $O/pytest1.txt 2>&1; echo "pytest rc=$?" >> $O/pytest1.txt
KG_EXP=111 python -m pytest tests/test_gpu_ops.py tests/test_gpu_golden.py -m gpu -q > $O/pytest_exp111.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_exp111.txt
