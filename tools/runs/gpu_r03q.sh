#!/bin/bash
O=gpurun_out/r03q; mkdir -p $O
timeout 200 python -m pytest tests -m gpu -q -x > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -1 $O/pytest.txt | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-220
