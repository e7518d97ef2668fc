#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/j_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/j_pytest.log
tail -4 gpurun_out/j_pytest.log
for k in 0 1 0 1; do FRB_SPEC_SWAP=$k timeout 300 python tools/prof_scan.py 40000000 5 24 2>&1 | tail -1 | sed "s/^/swap$k /" | tee -a gpurun_out/j_ab.log; done
for k in 0 1; do FRB_SPEC_SWAP=$k FRB_SCAN_TIMING=spec timeout 300 python tools/prof_scan.py 40000000 2 24 2>&1 | tail -9 | cut -c1-220 | tee -a gpurun_out/j_probe.log; done
