#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/i_pytest.log
tail -4 gpurun_out/i_pytest.log
for k in 1 2; do timeout 300 python tools/prof_scan.py 40000000 5 24 2>&1 | tail -1 | tee -a gpurun_out/i_ab.log; done
FRB_SCAN_TIMING=spec timeout 300 python tools/prof_scan.py 40000000 2 24 2>&1 | tail -6 | tee gpurun_out/i_probe.log
