#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in 0 2 3; do echo "wait_mode $k"; FRB_WAIT_MODE=$k FRB_SCAN_TIMING=spec timeout 300 python tools/prof_scan.py 40000000 2 24 2>&1 | tail -9 | cut -c1-220 | tee -a gpurun_out/k_probe.log; done
