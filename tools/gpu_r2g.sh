#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in 0 7 1 2 4 6 0 7; do
  FRB_WAIT_MODE=$k timeout 300 python tools/prof_scan.py 40000000 5 24 2>&1 | tail -1 | sed "s/^/wait$k /" | tee -a gpurun_out/g_ab.log
done
FRB_WAIT_MODE=7 FRB_SCAN_TIMING=spec timeout 300 python tools/prof_scan.py 40000000 3 24 2>&1 | tail -2 | tee -a gpurun_out/g_ab.log
