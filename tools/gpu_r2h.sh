#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
FRB_SCAN_TIMING=spec timeout 300 python tools/prof_scan.py 40000000 2 24 2>&1 | tail -10 | tee gpurun_out/h_probe.log
