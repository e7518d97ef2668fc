#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in 1 2 4 8 16 1 4; do
  FRB_COPY_SPLIT=$k timeout 300 python tools/prof_scan.py 40000000 5 24 2>&1 | tail -1 | sed "s/^/split$k /" | tee -a gpurun_out/e_ab.log
done
