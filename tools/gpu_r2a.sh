#!/bin/bash
# round 2, GPU call A: parity tests, same-box A/B of the scan kernel against round 1, role timing, ncu capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_box.txt
python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/a_pytest.log
tail -3 gpurun_out/a_pytest.log
for k in r1 new r1 new; do
  if [ $k = r1 ]; then export FRB_SCAN_KERNEL=r1; else unset FRB_SCAN_KERNEL; fi
  python tools/prof_scan.py 40000000 5 24 2>&1 | tail -1 | sed "s/^/$k /" | tee -a gpurun_out/a_ab.log
done
unset FRB_SCAN_KERNEL
FRB_SCAN_LEAN=0 python tools/prof_scan.py 40000000 4 24 2>&1 | tail -1 | sed "s/^/general /" | tee -a gpurun_out/a_ab.log
FRB_SCAN_TIMING=1 python tools/prof_scan.py 20000000 3 24 2>&1 | tail -3 | tee -a gpurun_out/a_ab.log
FRB_SCAN_KERNEL=r1 FRB_SCAN_TIMING=1 python tools/prof_scan.py 20000000 3 24 2>&1 | tail -3 | sed "s/^/r1 /" | tee -a gpurun_out/a_ab.log
python tools/prof_scan.py 40000000 2 24 > gpurun_out/a_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_ws_kernel -s 1 -c 1 -o gpurun_out/scan_r2a -f python tools/prof_scan.py 40000000 2 24 > gpurun_out/a_ncu.log 2>&1
tail -2 gpurun_out/a_ncu.log
