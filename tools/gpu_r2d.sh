#!/bin/bash
# round 2, GPU call D: committer-less speculative kernel (driver warp, per-lane deferred commits)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/d_pytest.log
tail -15 gpurun_out/d_pytest.log
for k in ring new regs48 new regs48; do
  unset FRB_SCAN_KERNEL FRB_WS_REGS
  [ $k = ring ] && export FRB_SCAN_KERNEL=ring
  [ $k = regs48 ] && export FRB_WS_REGS=48
  timeout 300 python tools/prof_scan.py 40000000 5 24 2>&1 | tail -1 | sed "s/^/$k /" | tee -a gpurun_out/d_ab.log
done
unset FRB_SCAN_KERNEL FRB_WS_REGS
timeout 300 python tools/prof_scan.py 40000000 2 24 > gpurun_out/d_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_spec_kernel -s 2 -c 1 -o gpurun_out/scan_r2d -f python tools/prof_scan.py 40000000 2 24 > gpurun_out/d_ncu.log 2>&1
tail -2 gpurun_out/d_ncu.log
timeout 600 python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; python -c "
import json;j=json.load(open('gpurun_out/d_bench.json'));print(j['value'],j['ms_per_step'],j['roofline']['frac'],j['roofline']['step_share'])"; tail -3 gpurun_out/d_bench.err
