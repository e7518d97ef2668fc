#!/bin/bash
# round 2, GPU call C: early stage release -- parity tests, A/B, ncu, full-size bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c_pytest.log
tail -15 gpurun_out/c_pytest.log
for k in r1 new regs48 new regs48; do
  unset FRB_SCAN_KERNEL FRB_WS_REGS
  [ $k = r1 ] && export FRB_SCAN_KERNEL=r1
  [ $k = regs48 ] && export FRB_WS_REGS=48
  python tools/prof_scan.py 40000000 5 24 2>&1 | tail -1 | sed "s/^/$k /" | tee -a gpurun_out/c_ab.log
done
unset FRB_SCAN_KERNEL FRB_WS_REGS
python tools/prof_scan.py 40000000 2 24 > gpurun_out/c_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_ws_kernel -s 2 -c 1 -o gpurun_out/scan_r2c -f python tools/prof_scan.py 40000000 2 24 > gpurun_out/c_ncu.log 2>&1
tail -2 gpurun_out/c_ncu.log
python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; tail -c 1500 gpurun_out/c_bench.json; tail -3 gpurun_out/c_bench.err
