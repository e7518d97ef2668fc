#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/prof_scan.py 40000000 2 24 > gpurun_out/l_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_spec_kernel -s 2 -c 1 -o gpurun_out/scan_r2l -f python tools/prof_scan.py 40000000 2 24 > gpurun_out/l_ncu.log 2>&1
tail -2 gpurun_out/l_ncu.log
timeout 600 python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; python -c "
import json;j=json.load(open('gpurun_out/l_bench.json'));print(j['value'],j['ms_per_step'],j['roofline']['frac'],j['roofline']['step_share'])"; tail -3 gpurun_out/l_bench.err
