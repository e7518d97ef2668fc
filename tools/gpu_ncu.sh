#!/bin/bash
# ncu evidence of the round: launch list of the bench command, full captures of the matcher, router and inflate kernels.
# Every command runs plain first (must exit 0), then under ncu; nothing printed under ncu is a bench number.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --reads 20000000 --steps 2 --warmup 3 --no-cpu --no-e2e --no-demux --no-check"
$B > gpurun_out/ncu_plain_bench.json 2> gpurun_out/ncu_plain_bench.err || { echo "plain bench failed"; exit 1; }
cut -c1-300 gpurun_out/ncu_plain_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_20Mreads.csv $B > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc $?"
ncu --set full --clock-control none --import-source on -k regex:match_cand_kernel -c 2 -f -o gpurun_out/match_cand $B > gpurun_out/ncu_match.log 2>&1; echo "match capture rc $?"
ncu -i gpurun_out/match_cand.ncu-rep --page raw --csv > gpurun_out/r02_match_cand_kernel_ncu_full_20Mreads.csv 2>/dev/null
D="python tools/bench_demux.py 2000000 64 1"
$D > gpurun_out/ncu_plain_demux.json 2>&1 || { echo "plain demux failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"route|scan_ws|scan_redo" -c 150 --csv --log-file gpurun_out/r02_route_launches_2Mpairs.csv $D > gpurun_out/ncu_route_list.log 2>&1; echo "route list rc $?"
ncu --set full --clock-control none --import-source on -k regex:"route_copy_kernel|route_hist_kernel|route_scan_kernel" -s 12 -c 4 -f -o gpurun_out/route $D > gpurun_out/ncu_route.log 2>&1; echo "route capture rc $?"
ncu -i gpurun_out/route.ncu-rep --page raw --csv > gpurun_out/r02_route_kernels_ncu_full_2Mpairs.csv 2>/dev/null
G="python tools/bench_gz_device.py 2000000 6"
$G > gpurun_out/ncu_plain_gz.log 2>&1 || { echo "plain gz failed"; exit 1; }
tail -3 gpurun_out/ncu_plain_gz.log
ncu --target-processes all --set full --clock-control none --import-source on -k regex:"gz_decode_kernel|gz_find_kernel" -c 2 -f -o gpurun_out/gz $G > gpurun_out/ncu_gz.log 2>&1; echo "gz capture rc $?"
ncu -i gpurun_out/gz.ncu-rep --page raw --csv > gpurun_out/r02_gz_kernels_ncu_full_2Mreads.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -20
