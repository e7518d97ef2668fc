#!/bin/bash
# ncu capture of the inflate kernels (after the same command ran plain)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
G="python tools/bench_gz_device.py 4000000 6"
$G > gpurun_out/ncu_plain_gz.log 2>&1 || { echo "plain gz failed"; exit 1; }
tail -3 gpurun_out/ncu_plain_gz.log
ncu --target-processes all --set full --clock-control none --import-source on -k regex:"gz_decode_kernel|gz_find_kernel|gz_crc_blocks_kernel|gz_resolve_kernel" -c 8 -f -o gpurun_out/gz $G > gpurun_out/ncu_gz.log 2>&1; echo "gz capture rc $?"
ncu -i gpurun_out/gz.ncu-rep --page raw --csv > gpurun_out/r02_gz_kernels_ncu_full_4Mreads.csv 2>/dev/null
ncu -i gpurun_out/gz.ncu-rep --page details --csv 2>/dev/null | grep -i "gz_find_kernel" | grep -i "stall\|Warp Cycles\|Issue\|Occupancy\|Waves\|Registers\|Duration\|Local" | head -60 > gpurun_out/gz_find_details.txt
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out/r02_gz_kernels_ncu_full_4Mreads.csv
