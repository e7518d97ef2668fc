#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gz.py -m gpu -x -q > gpurun_out/gz_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/gz_pytest.log
tail -30 gpurun_out/gz_pytest.log
FRB_GZ_TIMING=1 timeout 900 python tools/bench_gz_device.py 4000000 6 2>&1 | tail -30 | tee gpurun_out/gz_bench.log
