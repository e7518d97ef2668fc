#!/bin/bash
# what the driver runs at round end, in one call: build check is done in the container; here the -m gpu suite,
# smoke(), the default bench and the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/smoke.log
bash tools/gpu_bench.sh
