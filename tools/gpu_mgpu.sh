#!/bin/bash
# multi-GPU checks: N = number of visible GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(python -c "import ctypes;from frender_b200._lib import lib;n=ctypes.c_int();lib.frb_device_count(ctypes.byref(n));print(n.value)")
echo "GPUs: $N"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_check.py > gpurun_out/mgpu_check_${N}gpu.log 2>&1; echo "mgpu_check rc $?" >> gpurun_out/mgpu_check_${N}gpu.log; tail -3 gpurun_out/mgpu_check_${N}gpu.log
python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/mgpu_pytest_${N}gpu.log
for sc in weak strong; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 5 --warmup 3 --scaling $sc --no-cpu > gpurun_out/bench_${N}gpu_$sc.json 2> gpurun_out/bench_${N}gpu_$sc.err; echo "bench $sc rc $?"
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_${N}gpu_$sc.json').read().strip().splitlines()[-1])
    print("$sc", j["n_gpus"], j["scaling"], "value", j["value"], "ms", j["ms_per_step"], "e2e", j["e2e"] and j["e2e"]["value"], "pcie", j["e2e_pcie"] and (j["e2e_pcie"]["h2d_gbs"], j["e2e_pcie"]["pinned_memcpy_gbs"]), "checked", j["checked"], "uniq", j["unique_keys"], j["roofline"]["step_share"])
except Exception as e:
    print("parse failed", e)
PY
tail -3 gpurun_out/bench_${N}gpu_$sc.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 3 --warmup 3 --workload demux > gpurun_out/bench_${N}gpu_demux.json 2> gpurun_out/bench_${N}gpu_demux.err; echo "bench demux rc $?"
cut -c1-400 gpurun_out/bench_${N}gpu_demux.json; tail -3 gpurun_out/bench_${N}gpu_demux.err
