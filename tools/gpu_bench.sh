#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench rc $?"
tail -5 gpurun_out/bench_r2.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_r2.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","e2e_pcie","checked","cpu_baseline","gpu_launches"):
    print(k, j.get(k))
print("roofline", {k:v for k,v in j["roofline"].items() if k in ("achieved","frac","ms_per_launch","step_share")})
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; echo "ref rc $?"; cut -c1-600 gpurun_out/bench_r2_ref.json; tail -3 gpurun_out/bench_r2_ref.err
