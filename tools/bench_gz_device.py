"""One .fastq.gz through frb_scan_gz: device-side inflate against the host zlib thread.
usage: bench_gz_device.py [reads] [gzip level]"""
import os
import sys
import time
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frender_b200 import synth  # noqa: E402
from frender_b200.engine import Context  # noqa: E402

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
spec = synth.make_spec("C2")
path = f"/tmp/bench_{reads}_{level}.fastq.gz"
if not os.path.exists(path):
    z = zlib.compressobj(level, zlib.DEFLATED, 31)
    with open(path, "wb") as fh:
        for g in range(0, reads, 250_000):
            fh.write(z.compress(synth.generate_big(spec, g, min(g + 250_000, reads))))
        fh.write(z.flush())
gz_bytes = os.path.getsize(path)
for mode in ("device", "host", "device"):
    os.environ["FRB_GZ_DEVICE"] = "1" if mode == "device" else "0"
    # the switch is read once per process: run each mode in a child
    import subprocess
    code = f"""
import sys, time; sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
from frender_b200.engine import Context
ctx = Context(0, table_log2=22)
best = None
for _ in range(3):
    ctx.reset(); t0 = time.perf_counter(); r, u, raw = ctx.scan_gz({path!r}, 0); dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
print("{mode}", r, u, raw, round(best, 4), "s", round(raw / best / 1e9, 2), "GB/s inflated", round(r / best / 1e6, 2), "M reads/s")
"""
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ))
print("gz bytes", gz_bytes)
