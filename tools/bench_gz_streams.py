"""`-c N` scaling: N .fastq.gz files scanned at the same time on one GPU (python tools/bench_gz_streams.py [reads_per_file])."""
import os
import sys
import tempfile
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frender_b200 import synth  # noqa: E402
from frender_b200.cli import scan_files_concurrent  # noqa: E402
from frender_b200.engine import Context  # noqa: E402

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
spec = synth.make_spec("C2")
raw = synth.generate(spec, 0, reads) if reads <= 200_000 else b"".join(
    synth.generate(spec, o, min(o + 200_000, reads)) for o in range(0, reads, 200_000))
ctx = Context(0, table_log2=21)
with tempfile.TemporaryDirectory() as d:
    gz = os.path.join(d, "a_R1.fastq.gz")
    z = zlib.compressobj(1, zlib.DEFLATED, 31)
    with open(gz, "wb") as fh:
        fh.write(z.compress(raw))
        fh.write(z.flush())
    t0 = time.perf_counter()
    ctx.reset()
    ctx.scan_gz(gz, 0)
    one = time.perf_counter() - t0
    print(f"1 stream: {reads / one:.3e} reads/s ({len(raw) / one / 1e9:.2f} GB/s inflated)")
    for n in (2, 4, 8, 16):
        files = []
        for k in range(n):
            p = os.path.join(d, f"f{n}_{k}_R1.fastq.gz")
            os.link(gz, p)
            files.append(p)
        t0 = time.perf_counter()
        scan_files_concurrent(files, None, n, 0, 21, ctx)
        dt = time.perf_counter() - t0
        print(f"{n} streams: {n * reads / dt:.3e} reads/s ({n * len(raw) / dt / 1e9:.2f} GB/s inflated), {dt:.2f} s")
print("cores", len(os.sched_getaffinity(0)))
