"""`frender demux` end to end on a synthetic C4-shaped lane pair, host threads 1 vs all
(python tools/bench_demux_cli.py [pairs]).  Checks that both runs write byte-identical files."""
import gzip
import hashlib
import os
import sys
import tempfile
import time
import zlib
from argparse import Namespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frender_b200 import synth  # noqa: E402
from frender_b200.cli import frender_demux, frender_scan  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
spec = synth.make_spec("C4")


def gz_write(path, mate):
    z = zlib.compressobj(1, zlib.DEFLATED, 31)
    with open(path, "wb") as fh:
        for o in range(0, pairs, 200_000):
            fh.write(z.compress(synth.generate(spec, o, min(o + 200_000, pairs), read_no=mate)))
        fh.write(z.flush())


with tempfile.TemporaryDirectory() as d:
    os.chdir(d)
    r1, r2 = os.path.join(d, "L_R1_001.fastq.gz"), os.path.join(d, "L_R2_001.fastq.gz")
    gz_write(r1, 1)
    gz_write(r2, 2)
    sheet = os.path.join(d, "SampleSheet.csv")
    with open(sheet, "w") as fh:
        fh.write(spec.sheet_csv())
    t0 = time.perf_counter()
    frender_scan(Namespace(n=1, rc=True, c=1, s=None, o="res", p=None, b=sheet, files=[r1]))
    print(f"scan: {time.perf_counter() - t0:.2f} s")
    scan_csv = [f for f in os.listdir(d) if f.startswith("frender-scan-results_")][0]
    # the reference's demux only accepts the idx1,idx2,reads,... column order (F:306), not the order its
    # scan writes: reorder, as a user has to
    import csv
    rows = list(csv.reader(open(scan_csv, newline="")))
    order = [rows[0].index(c) for c in ("idx1", "idx2", "reads", "matched_idx1", "matched_idx2", "read_type",
                                        "sample_name", "demux_ok")]
    results = "results_for_demux.csv"
    with open(results, "w", newline="") as fh:
        csv.writer(fh).writerows([[r[i] for i in order] for r in rows])
    digests = {}
    for threads in ("1", str(len(os.sched_getaffinity(0)))):
        out = os.path.join(d, f"out{threads}")
        os.environ["FRENDER_DEMUX_THREADS"] = threads
        t0 = time.perf_counter()
        frender_demux(Namespace(r=os.path.join(d, results), d=out, o=None, no_index_hop=False,
                                no_ambiguous=False, no_undeter=False, no_samples=False, files=[r1, r2]))
        dt = time.perf_counter() - t0
        h = hashlib.sha256()
        for name in sorted(os.listdir(out)):
            h.update(name.encode())
            h.update(open(os.path.join(out, name), "rb").read())
        digests[threads] = h.hexdigest()
        print(f"demux, {threads} host thread(s): {dt:.2f} s = {pairs / dt:.3e} pairs/s, {len(os.listdir(out))} files")
    print("byte-identical outputs:", len(set(digests.values())) == 1)
