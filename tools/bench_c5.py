"""BASELINE configs[4] through the command line: `scan -n 0` of a 6 bp single-index run split over many files, the files
sharded over the visible GPUs (FRENDER_GPUS=N: file i on GPU i % N, tables merged over NCCL) against one GPU.

    python tools/bench_c5.py [files=64] [reads_per_file=100000]

Writes the files (zlib level 1) and the sample sheet to a temp directory, runs `frender.py scan -n 0 -c 4` on one GPU (runs of small files go through the device as one gzip stream)
and on all of them, checks that both write the same CSV bytes, and times the reference's tally_barcodes on the same
files beside them (the reference has no single-index MATCHER, F:104-107; its tally is what configs[4] pins)."""
import contextlib
import ctypes
import gzip
import io
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_scan(argv, cwd, gpus):
    from frender_b200 import cli
    os.environ["FRENDER_GPUS"] = str(gpus)
    os.environ["FRENDER_SINGLE_INDEX"] = "1"
    before = set(os.listdir(cwd))
    here = os.getcwd()
    os.chdir(cwd)
    try:
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            cli.main(argv)
        dt = time.perf_counter() - t0
    finally:
        os.chdir(here)
    made = [f for f in set(os.listdir(cwd)) - before if f.startswith("frender-scan-results_")]
    data = open(os.path.join(cwd, made[0]), "rb").read()
    os.remove(os.path.join(cwd, made[0]))
    return dt, data


def main():
    n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    per = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
    from frender_b200 import synth
    from frender_b200._lib import lib
    n = ctypes.c_int()
    lib.frb_device_count(ctypes.byref(n))
    spec = synth.make_spec("C5")
    with tempfile.TemporaryDirectory() as d:
        files, raw = [], 0
        t0 = time.perf_counter()
        for i in range(n_files):
            path = os.path.join(d, f"S{i:03d}_S{i}_L001_R1_001.fastq.gz")
            data = synth.generate_big(spec, i * per, (i + 1) * per)
            raw += len(data)
            with gzip.open(path, "wb", compresslevel=1) as fh:
                fh.write(data)
            files.append(path)
        gen_s = time.perf_counter() - t0
        sheet = os.path.join(d, "SampleSheet.csv")
        open(sheet, "w").write(spec.sheet_csv())
        argv = ["scan", "-n", "0", "-c", "4", "-o", "c5", "-b", sheet] + files
        for _ in range(3):                                             # warm-up: CUDA context, page cache, GPU clocks
            run_scan(argv, d, 1)
        one_s, one_csv = run_scan(argv, d, 1)
        out = {"files": n_files, "reads_per_file": per, "reads": n_files * per, "raw_bytes": raw,
               "gz_bytes": sum(os.path.getsize(f) for f in files), "generate_s": gen_s,
               "one_gpu": {"seconds": one_s, "reads_per_s": n_files * per / one_s, "how": "runs of small files inflated on the device as one multi-member stream (frb_scan_gz_batch)"}}
        if n.value > 1:
            many_s, many_csv = run_scan(argv, d, n.value)
            out["all_gpus"] = {"gpus": n.value, "seconds": many_s, "reads_per_s": n_files * per / many_s,
                               "csv_identical_to_one_gpu": many_csv == one_csv,
                               "note": "includes starting one process and CUDA context per GPU and the NCCL set-up"}
        ref_dir = os.path.join(ROOT, "baseline", "_ref")
        if os.path.exists(os.path.join(ref_dir, "frender.py")):
            import importlib
            import warnings
            warnings.simplefilter("ignore")
            sys.path.insert(0, ref_dir)
            ref = importlib.import_module("frender")
            cores = len(os.sched_getaffinity(0))
            sub = files[:min(n_files, 2 * cores)]
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                counter = ref.tally_barcodes(cores, [__import__("pathlib").Path(f) for f in sub], None)
            ref_s = time.perf_counter() - t0
            out["reference_tally"] = {"files": len(sub), "cores": cores, "seconds": ref_s,
                                      "reads_per_s": len(sub) * per / ref_s, "unique_keys": len(counter["total"]),
                                      "what": "tally_barcodes (F:183-207) only: the reference has no single-index matcher"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
