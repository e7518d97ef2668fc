"""Generate tests/golden/*.json by RUNNING the reference (/root/reference/frender.py).

Run in the build container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden.py
The fixtures pin the oracle (tests/test_oracle_golden.py) and, through it and
directly, the CUDA path (tests/test_gpu_*.py).  Inputs are stored next to the
reference's outputs so nothing has to be regenerated at test time.
"""
import argparse
import base64
import contextlib
import gzip
import io
import json
import os
import shutil
import sys
import tempfile
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.simplefilter("ignore")
import frender as F  # noqa: E402  (the reference itself)

from frender_b200 import synth  # noqa: E402


def b64(b):
    return base64.b64encode(b).decode()


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def matcher_kats():
    a1 = ["AAAAAAAA", "AAAAAAAT", "CCCCCCCC", "acgtacgt", "NNNNCCCC"]
    a2 = ["GGGGGGGG", "GGGGGGGG", "TTTTTTTT", "TTTTAAAA", "GGGGGGGG"]
    ids = ["s0", "s1", "s2", "s3", "s4"]
    queries = [
        (0, "AAAAAAAA", "GGGGGGGG"), (1, "AAAAAAAA", "GGGGGGGG"), (1, "AAAAAAAT", "GGGGGGGG"),
        (0, "AAAAAAAA", "TTTTTTTT"), (0, "ACGTACGT", "TTTTAAAA"), (0, "NNNNCCCC", "GGGGGGGG"),
        (0, "NAAAAAAA", "GGGGGGGG"), (1, "NAAAAAAA", "GGGGGGGG"), (0, "AAAAAAAA", "GGGGGGGC"),
        (2, "CCCCCCCC", "TTTTTTAA"), (8, "GATTACAG", "GATTACAG"), (3, "CCCCAAAA", "TTTTGGGG"),
        (1, "NNNNNNNN", "NNNNNNNN"), (2, "ACGTACGA", "TTTTAAAT"),
    ]
    out = {"idx1": a1, "idx2": a2, "id": ids, "classify": [], "approx": [], "rc": []}
    for n, i1, i2 in queries:
        out["classify"].append({"n": n, "idx1": i1, "idx2": i2,
                                "want": F.analyze_barcode(i1, i2, a1, a2, ids, n)})
        out["approx"].append({"n": n, "q": i1, "want": F.get_indexes_of_approx_matches(i1, a1, n)})
    out["revcomp"] = [[s, F.reverse_complement(s)] for s in ["ACGTNacgtnX", "", "A", "GGGGCCCC", "NNAC"]]
    tables = [
        (["AAAAAAAA", "CCCCCCCC"], ["ACGTTTTT", "AAAAACGT"], ["x", "y"]),
        (["AAAAAAAA", "AAAAAAAA"], ["ACGTTTTT", "AAAAACGT"], ["x", "y"]),
        (["AAAAAAAA", "AAAAAAAA"], ["ACGTTTTT", "AAAAACGT"], ["x", "x"]),
        (["AAAAAAAA", "CCCCCCCC"], ["ACGTACGT", "TTTTAAAA"], ["pal", "y"]),
    ]
    for t1, t2, tid in tables:
        for n in (0, 1):
            for key in ["AAAAAAAA+ACGTTTTT", "CCCCCCCC+ACGTTTTT", "AAAAAAAA+ACGTACGT",
                        "CCCCCCCC+TTTTAAAA", "GGGGGGGG+ACGTTTTT", "AAAAAAAA+ACGTTTTT+extra"]:
                for rc in (False, True):
                    out["rc"].append({"idx1": t1, "idx2": t2, "id": tid, "n": n, "key": key, "reads": 7,
                                      "rc_mode": rc,
                                      "want": F.analyze_barcodes_with_rc(key, 7, t1, t2, tid, n, rc)})
    return out


HEADER_CASES = [
    "@EAS139:136:FC706VJ:2:2104:15343:197393 1:Y:18:AAAAAAAA+GGGGGGGG\n",
    "@r 1:N:0:AAAA+CCCC extra:stuff\n",
    "@r  1:N:0:GG+TT\n",
    "@r 1:N:0:ACGT\n",
    "@r AACC+GGTT\n",
    "@r 1:N:0:\n",
    "@a:b:c 2:N:0:NNNN+ACGT",
    "@x 1:N:0:ACGTN+TTTTT\n\n",
    "@x:y 1:N:0:AC+GT:AA+TT\n",
]


def header_kats():
    out = []
    for line in HEADER_CASES:
        scan = line.rstrip("\n").split(" ")[1].split(":")[-1]          # F:169 verbatim expression
        demux = line.split(":")[-1].rstrip("\n")                       # F:778 verbatim expression
        out.append({"line": line, "scan": scan, "demux": demux})
    return out


def write_gz(path, data, members=1):
    step = (len(data) + members - 1) // members if members > 1 else len(data)
    with open(path, "wb") as fh:
        if not data:
            fh.write(gzip.compress(b"", 1, mtime=0))
        for off in range(0, len(data), max(step, 1)):
            fh.write(gzip.compress(data[off:off + step], 9, mtime=0))


def scan_case(name, config, reads, n, rc, tmp, n_samples=None, files=None, sample=None, prefix=""):
    """One full scan through the reference's own functions, stage by stage."""
    spec = synth.make_spec(config, n_samples=n_samples)
    case_dir = os.path.join(tmp, name)
    os.makedirs(case_dir)
    sheet = os.path.join(case_dir, "SampleSheet.csv")
    with open(sheet, "w") as fh:
        fh.write(spec.sheet_csv())
    indexes = F.get_indexes(sheet)
    inputs, paths = {}, []
    files = files or [("Undetermined_S0_L001_R1_001.fastq.gz", 0, reads)]
    for fname, g0, g1 in files:
        data = synth.generate_big(spec, g0, g1)
        p = os.path.join(case_dir, fname)
        write_gz(p, data)
        shutil.copy(p, os.path.join(HERE, f"{name}__{fname}"))
        inputs[fname] = [g0, g1]
        paths.append(p)
    counter = quiet(F.tally_barcodes, 1, paths, sample)
    first = quiet(F.process, 1, counter["total"], indexes, n, rc)
    out = {"config": config, "n_samples": n_samples, "n": n, "rc": rc, "sample": sample,
           "prefix": prefix, "files": inputs, "sheet_csv": spec.sheet_csv(), "indexes": indexes,
           "tally": {k: list(v.items()) for k, v in counter.items()},
           "first_pass": [[k, dict(v)] for k, v in first.items()]}
    results = first
    if rc:
        calls = F.call_rc_mode_per_id(F.flatten_results(first), indexes["id"])
        out["rc_calls"] = [[k, v] for k, v in calls.items()]
        cwd = os.getcwd()
        os.chdir(case_dir)
        quiet(F.report_rc_call_info, calls, indexes, "frender-scan-results_x.csv")
        out["rc_calls_csv"] = b64(open("frender-index-2-calls_x.csv", "rb").read())
        os.chdir(cwd)
        oriented = dict(indexes)
        oriented["idx2"] = [F.reverse_complement(s) if calls[i]["call"] else s
                            for s, i in zip(indexes["idx2"], indexes["id"])]
        out["oriented_idx2"] = oriented["idx2"]
        results = quiet(F.process, 1, counter["total"], oriented, n, False)
    results, bad = F.call_barcodes_correctly_distributed(counter, results, prefix)
    out["final"] = [[k, v] for k, v in results.items()]
    out["mismatching_files"] = sorted(bad)
    csv_path = os.path.join(case_dir, "out.csv")
    quiet(F.report_analysis, F.flatten_results(results), csv_path)
    out["scan_csv"] = b64(open(csv_path, "rb").read())
    return out



def tally_case(name, config, files, tmp, cores=1):
    """Tally stage only, over several files (config C5: single 6 bp index, file-level partitioning).
    The reference cannot run its matcher on single-index keys (F:104-107, F:306), so the tally of
    F:183-207 -- per-file dicts in file order and the merged "total" in first-appearance order -- is
    what pins this shape."""
    spec = synth.make_spec(config)
    case_dir = os.path.join(tmp, name)
    os.makedirs(case_dir)
    inputs, paths = {}, []
    for fname, g0, g1 in files:
        data = synth.generate_big(spec, g0, g1)
        p = os.path.join(case_dir, fname)
        write_gz(p, data)
        shutil.copy(p, os.path.join(HERE, f"{name}__{fname}"))
        inputs[fname] = [g0, g1]
        paths.append(p)
    counter = quiet(F.tally_barcodes, cores, paths, None)
    return {"config": config, "files": inputs, "cores": cores,
            "tally": {k: list(v.items()) for k, v in counter.items()}}


def cli_case(tmp):
    """The 3-file KAT of SURVEY Appendix A through the real CLI entry (frender_scan)."""
    d = os.path.join(tmp, "cli")
    os.makedirs(d)
    sheet = os.path.join(d, "sheet.csv")
    open(sheet, "w").write("Sample_ID,index,index2\nS1,AAAA,CCCC\nS2,GGGG,TTTT\n")

    def fq(keys):
        return "".join(f"@r{i}:1:2 1:N:0:{k}\nACGT\n+\nFFFF\n" for i, k in enumerate(keys)).encode()

    contents = {
        "S1_L001_R1_001.fastq.gz": fq(["AAAA+CCCC", "AAAT+CCCC", "GGGG+TTTT"]),
        "S2_L001_R1_001.fastq.gz": fq(["GGGG+TTTT", "GGGG+TTTT", "AAAA+TTTT"]),
        "Undetermined_L001_R1_001.fastq.gz": fq(["AAAA+CCCC", "AAAA+TTTT", "ACAC+ACAC"]),
    }
    files = []
    for name, data in contents.items():
        write_gz(os.path.join(d, name), data)
        files.append(os.path.join(d, name))
    cwd = os.getcwd()
    os.chdir(d)
    ns = argparse.Namespace(n=1, rc=True, c=1, s=None, o="kat", p=None, b=sheet, files=files)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        F.frender_scan(ns)
    os.chdir(cwd)
    produced = sorted(f for f in os.listdir(d) if f.startswith("frender-"))
    scan_csv = [f for f in produced if "scan-results" in f][0]
    calls_csv = [f for f in produced if "index-2-calls" in f][0]
    return {"sheet": open(sheet).read(), "files": {k: b64(v) for k, v in contents.items()},
            "args": {"n": 1, "rc": True, "o": "kat"},
            "scan_csv": b64(open(os.path.join(d, scan_csv), "rb").read()),
            "calls_csv": b64(open(os.path.join(d, calls_csv), "rb").read()),
            "stdout": buf.getvalue()}


def edge_inputs(tmp):
    """Tally edge cases: no trailing newline, CRLF, multi-member
    gzip, head sampling, empty file, extra header tokens."""
    d = os.path.join(tmp, "edge")
    os.makedirs(d)
    base = ("@a:1 1:N:0:AAAA+CCCC\nAC\n+\nFF\n@a:2 1:N:0:GGGG+TTTT x:y\nAC\n+\nFF\n"
            "@a:3 1:N:0:AAAA+CCCC\nAC\n+\nFF\n@a:4 2:Y:18:NNNN+ACGT\nAC\n+\nFF")
    cases = {
        "no_final_newline": (base.encode(), 1, None),
        "crlf": (base.replace("\n", "\r\n").encode() + b"\r\n", 1, None),
        "multi_member": ((base + "\n").encode() * 5, 4, None),
        "head_sample": ((base + "\n").encode() * 3, 1, 5),
        "empty": (b"", 1, None),
        "blank_tail_lines": ((base + "\n\n").encode(), 1, None),
        "no_space_header": (b"@nospace:AAAA+CCCC\nAC\n+\nFF\n", 1, None),
        "lowercase_key": (b"@a 1:N:0:acgt+CCCC\nAC\n+\nFF\n", 1, None),
    }
    out = {}
    for name, (data, members, sample) in cases.items():
        p = os.path.join(d, f"{name}_R1.fastq.gz")
        write_gz(p, data, members)
        rec = {"data": b64(data), "members": members, "sample": sample}
        try:
            counter = quiet(F.tally_barcodes, 1, [p], sample)
            rec["total"] = list(counter["total"].items())
        except Exception as exc:                                        # reference crashes: pin the type
            rec["raises"] = type(exc).__name__
        out[name] = rec
    return out


def reorder_for_demux(scan_csv_bytes):
    """Scan layout -> the layout parse_results_file asserts (SURVEY finding 1)."""
    import csv
    rows = list(csv.reader(io.StringIO(scan_csv_bytes.decode(), newline="")))
    h = rows[0]
    order = [h.index(c) for c in ["idx1", "idx2", "reads", "matched_idx1", "matched_idx2",
                                  "read_type", "sample_name", "demux_ok"]]
    buf = io.StringIO(newline="")
    csv.writer(buf).writerows([[r[i] for i in order] for r in rows])
    return buf.getvalue().encode()


def demux_case(name, scan, reads, tmp, flags, truncate_r2=None):
    spec = synth.make_spec(scan["config"], n_samples=scan["n_samples"])
    d = os.path.join(tmp, name)
    os.makedirs(d)
    r1 = synth.generate_big(spec, 0, reads, 1)
    r2 = synth.generate_big(spec, 0, reads, 2)
    if truncate_r2:
        # cut inside the sequence line of record `truncate_r2` so that the partial record still has
        # a complete header (F:777 pads it with "" and writes it out)
        pos = -1
        for _ in range(4 * truncate_r2 + 1):
            pos = r2.index(b"\n", pos + 1)
        r2 = r2[:pos + 11]
    p1 = os.path.join(d, "Undetermined_S0_L001_R1_001.fastq.gz")
    p2 = os.path.join(d, "Undetermined_S0_L001_R2_001.fastq.gz")
    write_gz(p1, r1)
    write_gz(p2, r2, 3)
    res = os.path.join(d, "results.csv")
    open(res, "wb").write(reorder_for_demux(base64.b64decode(scan["scan_csv"])))
    outdir = os.path.join(d, "out")
    ns = argparse.Namespace(no_index_hop=flags.get("i", False), no_ambiguous=flags.get("a", False),
                            no_undeter=False, no_samples=False, o=flags.get("o"), d=outdir, r=res,
                            files=[p1, p2])
    F.args = ns                                                         # open_files reads a global (F:672)
    quiet(F.frender_demux, ns)
    sinks = {}
    for f in sorted(os.listdir(outdir)):
        import hashlib
        raw = gzip.open(os.path.join(outdir, f), "rb").read()
        sinks[f] = {"bytes": len(raw), "sha256": hashlib.sha256(raw).hexdigest(),
                    "head": b64(raw[:160])}
    return {"scan_case": name.split("__")[0], "reads": reads, "flags": flags, "truncate_r2": truncate_r2,
            "results_csv": b64(open(res, "rb").read()), "sinks": sinks}


def main():
    tmp = tempfile.mkdtemp(prefix="frender_golden_")
    for f in os.listdir(HERE):
        if f.endswith(".gz") or f.endswith(".json"):
            os.remove(os.path.join(HERE, f))
    gold = {"matcher": matcher_kats(), "headers": header_kats(), "edge": edge_inputs(tmp),
            "cli3": cli_case(tmp), "scan": {}, "demux": {}}
    gold["scan"]["c1"] = scan_case("c1", "C1", 3000, 1, True, tmp)
    gold["scan"]["c2"] = scan_case("c2", "C2", 3000, 1, True, tmp, n_samples=48)
    gold["scan"]["c2n0"] = scan_case("c2n0", "C2", 1500, 0, False, tmp, n_samples=48)
    gold["scan"]["c3"] = scan_case("c3", "C3", 3000, 2, False, tmp)
    gold["scan"]["multi"] = scan_case(
        "multi", "C1", 0, 1, True, tmp, n_samples=12, prefix="S",
        files=[("S001_L001_R1_001.fastq.gz", 0, 700), ("S002_L001_R1_001.fastq.gz", 700, 1500),
               ("Undetermined_L001_R1_001.fastq.gz", 1500, 2400)])
    gold["scan"]["sampled"] = scan_case("sampled", "C1", 2000, 1, True, tmp, n_samples=24, sample=777)
    gold["demux"]["c1"] = demux_case("c1__demux", gold["scan"]["c1"], 3000, tmp, {})
    gold["demux"]["c1_ia"] = demux_case("c1__demux_ia", gold["scan"]["c1"], 3000, tmp,
                                        {"i": True, "a": True, "o": "tag"})
    gold["demux"]["c1_short_r2"] = demux_case("c1__demux_short", gold["scan"]["c1"], 3000, tmp, {},
                                              truncate_r2=2000)
    # the bench shape itself (BASELINE configs[1]): full 384-row sheet, -n 1 -rc
    gold["scan"]["c2_384"] = scan_case("c2_384", "C2", 20000, 1, True, tmp)
    # configs[3]: paired demux of that shape into 384 sample sinks + Index-hop / Ambiguous / Undetermined
    gold["demux"]["c4"] = demux_case("c2_384__demux", gold["scan"]["c2_384"], 5000, tmp, {})
    # configs[4]: single 6 bp index, many files (tally only: the reference's matcher needs index2)
    gold["tally"] = {"c5": tally_case(
        "c5", "C5", [(f"lane{k + 1}_S{k + 1}_L00{k + 1}_R1_001.fastq.gz", 600 * k, 600 * (k + 1)) for k in range(8)],
        tmp, cores=3)}
    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(gold, fh, indent=0, sort_keys=False)
    shutil.rmtree(tmp)
    print("wrote", os.path.join(HERE, "golden.json"),
          os.path.getsize(os.path.join(HERE, "golden.json")) // 1024, "KiB")


if __name__ == "__main__":
    main()
