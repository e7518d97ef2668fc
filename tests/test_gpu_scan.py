"""GPU parity: CUDA path (through the C-ABI) against the oracle and the reference's golden outputs."""
import gzip
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from frender_b200.engine import Context
    c = Context(0, table_log2=18)
    yield c
    c.close()


def test_pack_roundtrip():
    from frender_b200.engine import pack_keys, unpack_keys
    keys = ["AAAAAAAA+GGGGGGGG", "ACGTN+NTGCA", "", "ACGTAC", "NNNNNNNNNN+ACGTACGTAC", "A+C+G"]
    assert unpack_keys(pack_keys(keys)) == keys


def test_header_rules(ctx, golden):
    """Every header KAT through the kernel, both key rules (F:169, F:778)."""
    import frender_b200._lib as L
    from frender_b200.engine import FrbError
    for case in golden["headers"]:
        line = case["line"].rstrip("\n")
        rec = (line + "\nACGT\n+\nFFFF\n").encode()
        for rule, want in ((L.RULE_SCAN, case["scan"]), (L.RULE_DEMUX, case["demux"])):
            ctx.reset()
            ok = all(ch in "ACGTN+" for ch in want) and len(want) <= 21
            if ok:
                reads, uniq = ctx.scan_bytes(rec * 3, rule=rule)
                assert reads == 3 and uniq == 1
                assert ctx.counter()["total"] == {want: 3}, (case, rule)
            else:
                with pytest.raises(FrbError):
                    ctx.scan_bytes(rec * 3, rule=rule)


def test_tally_edges(ctx, golden, tmp_path):
    from conftest import unb64
    from frender_b200.engine import FrbError, tally_barcodes
    for name, case in golden["edge"].items():
        data = unb64(case["data"])
        path = tmp_path / f"{name}_R1.fastq.gz"
        step = (len(data) + case["members"] - 1) // case["members"] if case["members"] > 1 else max(len(data), 1)
        with open(path, "wb") as fh:
            if not data:
                fh.write(gzip.compress(b""))
            for off in range(0, len(data), step):
                fh.write(gzip.compress(data[off:off + step]))
        if "raises" in case or name == "lowercase_key":
            with pytest.raises(FrbError):
                tally_barcodes(1, [path], case["sample"], ctx=ctx)
        else:
            got = tally_barcodes(1, [path], case["sample"], ctx=ctx)
            assert [list(x) for x in got["total"].items()] == case["total"], name


@pytest.mark.parametrize("name", ["c1", "c2", "c2n0", "c3", "multi", "sampled", "c2_384"])
def test_scan_stages_golden(ctx, golden, golden_dir, name):
    """tally -> first pass -> orientation call -> second pass against the reference's outputs."""
    from frender_b200.engine import tally_barcodes
    case = golden["scan"][name]
    files = [os.path.join(golden_dir, f"{name}__{f}") for f in case["files"]]
    counter = tally_barcodes(1, files, case["sample"], ctx=ctx)
    counter = {k.split("__", 1)[-1]: v for k, v in counter.items()}
    assert {k: [list(x) for x in v.items()] for k, v in counter.items()} == case["tally"]
    assert list(counter) == list(case["tally"])
    indexes = case["indexes"]
    first, raw = ctx.process(None, indexes, case["n"], case["rc"])
    assert [[k, v] for k, v in first.items()] == case["first_pass"]
    results, calls, oriented = ctx.analyze(indexes, case["n"], case["rc"])
    if case["rc"]:
        assert [[k, v] for k, v in calls.items()] == case["rc_calls"]
        assert oriented == case["oriented_idx2"]
    want = [[k, {f: v for f, v in rec.items() if f != "demux_ok"}] for k, rec in case["final"]]
    assert [[k, v] for k, v in results.items()] == want


def test_matcher_kats(ctx, golden):
    from frender_b200.engine import process
    m = golden["matcher"]
    idx = {"id": m["id"], "idx1": m["idx1"], "idx2": m["idx2"]}
    for c in m["classify"]:
        got = process(1, {c["idx1"] + "+" + c["idx2"]: 5}, idx, c["n"], False, ctx=ctx)
        want = dict(c["want"], reads=5)
        assert got == {c["idx1"] + "+" + c["idx2"]: want}, c
    for c in m["rc"]:
        idx = {"id": c["id"], "idx1": c["idx1"], "idx2": c["idx2"]}
        got = process(1, {c["key"]: c["reads"]}, idx, c["n"], c["rc_mode"], ctx=ctx)
        assert got == {c["key"]: c["want"]} and list(got[c["key"]]) == list(c["want"]), c


def test_matcher_random_sheets_vs_oracle(ctx):
    """Random sheets (duplicated index values as in combinatorial designs, 6/8/10 bp, n = 0..4) and keys a few
    substitutions away from sheet rows, N included: the device matcher (candidate tables for n <= 3, sweep
    beyond) must give the oracle's classification, matched strings, sample names and orientation calls."""
    import random

    import frender_oracle as O
    from frender_b200.engine import process
    rnd = random.Random(23)
    for trial in range(24):
        l = rnd.choice([6, 8, 10])
        n = rnd.choice([0, 1, 1, 2, 3, 4])
        rows = rnd.choice([3, 17, 96])
        pool1 = ["".join(rnd.choice("ACGT") for _ in range(l)) for _ in range(max(2, rows // rnd.choice([1, 1, 4])))]
        pool2 = ["".join(rnd.choice("ACGT") for _ in range(l)) for _ in range(max(2, rows // rnd.choice([1, 1, 4])))]
        idx = {"id": [f"s{r % max(1, rows - 2)}" for r in range(rows)],          # a few repeated sample names
               "idx1": [rnd.choice(pool1) for _ in range(rows)], "idx2": [rnd.choice(pool2) for _ in range(rows)]}
        counter = {}
        for _ in range(400):
            r = rnd.randrange(rows)
            a = list(idx["idx1"][r])
            b = list(rnd.choice([idx["idx2"][rnd.randrange(rows)], O.reverse_complement(idx["idx2"][r])]))
            for word in (a, b):
                for _ in range(rnd.choice([0, 0, 1, 2, 3])):
                    word[rnd.randrange(l)] = rnd.choice("ACGTN")
            counter.setdefault("".join(a) + "+" + "".join(b), rnd.randrange(1, 50))
        rc_mode = trial % 2 == 0
        got = process(1, counter, idx, n, rc_mode, ctx=ctx)
        want = O.process(1, counter, idx, n, rc_mode)
        assert got == want, (trial, l, n, rows, rc_mode)
        if rc_mode:
            res, calls, _ = ctx.analyze(idx, n, True, counter=counter)
            want_res, want_calls, _ = O.scan_analysis(1, {"total": counter}, idx, n, True)
            assert res == want_res and calls == want_calls, (trial, "second pass")


def test_length_mismatch_raises(ctx):
    from frender_b200.engine import process
    idx = {"id": ["a"], "idx1": ["AAAAAAAA"], "idx2": ["CCCCCCCC"]}
    with pytest.raises(AssertionError):
        process(1, {"AAAA+CCCCCCCC": 1}, idx, 1, False, ctx=ctx)
    with pytest.raises(ValueError):                    # no '+': unpacking fails as in F:306
        process(1, {"AAAAAAAA": 1}, idx, 1, False, ctx=ctx)
    with pytest.raises(AssertionError):
        process(1, {"AAAAAAAA+CCCCCCC": 1}, idx, 1, False, ctx=ctx)


@pytest.mark.parametrize("chunk", [None, 4096, 40000, 1 << 20])
def test_chunking_invariance(ctx, chunk):
    """Any chunking of the stream gives the oracle's counts in the oracle's order."""
    import frender_oracle as O
    from frender_b200 import synth
    spec = synth.make_spec("C2", n_samples=32)
    data = synth.generate(spec, 100, 20100)
    want, visited = O.tally_text(data.decode().splitlines(keepends=True))
    ctx.reset()
    reads, uniq = ctx.scan_bytes(data, chunk=chunk)
    assert reads == visited == 20000 and uniq == len(want)
    assert list(ctx.counter()["total"].items()) == list(want.items())


def test_keys_per_read_and_offsets(ctx):
    """Per-read packed keys and record offsets from the resident-buffer entry point."""
    import frender_b200._lib as L
    from frender_b200 import synth
    from frender_b200.engine import C, unpack_keys
    spec = synth.make_spec("C1", n_samples=16)
    n = 30000
    data = synth.generate(spec, 0, n)
    h = ctx._h
    ctx.reset()
    dbuf, dkeys, doffs = C.c_void_p(), C.c_void_p(), C.c_void_p()
    ctx._ck(L.lib.frb_dev_alloc(h, len(data) + 64, C.byref(dbuf)))
    ctx._ck(L.lib.frb_dev_alloc(h, n * 8, C.byref(dkeys)))
    ctx._ck(L.lib.frb_dev_alloc(h, n * 8, C.byref(doffs)))
    arr = np.frombuffer(data, np.uint8)
    ctx._ck(L.lib.frb_h2d(h, dbuf, arr.ctypes.data_as(C.c_void_p), len(data)))
    ctx._ck(L.lib.frb_scan_begin(h, 0, 0))
    ctx._ck(L.lib.frb_scan_chunk_dev(h, dbuf, len(data), 0, L.RULE_SCAN, dkeys, doffs))
    reads, uniq = C.c_uint64(), C.c_uint64()
    ctx._ck(L.lib.frb_scan_end(h, C.byref(reads), C.byref(uniq)))
    keys, offs = np.empty(n, np.uint64), np.empty(n, np.uint64)
    ctx._ck(L.lib.frb_d2h(h, keys.ctypes.data_as(C.c_void_p), dkeys, n * 8))
    ctx._ck(L.lib.frb_d2h(h, offs.ctypes.data_as(C.c_void_p), doffs, n * 8))
    for p in (dbuf, dkeys, doffs):
        ctx._ck(L.lib.frb_dev_free(h, p))
    assert reads.value == n
    assert unpack_keys(keys) == synth.keys_of(spec, 0, n)
    starts = np.flatnonzero(arr == 10)[3::4][:-1] + 1
    assert offs[0] == 0 and (offs[1:] == starts).all()


def test_short_lines_serial_fallback(ctx):
    """Records so short that a tile holds more newlines than the kernel's position list: the exact
    serial fallback must give the oracle's tally (and per-read keys) too."""
    import random

    import frender_oracle as O
    rnd = random.Random(7)
    keys = ["".join(rnd.choice("ACGTN") for _ in range(2)) + "+" + "".join(rnd.choice("ACGT") for _ in range(2))
            for _ in range(40)]
    recs = "".join(f"@r 1:N:0:{rnd.choice(keys)}\nA\n+\nF\n" for _ in range(20000))
    want, visited = O.tally_text(recs.splitlines(keepends=True))
    for chunk in (None, 50_000):
        ctx.reset()
        reads, uniq = ctx.scan_bytes(recs.encode(), chunk=chunk)
        assert reads == visited == 20000 and uniq == len(want)
        assert list(ctx.counter()["total"].items()) == list(want.items())


def test_line_phase_is_by_count_not_by_content(ctx):
    """The kernel guesses which lines are headers from the text ('@' ... '+' ... '@') and checks the guess
    against the newline count before anything reaches the table.  Here the text says one thing and the
    count another (F:161-169 takes every 4th line, whatever it looks like): the result must follow the
    count."""
    import random

    import frender_oracle as O
    rnd = random.Random(11)
    keys = ["".join(rnd.choice("ACGT") for _ in range(8)) for _ in range(50)]
    # by count the headers are the '@r' lines; by content ('@', then '+' two lines on) the '@q' lines
    recs = "".join(f"@r{i} 1:N:0:{rnd.choice(keys)}\n@q{i} 1:N:0:{rnd.choice(keys)}\n{'ACGT' * 30}\n+{'F' * 100}\n"
                   for i in range(30000))
    want, visited = O.tally_text(recs.splitlines(keepends=True))
    for chunk in (None, 1 << 20):
        ctx.reset()
        reads, uniq = ctx.scan_bytes(recs.encode(), chunk=chunk)
        assert reads == visited == 30000 and uniq == len(want)
        assert list(ctx.counter()["total"].items()) == list(want.items())
    # and a file that only goes wrong half way: a stray line shifts the phase of everything after it
    good = "".join(f"@r{i} 1:N:0:{rnd.choice(keys)}\n{'ACGT' * 30}\n+\n{'F' * 120}\n" for i in range(20000))
    lines = good.splitlines(keepends=True)
    shifted = "".join(lines[:40000]) + "@x 1:N:0:ACGTACGT\n" + "".join(
        f"@s{i} 1:N:0:{rnd.choice(keys)}\n{'ACGT' * 30}\n+\n@t{i} 1:N:0:{rnd.choice(keys)}\n" for i in range(15000))
    want, visited = O.tally_text(shifted.splitlines(keepends=True))
    ctx.reset()
    reads, uniq = ctx.scan_bytes(shifted.encode())
    assert reads == visited and uniq == len(want)
    assert list(ctx.counter()["total"].items()) == list(want.items())


def test_long_lines(ctx):
    """Header lines longer than the 512-byte halo (their start is outside the staged bytes) and reads of
    several kilobytes (a tile holds too few lines for a phase guess): both take the exact slow paths."""
    import random

    import frender_oracle as O
    rnd = random.Random(5)
    keys = ["".join(rnd.choice("ACGTN") for _ in range(10)) + "+" + "".join(rnd.choice("ACGT") for _ in range(10))
            for _ in range(30)]
    recs = "".join(f"@{'x' * rnd.choice((30, 700, 1500))}:{i} 1:N:0:{rnd.choice(keys)}\n{'ACGT' * rnd.choice((40, 1500))}\n+\n"
                   f"{'F' * 100}\n" for i in range(4000))
    want, visited = O.tally_text(recs.splitlines(keepends=True))
    for chunk in (None, 1 << 20):
        ctx.reset()
        reads, uniq = ctx.scan_bytes(recs.encode(), chunk=chunk)
        assert reads == visited == 4000 and uniq == len(want)
        assert list(ctx.counter()["total"].items()) == list(want.items())


def test_single_index_c5(ctx, tmp_path):
    """Config 5 shape: 6 bp single index.  The tally is pinned by the oracle (the reference can run it,
    F:154-207); the matcher for single-index sheets is an extension (the reference cannot: F:104-107,
    F:306) checked against the oracle's statement of it (parity unpinned)."""
    import frender_oracle as O
    from frender_b200 import synth
    from frender_b200.engine import tally_barcodes
    spec = synth.make_spec("C5")
    files = []
    for i in range(3):
        p = tmp_path / f"S{i}_L001_R1_001.fastq.gz"
        p.write_bytes(gzip.compress(synth.generate(spec, i * 4000, (i + 1) * 4000), 1))
        files.append(p)
    want = O.tally_barcodes(1, files)
    got = tally_barcodes(1, files, ctx=ctx)
    assert {k: list(v.items()) for k, v in got.items()} == {k: list(v.items()) for k, v in want.items()}
    idx = spec.indexes()
    assert idx["idx2"] is None
    res, _ = ctx.process(None, idx, 0, False)
    for key, rec in res.items():
        exp = O.match_single_index(key, idx["idx1"], idx["id"], 0)
        assert (rec["matched_idx1"], rec["read_type"], rec["sample_name"]) == \
               (exp["matched_idx1"], exp["read_type"], exp["sample_name"]), key
    res1, _ = ctx.process(None, idx, 1, False)
    for key, rec in list(res1.items())[:500]:
        exp = O.match_single_index(key, idx["idx1"], idx["id"], 1)
        assert (rec["matched_idx1"], rec["read_type"], rec["sample_name"]) == \
               (exp["matched_idx1"], exp["read_type"], exp["sample_name"]), key


@pytest.mark.parametrize("streams", [1, 4])
def test_tally_c5_golden(ctx, golden, golden_dir, streams):
    """BASELINE configs[4] shape against what the REFERENCE tallied (F:183-207): eight single-index files,
    per-file dicts in file order and "total" in first-appearance order; also through the `-c N` path (N files
    scanned at the same time on N contexts of this GPU)."""
    from frender_b200.cli import ScanTables, scan_files_concurrent
    from frender_b200.engine import tally_barcodes, unpack_keys
    case = golden["tally"]["c5"]
    files = [os.path.join(golden_dir, f"c5__{f}") for f in case["files"]]
    if streams == 1:
        got = tally_barcodes(1, files, ctx=ctx)
        got = {k.split("__", 1)[-1]: list(v.items()) for k, v in got.items()}
    else:
        per_file, total = scan_files_concurrent(files, None, streams, 0, 16, ctx)
        got = {"total": list(zip(unpack_keys(total[0]), total[1].tolist()))}
        for i, f in enumerate(case["files"]):
            got[f] = list(zip(unpack_keys(per_file[i][2]), per_file[i][3].tolist()))
    assert list(got) == list(case["tally"])
    assert {k: [list(x) for x in v] for k, v in got.items()} == case["tally"]


def test_empty_sheet_is_all_undetermined(ctx):
    """A sheet with no rows: get_indexes_of_approx_matches returns [] (F:220-234), so every key is undetermined
    (F:280-284) whatever -n is."""
    from frender_b200.engine import process
    idx = {"id": [], "idx1": [], "idx2": []}
    for n in (0, 1):
        got = process(1, {"ACGTACGT+TTTTAAAA": 3, "NNNNNNNN+ACGTACGT": 1}, idx, n, False, ctx=ctx)
        assert [v["read_type"] for v in got.values()] == ["undetermined", "undetermined"]
        assert all(v["matched_idx1"] == "" and v["sample_name"] == "" for v in got.values())


def test_demux_ok_on_the_device(ctx, golden, golden_dir):
    """frb_demux_ok (F:504-564) from the context's own per-file lists and from host lists: the reference's flags."""
    import numpy as np
    import frender_b200.cli as cli
    case = golden["scan"]["multi"]
    names = list(case["files"])
    ctx.reset()
    for i, f in enumerate(names):
        ctx.scan_gz(os.path.join(golden_dir, f"multi__{f}"), i)
    ctx.file_names[:] = names
    tables = cli.ScanTables.from_ctx(ctx, names)
    results, _, _ = ctx.analyze(case["indexes"], case["n"], case["rc"])
    assert list(results) == [k for k, _ in case["final"]]
    m = cli.class_file_matrix(names, case["indexes"]["id"], case["prefix"] or "")
    ok_dev, bad_dev, err = ctx.demux_ok(m, None)
    ok_host, bad_host, _ = ctx.demux_ok(m, tables.files)
    want = [rec["demux_ok"] for _, rec in case["final"]]
    assert err is None and ok_dev.tolist() == want and ok_host.tolist() == want
    assert bad_dev.tolist() == bad_host.tolist()
    assert sorted(names[f] for f in np.flatnonzero(bad_dev)) == sorted(case["mismatching_files"])
    # a sample name that is not a valid pattern only matters when a key of that sample sits in a file
    rows = sorted({case["indexes"]["id"].index(rec["sample_name"]) for _, rec in case["final"] if rec["sample_name"]})
    m2 = m.copy()
    m2[4:, :] = 2
    assert ctx.demux_ok(m2, None)[2] == rows[0]
    unused = [r for r in range(len(case["indexes"]["id"])) if r not in rows]
    if unused:
        m3 = m.copy()
        m3[4 + unused[0], :] = 2
        assert ctx.demux_ok(m3, None)[2] is None


def test_scan_gz_batch_equals_one_by_one(ctx, golden, golden_dir, tmp_path):
    """A run of small files as ONE multi-member gzip stream (frb_scan_gz_batch): per-file tallies, the total and its
    order are what the reference tallied file by file (configs[4] golden), whatever the files look like inside --
    several members, stored blocks, no newline at the end, no reads at all."""
    import gzip
    import zlib
    import frender_oracle as O
    from frender_b200 import synth
    case = golden["tally"]["c5"]
    files = [os.path.join(golden_dir, f"c5__{f}") for f in case["files"]]
    ctx.reset()
    done = ctx.scan_gz_batch(files, 0)
    assert done is not None, "the library declined the run"
    got = {k.split("__", 1)[-1]: list(v.items()) for k, v in ctx.counter().items()}
    assert list(got) == list(case["tally"])
    assert {k: [list(x) for x in v] for k, v in got.items()} == case["tally"]
    assert [r for r, _, _ in done] == [sum(n for _, n in case["tally"][f]) for f in case["files"]]
    # odd files in one run, against the oracle file by file
    spec = synth.make_spec("C1", n_samples=12)
    texts = [synth.generate(spec, 0, 3000), synth.generate(spec, 3000, 3001)[:-1], b"", synth.generate(spec, 4000, 9000),
             synth.generate(spec, 9000, 9500)]
    blobs = [gzip.compress(texts[0], 6), gzip.compress(texts[1], 9), gzip.compress(b""),
             gzip.compress(texts[3][:200_000], 1) + gzip.compress(texts[3][200_000:], 6)]
    stored = zlib.compressobj(0, zlib.DEFLATED, 31)
    blobs.append(stored.compress(texts[4]) + stored.flush())
    paths = []
    for i, blob in enumerate(blobs):
        p = tmp_path / f"f{i}_R1_001.fastq.gz"
        p.write_bytes(blob)
        paths.append(p)
    ctx.reset()
    done = ctx.scan_gz_batch(paths, 0)
    assert done is not None, "the library declined the run"
    want = O.tally_barcodes(1, paths)
    have = ctx.counter()
    assert {k: list(v.items()) for k, v in have.items()} == {k: list(v.items()) for k, v in want.items()}
    assert [raw for _, _, raw in done] == [len(t) for t in texts]
    # a damaged file in the run: declined as a whole, nothing kept; one by one names the file
    from frender_b200.engine import FrbError
    bad = tmp_path / "bad_R1_001.fastq.gz"
    bad.write_bytes(blobs[0][:-8] + bytes([blobs[0][-8] ^ 1]) + blobs[0][-7:])
    ctx.reset()
    assert ctx.scan_gz_batch([paths[0], bad, paths[4]], 0) is None
    assert ctx.counter() == {"total": {}} or list(ctx.counter()) == ["total"]
    with pytest.raises(FrbError, match="bad_R1_001"):
        ctx.scan_gz(bad, 0)
