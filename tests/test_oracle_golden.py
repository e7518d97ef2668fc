"""The oracle against outputs of the reference itself (tests/golden, made by
make_golden.py which imports /root/reference/frender.py).  CPU only."""
import gzip
import hashlib
import io
import os

import pytest

import frender_oracle as O
from conftest import unb64
from frender_b200 import synth


def test_header_rules(golden):
    for case in golden["headers"]:
        assert O.scan_key(case["line"]) == case["scan"]
        assert O.demux_key(case["line"]) == case["demux"]


def test_matcher_kats(golden):
    m = golden["matcher"]
    for c in m["classify"]:
        assert O.classify(c["idx1"], c["idx2"], m["idx1"], m["idx2"], m["id"], c["n"]) == c["want"]
    for c in m["approx"]:
        assert O.approx_match_rows(c["q"], m["idx1"], c["n"]) == c["want"]
    for s, want in m["revcomp"]:
        assert O.reverse_complement(s) == want
    for c in m["rc"]:
        got = O.classify_with_rc(c["key"], c["reads"], c["idx1"], c["idx2"], c["id"], c["n"], c["rc_mode"])
        assert got == c["want"] and list(got) == list(c["want"])


def test_length_mismatch_message():
    with pytest.raises(AssertionError, match="Barcode aaaa doesn't match length of supplied barcode aaaaaaaa"):
        O.approx_match_rows("AAAA", ["AAAAAAAA"], 1)


def test_tally_edges(golden, tmp_path):
    for name, case in golden["edge"].items():
        data = unb64(case["data"])
        path = tmp_path / f"{name}_R1.fastq.gz"
        step = (len(data) + case["members"] - 1) // case["members"] if case["members"] > 1 else max(len(data), 1)
        with open(path, "wb") as fh:
            if not data:
                fh.write(gzip.compress(b""))
            for off in range(0, len(data), step):
                fh.write(gzip.compress(data[off:off + step]))
        if "raises" in case:
            with pytest.raises(Exception) as info:
                O.tally_barcodes(1, [path], case["sample"])
            assert type(info.value).__name__ == case["raises"]
        else:
            got = O.tally_barcodes(1, [path], case["sample"])
            assert [list(x) for x in got["total"].items()] == case["total"], name


@pytest.mark.parametrize("name", ["c1", "c2", "c2n0", "c3", "multi", "sampled", "c2_384"])
def test_scan_stages(golden, golden_dir, name):
    case = golden["scan"][name]
    files = [os.path.join(golden_dir, f"{name}__{f}") for f in case["files"]]
    counter = O.tally_barcodes(1, files, case["sample"])
    # fixture files carry a "<case>__" prefix; the reference saw the bare names
    counter = {k.split("__", 1)[-1]: v for k, v in counter.items()}
    assert {k: [list(x) for x in v.items()] for k, v in counter.items()} == case["tally"]
    assert list(counter) == list(case["tally"])
    indexes = case["indexes"]
    first = O.process(1, counter["total"], indexes, case["n"], case["rc"])
    assert [[k, v] for k, v in first.items()] == case["first_pass"]
    results, calls, used = O.scan_analysis(2, counter, dict(indexes), case["n"], case["rc"])
    if case["rc"]:
        assert [[k, v] for k, v in calls.items()] == case["rc_calls"]
        assert used["idx2"] == case["oriented_idx2"]
        assert O.rc_calls_csv_bytes(calls, indexes) == unb64(case["rc_calls_csv"])
    results, bad = O.demux_ok(counter, results, case["prefix"])
    assert [[k, v] for k, v in results.items()] == case["final"]
    assert sorted(bad) == case["mismatching_files"]
    assert O.scan_csv_bytes(results) == unb64(case["scan_csv"])


def test_sheet_reader(golden, tmp_path):
    case = golden["scan"]["c1"]
    p = tmp_path / "SampleSheet.csv"
    p.write_text(case["sheet_csv"])
    assert O.read_sheet(p) == case["indexes"]
    p.write_text(golden["cli3"]["sheet"])
    assert O.read_sheet(p) == {"id": ["S1", "S2"], "idx1": ["AAAA", "GGGG"], "idx2": ["CCCC", "TTTT"]}


def test_synth_matches_fixture(golden, golden_dir):
    """The committed inputs are exactly what the generator still produces."""
    case = golden["scan"]["c1"]
    spec = synth.make_spec(case["config"], n_samples=case["n_samples"])
    (fname, (g0, g1)), = case["files"].items()
    raw = gzip.open(os.path.join(golden_dir, f"c1__{fname}"), "rb").read()
    assert raw == synth.generate_big(spec, g0, g1)
    assert spec.sheet_csv() == case["sheet_csv"]
    keys = synth.keys_of(spec, g0, g1)
    tally = {}
    for k in keys:
        tally[k] = tally.get(k, 0) + 1
    assert [list(x) for x in tally.items()] == case["tally"]["total"]


@pytest.mark.parametrize("name", ["c1", "c1_ia", "c1_short_r2", "c4"])
def test_demux_streams(golden, tmp_path, name):
    case = golden["demux"][name]
    scan = golden["scan"][case["scan_case"]]
    spec = synth.make_spec(scan["config"], n_samples=scan["n_samples"])
    r1 = synth.generate_big(spec, 0, case["reads"], 1)
    r2 = synth.generate_big(spec, 0, case["reads"], 2)
    if case["truncate_r2"]:
        pos = -1
        for _ in range(4 * case["truncate_r2"] + 1):
            pos = r2.index(b"\n", pos + 1)
        r2 = r2[:pos + 11]
    res = tmp_path / "results.csv"
    res.write_bytes(unb64(case["results_csv"]))
    table = O.parse_results_file(res)
    flags = case["flags"]
    roles = O.sink_names(table, index_hop=not flags.get("i", False), ambiguous=not flags.get("a", False))
    streams = O.route_pairs(io.StringIO(r1.decode()), io.StringIO(r2.decode()), table, roles)
    infix = (flags["o"] + "_") if flags.get("o") else ""
    got = {}
    for sink, (a, b) in streams.items():
        got[f"{sink}_frender-demux_{infix}R1.fq.gz"] = a
        got[f"{sink}_frender-demux_{infix}R2.fq.gz"] = b
    assert sorted(got) == sorted(case["sinks"])
    for fname, want in case["sinks"].items():
        assert len(got[fname]) == want["bytes"], fname
        assert hashlib.sha256(got[fname]).hexdigest() == want["sha256"], fname


def test_tally_c5_many_files(golden, golden_dir):
    """BASELINE configs[4] shape: single 6 bp index, eight files; per-file dicts and "total" of F:183-207."""
    case = golden["tally"]["c5"]
    files = [os.path.join(golden_dir, f"c5__{f}") for f in case["files"]]
    for cores in (1, 2):
        counter = O.tally_barcodes(cores, files)
        counter = {k.split("__", 1)[-1]: v for k, v in counter.items()}
        assert list(counter) == list(case["tally"])
        assert {k: [list(x) for x in v.items()] for k, v in counter.items()} == case["tally"]
