"""The C oracle against the Python oracle / the reference's golden outputs (CPU)."""
import gzip
import os

import pytest

import c_oracle
import frender_oracle as O
from conftest import unb64

TYPES = ("undetermined", "index_hop", "demuxable", "ambiguous")


def test_c_key_rules(golden):
    for case in golden["headers"]:
        line = case["line"].rstrip("\n")
        got, reads = c_oracle.tally((line + "\nAC\n+\nFF\n").encode(), rule=0)
        assert got == {case["scan"]: 1} and reads == 1
        got, _ = c_oracle.tally((line + "\nAC\n+\nFF\n").encode(), rule=1)
        assert got == {case["demux"]: 1}


def test_c_tally_edges(golden):
    for name, case in golden["edge"].items():
        data = unb64(case["data"])
        if name == "crlf":
            continue                                   # universal newlines are the reader's job
        if "raises" in case:
            with pytest.raises(IndexError):
                c_oracle.tally(data, sample=case["sample"] or 0)
        else:
            got, _ = c_oracle.tally(data, sample=case["sample"] or 0)
            assert [list(x) for x in got.items()] == case["total"], name


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c2_384"])
def test_c_tally_and_classify_golden(golden, golden_dir, name):
    case = golden["scan"][name]
    (fname, _), = case["files"].items()
    raw = gzip.open(os.path.join(golden_dir, f"{name}__{fname}"), "rb").read()
    got, reads = c_oracle.tally(raw)
    assert [list(x) for x in got.items()] == case["tally"]["total"]
    idx = dict(case["indexes"])
    if case["rc"]:
        idx["idx2"] = case["oriented_idx2"]
    res = c_oracle.classify_all(list(got), idx, case["n"])
    for (key, want), (m1, m2, kind, row) in zip(case["final"], res):
        assert TYPES[kind] == want["read_type"], key
        assert (idx["idx1"][m1] if m1 >= 0 else "") == want["matched_idx1"]
        assert (idx["idx2"][m2] if m2 >= 0 else "") == want["matched_idx2"]
        assert (idx["id"][row] if row >= 0 else "") == want["sample_name"]


@pytest.mark.parametrize("name", ["c1", "c2", "c2_384", "multi"])
def test_c_rc_first_pass_golden(golden, name):
    """`-rc` first pass (F:294-351) and the per-sample orientation call (F:354-388) of the C oracle against
    what the reference produced."""
    case = golden["scan"][name]
    keys = [k for k, _ in case["tally"]["total"]]
    counts = [n for _, n in case["tally"]["total"]]
    idx = case["indexes"]
    res = c_oracle.classify_all_rc(keys, idx, case["n"])
    rc_idx2 = [O.reverse_complement(s) for s in idx["idx2"]]
    pick = lambda table, row: table[row] if row >= 0 else ""
    for (key, want), r in zip(case["first_pass"], res):
        got = {"matched_idx1": pick(idx["idx1"], r[0]), "matched_idx2": pick(idx["idx2"], r[1]),
               "read_type": TYPES[r[2]], "sample_name": pick(idx["id"], r[3]),
               "matched_rc_idx2": pick(rc_idx2, r[4]), "rc_read_type": TYPES[r[5]],
               "rc_sample_name": pick(idx["id"], r[6])}
        assert {f: want[f] for f in got} == got, key
    calls = c_oracle.rc_calls(keys, counts, res, idx)
    assert calls == {k: (v["call"], v["reads_f"], v["reads_rc"]) for k, v in case["rc_calls"]}
