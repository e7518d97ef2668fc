"""The C oracle against the Python oracle / the reference's golden outputs (CPU)."""
import gzip
import os
import sys

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


def test_c_route_digests_equal_the_python_oracle():
    """oracle_route (the demux loop F:774-810 in C, per-sink byte counts and order-sensitive digests) against
    route_pairs, which the reference's golden demux outputs pin: a short mate, a partial tail, an unknown key."""
    import random

    import numpy as np
    import c_oracle
    import frender_oracle as O
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_gpu_route import make_pair
    role_of = {"index_hop": "#hop", "ambiguous": "#amb", "undetermined": "#und"}
    for seed, kwargs in ((1, {}), (2, {"crop_tail": 9}), (3, {"r2_records": 1700})):
        t1, t2, table = make_pair(random.Random(seed), 2500, **kwargs)
        roles = O.sink_names(table)
        names = sorted({n for n in roles.values() if n})
        sid = {n: i for i, n in enumerate(names)}
        keys = list(table)
        routes = [sid[roles[table[k][1]] if table[k][0] == "demuxable" else roles[role_of[table[k][0]]]] for k in keys]
        want = O.route_pairs(t1.splitlines(keepends=True), t2.splitlines(keepends=True), table, roles)
        got = c_oracle.route_sums(t1.encode(), 0, t2.encode(), 0, keys, routes, len(names))
        for name, i in sid.items():
            for mate, stream in enumerate(want[name]):
                h = np.full(1, c_oracle.FNV_BASIS, np.uint64)
                if stream:
                    c_oracle.fnv1a_segments(h, np.frombuffer(stream, np.uint8).ctypes.data, [0, len(stream)])
                assert got[i][f"bytes{mate + 1}"] == len(stream) and got[i][f"hash{mate + 1}"] == h[0], (seed, name, mate)
        assert int(got["records"].sum()) == sum(v[1].count(b"\n@M0:") + (1 if v[1] else 0) for v in want.values())
    t1, t2, table = make_pair(random.Random(7), 100)
    headers = t2.splitlines()[::4]
    lost = headers[40].rsplit(":", 1)[1]
    first = next(i for i, hd in enumerate(headers) if hd.rsplit(":", 1)[1] == lost)
    keys = [k for k in table if k != lost]
    with pytest.raises(SystemExit, match=f"record {first} "):
        c_oracle.route_sums(t1.encode(), 0, t2.encode(), 0, keys, [0] * len(keys), 1)
