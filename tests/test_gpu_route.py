"""The demux stream (frb_route_push / frb_route_pop) against the oracle's route_pairs (F:774-810): chunks cut anywhere
and differently for the two mates, carries on the device, a mate that ends early, a trailing partial record."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KINDS = ["demuxable", "index_hop", "ambiguous", "undetermined"]


def make_pair(rng, n, n_keys=40, r2_records=None, crop_tail=0):
    bases = "ACGTN"
    keys = ["".join(rng.choice(bases) for _ in range(8)) + "+" + "".join(rng.choice(bases) for _ in range(8))
            for _ in range(n_keys)]
    r1, r2 = [], []
    for i in range(n):
        k = rng.choice(keys)
        l1, l2 = rng.randint(1, 300), rng.randint(1, 300)
        name = f"@M0:{i}:FC:1:{rng.randint(1, 99999)}:{rng.randint(1, 99999)}"
        r1.append(f"{name} 1:N:0:{k}\n{'A' * l1}\n+\n{'F' * l1}\n")
        r2.append(f"{name} 2:N:0:{k}\n{'C' * l2}\n+\n{'#' * l2}\n")
    if r2_records is not None:
        r2 = r2[:r2_records]
    t1, t2 = "".join(r1), "".join(r2)
    if crop_tail:
        t2 = t2[:-crop_tail]
    table = {k: (KINDS[j % 4], f"S{j % 7}" if j % 4 == 0 else "") for j, k in enumerate(keys)}
    return t1, t2, table


def stream(ctx, t1, t2, table, roles, cut1, cut2):
    """Push the two texts in pieces of cut1() / cut2() bytes, two chunks in flight; returns {sink: (R1, R2)}."""
    from frender_b200.engine import pack_keys
    names = sorted({n for n in roles.values() if n})
    sid = {n: i for i, n in enumerate(names)}
    role_of = {"index_hop": "#hop", "ambiguous": "#amb", "undetermined": "#und"}
    keys = list(table)
    routes = np.array([sid[roles[table[k][1]] if table[k][0] == "demuxable" else roles[role_of[table[k][0]]]]
                       for k in keys], np.uint32)
    ctx.route_load(pack_keys(keys), routes, len(names))
    ctx.route_reset()
    out = {n: (bytearray(), bytearray()) for n in names}
    b1, b2 = t1.encode(), t2.encode()
    p1 = p2 = 0
    queued = 0
    total = 0

    def take():
        nonlocal total
        o1, o2, off1, off2, pairs, _, _ = ctx.route_pop()
        total += pairs
        for n, i in sid.items():
            out[n][0].extend(o1[off1[i]:off1[i + 1]].tobytes())
            out[n][1].extend(o2[off2[i]:off2[i + 1]].tobytes())

    while True:
        n1, n2 = cut1(), cut2()
        d1, d2 = b1[p1:p1 + n1], b2[p2:p2 + n2]
        p1, p2 = p1 + len(d1), p2 + len(d2)
        final = (1 if p1 >= len(b1) else 0) | (2 if p2 >= len(b2) else 0)
        ctx.route_push(d1, d2, final)
        queued += 1
        if queued == 2:
            take()
            queued -= 1
        if final == 3:
            break
    while queued:
        take()
        queued -= 1
    return {n: (bytes(a), bytes(b)) for n, (a, b) in out.items()}, total


@pytest.fixture(scope="module")
def ctx():
    from frender_b200.engine import Context
    c = Context(0, table_log2=16)
    yield c
    c.close()


@pytest.mark.parametrize("case", ["even", "r2_short", "r1_short_partial", "r2_partial_tail", "tiny_cuts"])
def test_stream_matches_the_oracle(ctx, case):
    import frender_oracle as O
    rng = random.Random(sum(case.encode()))
    n = 6000
    if case == "r2_short":
        t1, t2, table = make_pair(rng, n, r2_records=n - 1500)
    elif case == "r1_short_partial":
        t2, t1, table = make_pair(rng, n, r2_records=n - 2500, crop_tail=7)   # R1 is the short one and ends mid-record
        # the key comes from R2: give the table R2's keys (same keys, mates swapped in the text only)
    elif case == "r2_partial_tail":
        t1, t2, table = make_pair(rng, n, crop_tail=3)
    else:
        t1, t2, table = make_pair(rng, n)
    roles = O.sink_names(table)
    want = O.route_pairs(t1.splitlines(keepends=True), t2.splitlines(keepends=True), table, roles)
    if case == "tiny_cuts":
        cut1, cut2 = (lambda: rng.randint(1, 3000)), (lambda: rng.randint(1, 3000))
    else:
        cut1, cut2 = (lambda: rng.randint(100_000, 400_000)), (lambda: rng.randint(50_000, 500_000))
    got, pairs = stream(ctx, t1, t2, table, roles, cut1, cut2)
    assert pairs == sum(v[1].count(b"\n@M0:") + (1 if v[1] else 0) for v in want.values())
    for name, (w1, w2) in want.items():
        assert got[name][0] == w1, f"{case}: R1 of sink {name}"
        assert got[name][1] == w2, f"{case}: R2 of sink {name}"


def test_stream_reports_an_unknown_key(ctx):
    import frender_oracle as O
    from frender_b200 import _lib
    from frender_b200.engine import FrbError
    rng = random.Random(5)
    t1, t2, table = make_pair(rng, 500)
    roles = O.sink_names(table)
    lost = t2.splitlines()[4 * 321].rsplit(":", 1)[1]
    del table[lost]
    with pytest.raises(FrbError) as err:
        stream(ctx, t1, t2, table, roles, lambda: 1 << 20, lambda: 1 << 20)
    assert err.value.code == _lib.ERR_KEY_NOT_FOUND and lost in err.value.message
    ctx.route_reset()


def test_route_pair_is_a_stream_of_one(ctx):
    """The synchronous entry point (round 1) still answers like the oracle."""
    import frender_oracle as O
    from frender_b200.engine import pack_keys
    rng = random.Random(11)
    t1, t2, table = make_pair(rng, 3000, crop_tail=5)
    roles = O.sink_names(table)
    want = O.route_pairs(t1.splitlines(keepends=True), t2.splitlines(keepends=True), table, roles)
    names = sorted({n for n in roles.values() if n})
    sid = {n: i for i, n in enumerate(names)}
    role_of = {"index_hop": "#hop", "ambiguous": "#amb", "undetermined": "#und"}
    keys = list(table)
    routes = np.array([sid[roles[table[k][1]] if table[k][0] == "demuxable" else roles[role_of[table[k][0]]]]
                       for k in keys], np.uint32)
    ctx.route_load(pack_keys(keys), routes, len(names))
    o1, o2, off1, off2, pairs, used1, used2 = ctx.route_pair(t1.encode(), t2.encode(), 3)
    assert pairs == 3000 and used1 == len(t1) and used2 == len(t2)
    for n, i in sid.items():
        assert bytes(o1[off1[i]:off1[i + 1]]) == want[n][0]
        assert bytes(o2[off2[i]:off2[i + 1]]) == want[n][1]
