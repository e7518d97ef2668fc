"""At-scale checks: the C oracle on millions of reads, and size-independent properties on a lane
slice far larger than the Python oracle can follow."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def device_lane(ctx, spec, reads, chunk=4_000_000):
    """Generate `reads` reads on the device into one resident buffer; returns (dptr, nbytes)."""
    import frender_b200._lib as L
    from frender_b200.engine import C
    h, lib = ctx._h, L.lib
    pk = lambda rows: np.array([sum(int(c) << (2 * p) for p, c in enumerate(r)) for r in rows], np.uint32)
    i7, i5, cdf = pk(spec.sheet_i7), pk(spec.emit_i5()), np.ascontiguousarray(spec.cdf, np.uint64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    ctx._ck(lib.frb_synth_load(h, spec.seed, spec.l1, spec.l2, spec.n_samples, vp(i7), vp(i5), vp(cdf), spec.lane,
                               spec.read_len, spec.sub_t, spec.n_t, spec.rand_t, spec.hop_t))
    cap = reads * 376 + (1 << 20)
    dbuf = C.c_void_p()
    ctx._ck(lib.frb_dev_alloc(h, cap, C.byref(dbuf)))
    off = 0
    for g in range(0, reads, chunk):
        n = C.c_uint64()
        ctx._ck(lib.frb_synth_generate(h, g, min(g + chunk, reads), 1, C.c_void_p(dbuf.value + off), cap - off,
                                       C.byref(n)))
        off += n.value
    return dbuf, off


def scan_resident(ctx, dbuf, nbytes):
    import frender_b200._lib as L
    from frender_b200.engine import C
    ctx.reset()
    ctx._ck(L.lib.frb_scan_begin(ctx._h, 0, 0))
    ctx._ck(L.lib.frb_scan_chunk_dev(ctx._h, dbuf, nbytes, 0, L.RULE_SCAN, None, None))
    r, u = C.c_uint64(), C.c_uint64()
    ctx._ck(L.lib.frb_scan_end(ctx._h, C.byref(r), C.byref(u)))
    return r.value, u.value


def test_c_oracle_at_3m_reads():
    """3 M reads of the C2 lane, the bench shape itself (384-row sheet, -n 1 -rc): counts and first-appearance
    order against the C oracle; BOTH matcher passes of every unique key -- the forward + reverse-complement
    first pass with its per-sample orientation sums (F:294-388) and the oriented second pass (F:618-630) --
    against the C oracle (itself pinned on the reference's c2_384 golden case)."""
    import c_oracle
    import frender_b200._lib as L
    from frender_b200 import synth
    from frender_b200.engine import C, Context, reverse_complement, unpack_keys
    reads = 3_000_000
    spec = synth.make_spec("C2")
    ctx = Context(0, table_log2=21)
    dbuf, nbytes = device_lane(ctx, spec, reads)
    got_reads, got_uniq = scan_resident(ctx, dbuf, nbytes)
    host = np.empty(nbytes, np.uint8)
    ctx._ck(L.lib.frb_d2h(ctx._h, host.ctypes.data_as(C.c_void_p), dbuf, nbytes))
    ctx._ck(L.lib.frb_dev_free(ctx._h, dbuf))
    want, want_reads = c_oracle.tally(host.tobytes())
    keys, counts, first = ctx.total_arrays()
    names = unpack_keys(keys)
    assert got_reads == want_reads == reads and got_uniq == len(want)
    assert names == list(want) and counts.tolist() == list(want.values())
    assert (np.diff(first.astype(np.int64)) > 0).all()
    idx = spec.indexes()
    assert len(idx["id"]) == 384
    sheet = ctx.load_sheet(idx)
    # pass 1: forward and reverse-complement i5 together
    res = ctx.match(1, True)
    ref = np.array(c_oracle.classify_all_rc(names, idx, 1), np.int32)
    for col, name in enumerate(("m1", "m2", "type", "srow", "m2rc", "type_rc", "srow_rc")):
        got = res[name].astype(np.int32)
        if name == "m1":   # the rc pass's idx1 row fills an empty forward one (F:319-323); same row either way
            pass
        assert (got == ref[:, col]).all(), name
    calls = c_oracle.rc_calls(names, counts.tolist(), ref.tolist(), idx)
    got_calls = ctx.rc_calls(res)
    assert {k: (v["call"], v["reads_f"], v["reads_rc"]) for k, v in got_calls.items()} == calls
    assert any(v[0] for v in calls.values()) and not all(v[0] for v in calls.values())
    # pass 2: every row's idx2 in the orientation its sample was called in
    use = np.array([calls[name][0] for name in idx["id"]], np.uint8)
    oriented = dict(idx, idx2=[reverse_complement(s) if u else s for s, u in zip(idx["idx2"], use)])
    res2 = ctx.match(1, False, use)
    ref2 = np.array(c_oracle.classify_all(names, oriented, 1), np.int32)
    assert (res2["m1"] == ref2[:, 0]).all() and (res2["m2"] == ref2[:, 1]).all()
    assert (res2["type"] == ref2[:, 2]).all() and (res2["srow"] == ref2[:, 3]).all()
    ctx.close()


def test_properties_at_40m_reads():
    """15 GB lane slice: conservation (sum of counts = reads), strict first-appearance order,
    idempotence, and invariance under host-side chunked feeding of a prefix."""
    import frender_b200._lib as L
    from frender_b200 import synth
    from frender_b200.engine import C, Context
    reads = 40_000_000
    spec = synth.make_spec("C2")
    ctx = Context(0, table_log2=23)
    dbuf, nbytes = device_lane(ctx, spec, reads)
    r1, u1 = scan_resident(ctx, dbuf, nbytes)
    k1, c1, f1 = ctx.total_arrays()
    assert r1 == reads and int(c1.sum()) == reads and len(k1) == u1
    assert (np.diff(f1.astype(np.int64)) > 0).all() and f1[0] == 0
    assert len(np.unique(k1)) == len(k1)
    r2, u2 = scan_resident(ctx, dbuf, nbytes)                        # idempotent
    k2, c2, f2 = ctx.total_arrays()
    assert (k1 == k2).all() and (c1 == c2).all() and (f1 == f2).all()
    # the first 2 GiB fed from the host in ragged chunks == the same bytes scanned resident
    part = 2 << 30
    host = np.empty(part, np.uint8)
    ctx._ck(L.lib.frb_d2h(ctx._h, host.ctypes.data_as(C.c_void_p), dbuf, part))
    part = int(np.flatnonzero(host[-4096:] == 10)[-1]) + part - 4096 + 1     # cut at a line end
    ra, _ = scan_resident(ctx, dbuf, part)
    ka, ca, fa = ctx.total_arrays()
    ctx.reset()
    rb, _ = ctx.scan_bytes(memoryview(host)[:part], chunk=97_000_001)
    kb, cb, fb = ctx.total_arrays()
    assert ra == rb and (ka == kb).all() and (ca == cb).all() and (fa == fb).all()
    ctx._ck(L.lib.frb_dev_free(ctx._h, dbuf))
    ctx.close()
