"""Demux router (hot path C) as a stream: python tools/bench_demux.py [pairs] [chunk_mb] [steps]

R1/R2 of a C4-shaped lane are device-generated (same read ordinals, mate 1 / mate 2) and pulled to pinned host
memory; the results table comes from a scan of R1 and both matcher passes.  The two mates are then pushed through
frb_route_push / frb_route_pop in chunks of chunk_mb, cut at fixed byte counts (not at record ends), two chunks in
flight -- exactly what the CLI's demux loop does between its inflate and deflate threads.  Reported:
  * pairs/s host -> host: pinned host buffers in, pinned host buffers out (H2D + parser + router + D2H overlapped);
  * kernel-only time from the library's per-class events (parser = K_SCAN, router = K_ROUTE) and the GB/s that is
    on the algorithmic bytes 2 x (R1 + R2);
  * checks: every pair routed once, per-sink record counts equal the classification's read counts, sink bytes add up.
`measure()` is what bench.py's demux leg calls."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def measure(device=0, pairs=4_000_000, chunk_mb=64, steps=3, warmup=1, ctx=None, oracle_check=True):
    from frender_b200 import _lib as L
    from frender_b200 import synth
    from frender_b200.engine import C, Context
    own = ctx is None
    if own:
        ctx = Context(device, table_log2=22)
    h, lib, ck = ctx._h, L.lib, ctx._ck
    spec = synth.make_spec("C4")
    pk = lambda rows: np.array([sum(int(c) << (2 * p) for p, c in enumerate(r)) for r in rows], np.uint32)
    i7, i5, cdf = pk(spec.sheet_i7), pk(spec.emit_i5()), np.ascontiguousarray(spec.cdf, np.uint64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    ck(lib.frb_synth_load(h, spec.seed, spec.l1, spec.l2, spec.n_samples, vp(i7), vp(i5), vp(cdf), spec.lane,
                          spec.read_len, spec.sub_t, spec.n_t, spec.rand_t, spec.hop_t))
    cap = pairs * 376 + (1 << 20)
    dbuf = C.c_void_p()
    ck(lib.frb_dev_alloc(h, cap, C.byref(dbuf)))
    mates = []
    try:
        for mate in (1, 2):
            nb = C.c_uint64()
            ck(lib.frb_synth_generate(h, 0, pairs, mate, dbuf, cap, C.byref(nb)))
            host = C.c_void_p()
            ck(lib.frb_host_alloc(C.byref(host), nb.value))
            ck(lib.frb_d2h(h, host, dbuf, nb.value))
            mates.append((host, nb.value))
        # results table: scan R1 (resident), classify, sink = sample row or one of 3 extra sinks
        ck(lib.frb_synth_generate(h, 0, pairs, 1, dbuf, cap, C.byref(C.c_uint64())))
        ctx.reset()
        ck(lib.frb_scan_begin(h, 0, 0))
        ck(lib.frb_scan_chunk_dev(h, dbuf, mates[0][1], 0, L.RULE_SCAN, None, None))
        r, u = C.c_uint64(), C.c_uint64()
        ck(lib.frb_scan_end(h, C.byref(r), C.byref(u)))
    finally:
        ck(lib.frb_dev_free(h, dbuf))
    keys, counts, _ = ctx.total_arrays()
    sheet = ctx.load_sheet(spec.indexes())
    first = ctx.match(1, True, None, want_outputs=False)
    use = np.array([first["f_sum"][g] < first["rc_sum"][g] for g in sheet.group], np.uint8)
    res = ctx.match(1, False, use)
    S = spec.n_samples
    sink = np.where(res["type"] == 2, res["srow"], S + np.where(res["type"] == 1, 0, np.where(res["type"] == 3, 1, 2)))
    n_sinks = S + 3
    ctx.route_load(keys, sink.astype(np.uint32), n_sinks)
    want_per_sink = np.bincount(sink, weights=counts.astype(np.float64), minlength=n_sinks).astype(np.int64)

    chunk = chunk_mb << 20
    o1, o2 = C.c_void_p(), C.c_void_p()
    off1, off2 = np.zeros(n_sinks + 1, np.uint64), np.zeros(n_sinks + 1, np.uint64)
    npairs, c1, c2, bad = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    (h1, n1), (h2, n2) = mates
    n_chunks = max((n1 + chunk - 1) // chunk, (n2 + chunk - 1) // chunk)

    digest = None
    if oracle_check:   # the checker (oracle/, test infrastructure): the reference's loop restated in C, per-sink digests
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
        import c_oracle
        from frender_b200.engine import unpack_keys
        digest = [np.full(n_sinks, c_oracle.FNV_BASIS, np.uint64), np.full(n_sinks, c_oracle.FNV_BASIS, np.uint64)]

    def stream(check):
        """One pass of both mates through the router; check: count the records of every sink on the way (and, with
        the oracle, continue every sink's order-sensitive digest over the bytes it receives)."""
        ck(lib.frb_route_reset(h))
        total, in_flight = 0, 0
        per_sink = np.zeros(n_sinks, np.int64)
        bytes1 = bytes2 = 0

        def pop():
            nonlocal total, bytes1, bytes2
            ck(lib.frb_route_pop(h, C.byref(o1), C.byref(o2), vp(off1), vp(off2), C.byref(npairs), C.byref(c1), C.byref(c2),
                                 C.byref(bad)))
            total += npairs.value
            bytes1 += int(off1[-1])
            bytes2 += int(off2[-1])
            if check and off1[-1]:
                out = np.ctypeslib.as_array((C.c_uint8 * int(off1[-1])).from_address(o1.value))
                nl = np.flatnonzero(out == 10)
                per_sink[:] += np.diff(np.searchsorted(nl, off1.astype(np.int64))) // 4
            if check and digest is not None:
                c_oracle.fnv1a_segments(digest[0], o1.value, off1)
                c_oracle.fnv1a_segments(digest[1], o2.value, off2)

        for k in range(n_chunks):
            a0, b0 = min(k * chunk, n1), min(k * chunk, n2)
            a1, b1 = min(a0 + chunk, n1), min(b0 + chunk, n2)
            final = (1 if a1 >= n1 else 0) | (2 if b1 >= n2 else 0)
            ck(lib.frb_route_push(h, C.c_void_p(h1.value + a0), a1 - a0, C.c_void_p(h2.value + b0), b1 - b0, final))
            in_flight += 1
            if in_flight == 2:
                pop()
                in_flight -= 1
        while in_flight:
            pop()
            in_flight -= 1
        return total, per_sink, bytes1, bytes2

    total, per_sink, bytes1, bytes2 = stream(True)
    assert total == pairs, (total, pairs)
    assert bytes1 == n1 and bytes2 == n2, "sink bytes do not add up to the input"
    assert (per_sink == want_per_sink).all(), "per-sink record counts differ from the classification"
    checked = "every pair routed once; per-sink record counts == classification's read counts; sink bytes == input bytes"
    if digest is not None:
        want = c_oracle.route_sums(h1.value, n1, h2.value, n2, unpack_keys(keys), sink.astype(np.uint32), n_sinks)
        assert (want["records"].astype(np.int64) == per_sink).all(), "per-sink records differ from the C oracle's demux loop"
        assert (want["hash1"] == digest[0]).all() and (want["hash2"] == digest[1]).all(), \
            "a sink's byte stream differs from the C oracle's demux loop"
        checked += "; every sink's R1 and R2 byte stream (order-sensitive FNV-1a digest) == the C oracle's restatement of F:774-810"
    for _ in range(warmup):
        stream(False)
    ctx.prof(True)
    for k in range(L.K_NUM):
        ctx.prof_read(k)
    launches0 = ctx.launches() if hasattr(ctx, "launches") else 0
    t0 = time.perf_counter()
    for _ in range(steps):
        stream(False)
    dt = (time.perf_counter() - t0) / steps
    launches = (ctx.launches() - launches0) // steps if hasattr(ctx, "launches") else None
    scan_ms, _ = ctx.prof_read(L.K_SCAN)
    route_ms, _ = ctx.prof_read(L.K_ROUTE)
    ctx.prof(False)
    ck(lib.frb_route_reset(h))
    for host, _ in mates:
        lib.frb_host_free(host)
    kernel_ms = (scan_ms + route_ms) / steps
    moved = 2 * (n1 + n2)
    out = {"pairs": pairs, "chunk_mb": chunk_mb, "chunks": int(n_chunks), "sinks": int(n_sinks),
           "pairs_per_s_host_to_host": pairs / dt, "ms_per_pass_host_to_host": dt * 1e3,
           "host_to_host_gbs_in_plus_out": moved / dt / 1e9,
           "pairs_per_s_kernel_only": pairs / (kernel_ms / 1e3), "kernel_ms": kernel_ms,
           "parser_ms": scan_ms / steps, "router_ms": route_ms / steps,
           "algorithmic_bytes": int(moved), "kernel_gbs": moved / (kernel_ms / 1e3) / 1e9,
           "gpu_launches_per_pass": launches,
           "checked": checked}
    if own:
        ctx.close()
    return out


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    mb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    st = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    print(json.dumps(measure(0, n, mb, st)))
