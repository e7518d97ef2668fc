"""Demux router (hot path C) on a C4-shaped chunk pair: python tools/bench_demux.py [pairs]

R1/R2 are device-generated (same read ordinals, mate 1 / mate 2), pulled to pinned host memory, the
results table comes from a scan of R1 + both matcher passes.  Reports pairs/s through frb_route_pair
(host buffers in, host buffers out: H2D + kernels + D2H) and the kernel-only time from the library's
per-class events, and checks conservation (every byte of every record lands in exactly one sink,
per-sink record counts equal the classification's read counts)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frender_b200 import _lib as L  # noqa: E402
from frender_b200 import synth  # noqa: E402
from frender_b200.engine import C, Context  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
spec = synth.make_spec("C4")
ctx = Context(0, table_log2=22)
h, lib, ck = ctx._h, L.lib, ctx._ck
pk = lambda rows: np.array([sum(int(c) << (2 * p) for p, c in enumerate(r)) for r in rows], np.uint32)
i7, i5, cdf = pk(spec.sheet_i7), pk(spec.emit_i5()), np.ascontiguousarray(spec.cdf, np.uint64)
vp = lambda a: a.ctypes.data_as(C.c_void_p)
ck(lib.frb_synth_load(h, spec.seed, spec.l1, spec.l2, spec.n_samples, vp(i7), vp(i5), vp(cdf), spec.lane,
                      spec.read_len, spec.sub_t, spec.n_t, spec.rand_t, spec.hop_t))
cap = pairs * 376 + (1 << 20)
dbuf = C.c_void_p()
ck(lib.frb_dev_alloc(h, cap, C.byref(dbuf)))
mates = []
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

o1, o2 = np.empty(mates[0][1], np.uint8), np.empty(mates[1][1], np.uint8)
off1, off2 = np.zeros(n_sinks + 1, np.uint64), np.zeros(n_sinks + 1, np.uint64)
npairs, u1, u2, bad = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()


def step():
    ck(lib.frb_route_pair(h, mates[0][0], mates[0][1], mates[1][0], mates[1][1], 3, vp(o1), vp(o2), vp(off1),
                          vp(off2), C.byref(npairs), C.byref(u1), C.byref(u2), C.byref(bad)))


for _ in range(3):
    step()
ctx.prof(True)
for k in range(5):
    ctx.prof_read(k)
t0 = time.perf_counter()
steps = 5
for _ in range(steps):
    step()
dt = (time.perf_counter() - t0) / steps
scan_ms, _ = ctx.prof_read(L.K_SCAN)
route_ms, _ = ctx.prof_read(L.K_ROUTE)
assert npairs.value == pairs and u1.value == mates[0][1] and u2.value == mates[1][1]
assert off1[-1] == mates[0][1] and off2[-1] == mates[1][1]
# per-sink record counts: count newlines / 4 in each sink region of mate 1
nl = np.flatnonzero(o1 == 10)
per_sink = np.diff(np.searchsorted(nl, off1.astype(np.int64))) // 4
assert (per_sink == want_per_sink).all(), "per-sink record counts differ from the classification"
kernel_ms = (scan_ms + route_ms) / steps
bytes_moved = 2 * (mates[0][1] + mates[1][1])
print(json.dumps({"pairs": pairs, "pairs_per_s_e2e_host_buffers": pairs / dt, "ms_per_chunk_e2e": dt * 1e3,
                  "pairs_per_s_kernel_only": pairs / (kernel_ms / 1e3), "kernel_ms": kernel_ms,
                  "scan_ms": scan_ms / steps, "route_ms": route_ms / steps,
                  "algorithmic_bytes": bytes_moved, "kernel_gbs": bytes_moved / (kernel_ms / 1e3) / 1e9,
                  "h2d_plus_d2h_gbs": bytes_moved / dt / 1e9, "sinks": n_sinks, "checks": "conservation ok"}))
