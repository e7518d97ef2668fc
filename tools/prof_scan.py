"""Small resident-input scan used under ncu (python tools/prof_scan.py [reads] [steps])."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frender_b200 import _lib as L  # noqa: E402
from frender_b200 import synth  # noqa: E402
from frender_b200.engine import C, Context, PackedSheet  # noqa: E402

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
spec = synth.make_spec("C2")
ctx = Context(0, table_log2=int(sys.argv[3]) if len(sys.argv) > 3 else 22)
h, lib, ck = ctx._h, L.lib, ctx._ck
pk = lambda rows: np.array([sum(int(c) << (2 * p) for p, c in enumerate(r)) for r in rows], np.uint32)
i7, i5, cdf = pk(spec.sheet_i7), pk(spec.emit_i5()), np.ascontiguousarray(spec.cdf, np.uint64)
vp = lambda a: a.ctypes.data_as(C.c_void_p)
ck(lib.frb_synth_load(h, spec.seed, spec.l1, spec.l2, spec.n_samples, vp(i7), vp(i5), vp(cdf), spec.lane,
                      spec.read_len, spec.sub_t, spec.n_t, spec.rand_t, spec.hop_t))
cap = reads * 375 + (1 << 20)
dbuf, nb = C.c_void_p(), C.c_uint64()
ck(lib.frb_dev_alloc(h, cap, C.byref(dbuf)))
ck(lib.frb_synth_generate(h, 0, reads, 1, dbuf, cap, C.byref(nb)))
sheet = ctx.load_sheet(PackedSheet(spec.indexes()))
ctx.prof(True)
times = []
for _ in range(steps):
    ck(lib.frb_reset(h))
    ck(lib.frb_scan_begin(h, 0, 0))
    ck(lib.frb_scan_chunk_dev(h, dbuf, nb.value, 0, L.RULE_SCAN, None, None))
    r, u = C.c_uint64(), C.c_uint64()
    ck(lib.frb_scan_end(h, C.byref(r), C.byref(u)))
    m = ctx.match(1, True, None, want_outputs=False)
    use = np.array([m["f_sum"][g] < m["rc_sum"][g] for g in sheet.group], np.uint8)
    ctx.match(1, False, use, want_outputs=False)
    times.append(ctx.prof_read(L.K_SCAN)[0])
best = min(times[1:] or times)
print("reads", r.value, "unique", u.value, "bytes", nb.value, "scan_ms", round(best, 4), "GB/s", round(nb.value / best / 1e6, 1))
