import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import frender_b200._lib as L
from frender_b200 import synth
from frender_b200.engine import Context, C, unpack_keys, FrbError
spec = synth.make_spec("C1", n_samples=16)
ctx = Context(0, table_log2=16)
h = ctx._h
for n in (50, 100, 300, 3000):
    data = synth.generate(spec, 0, n)
    ctx.reset()
    dbuf, dkeys, doffs = C.c_void_p(), C.c_void_p(), C.c_void_p()
    ctx._ck(L.lib.frb_dev_alloc(h, len(data) + 64, C.byref(dbuf)))
    ctx._ck(L.lib.frb_dev_alloc(h, n * 8, C.byref(dkeys)))
    ctx._ck(L.lib.frb_dev_alloc(h, n * 8, C.byref(doffs)))
    arr = np.frombuffer(data, np.uint8)
    init = np.full(n, 0x7777777777777777, np.uint64)
    ctx._ck(L.lib.frb_h2d(h, dbuf, arr.ctypes.data_as(C.c_void_p), len(data)))
    ctx._ck(L.lib.frb_h2d(h, dkeys, init.ctypes.data_as(C.c_void_p), n*8))
    ctx._ck(L.lib.frb_h2d(h, doffs, init.ctypes.data_as(C.c_void_p), n*8))
    ctx._ck(L.lib.frb_scan_begin(h, 0, 0))
    ctx._ck(L.lib.frb_scan_chunk_dev(h, dbuf, len(data), 0, L.RULE_SCAN, dkeys, doffs))
    reads, uniq = C.c_uint64(), C.c_uint64()
    try:
        ctx._ck(L.lib.frb_scan_end(h, C.byref(reads), C.byref(uniq)))
    except FrbError as e:
        print("n", n, "ERR", e)
    keys, offs = np.empty(n, np.uint64), np.empty(n, np.uint64)
    ctx._ck(L.lib.frb_d2h(h, keys.ctypes.data_as(C.c_void_p), dkeys, n * 8))
    ctx._ck(L.lib.frb_d2h(h, offs.ctypes.data_as(C.c_void_p), doffs, n * 8))
    want = synth.keys_of(spec, 0, n)
    starts = np.concatenate([[0], np.flatnonzero(arr == 10)[3::4][:-1] + 1])
    unset = keys == 0x7777777777777777
    got = unpack_keys(np.where(unset, 0, keys))
    badk = [i for i in range(n) if unset[i] or got[i] != want[i]]
    bado = [i for i in range(n) if offs[i] != starts[i]]
    print("n", n, "bytes", len(data), "reads", reads.value, "bad keys", len(badk), badk[:20], "bad offs", len(bado), bado[:20])
    for i in badk[:5]:
        print("   read", i, "start", starts[i], "tile", starts[i]//32768, "thread", (starts[i]%32768)//128, "unset", bool(unset[i]), "got", got[i], "want", want[i], "off", offs[i] if offs[i]!=0x7777777777777777 else None)
