"""Device generator == host generator, byte for byte (so host slices can stand in for device lanes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("config,read_no,g0,n", [("C1", 1, 0, 3000), ("C2", 1, 12345, 5000), ("C2", 2, 7, 2000),
                                                 ("C3", 1, 10**9, 2500), ("C5", 1, 0, 4000)])
def test_device_generator_matches_host(config, read_no, g0, n):
    import frender_b200._lib as L
    from frender_b200 import synth
    from frender_b200.engine import C, Context
    spec = synth.make_spec(config, lane=3)
    want = synth.generate(spec, g0, g0 + n, read_no)
    ctx = Context(0, table_log2=12)
    h, lib = ctx._h, L.lib
    pk = lambda rows: np.array([sum(int(c) << (2 * p) for p, c in enumerate(r)) for r in rows], np.uint32)
    i7, i5, cdf = pk(spec.sheet_i7), pk(spec.emit_i5()), np.ascontiguousarray(spec.cdf, np.uint64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    ctx._ck(lib.frb_synth_load(h, spec.seed, spec.l1, spec.l2, spec.n_samples, vp(i7), vp(i5), vp(cdf), spec.lane,
                               spec.read_len, spec.sub_t, spec.n_t, spec.rand_t, spec.hop_t))
    dbuf, nb = C.c_void_p(), C.c_uint64()
    cap = len(want) + 4096
    ctx._ck(lib.frb_dev_alloc(h, cap, C.byref(dbuf)))
    ctx._ck(lib.frb_synth_generate(h, g0, g0 + n, read_no, dbuf, cap, C.byref(nb)))
    got = np.empty(nb.value, np.uint8)
    ctx._ck(lib.frb_d2h(h, vp(got), dbuf, nb.value))
    ctx._ck(lib.frb_dev_free(h, dbuf))
    ctx.close()
    assert nb.value == len(want)
    assert got.tobytes() == want
