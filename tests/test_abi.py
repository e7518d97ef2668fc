"""The C-ABI library loads and exports every symbol include/frender_b200.h declares; the ctypes
table matches the header; without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "frender_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(frb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    from frender_b200 import _lib
    names = declared_functions()
    assert len(names) >= 35
    for name in names:
        assert hasattr(_lib.lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_header_arity_matches_ctypes_table():
    from frender_b200 import _lib
    text = open(os.path.join(ROOT, "include", "frender_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, args in re.findall(r"\b(frb_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = args.strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        assert n == len(_lib.SIGNATURES[name][1]), name


def test_key_packing_roundtrip_host_side():
    from frender_b200 import _lib
    out = ctypes.c_uint64()
    assert _lib.lib.frb_pack_key(b"ACGTN+TTGCA", 11, 0, ctypes.byref(out)) == 0
    buf = ctypes.create_string_buffer(24)
    assert _lib.lib.frb_unpack_key(out.value, buf) == 11 and buf.value == b"ACGTN+TTGCA"
    assert _lib.lib.frb_pack_key(b"acgt", 4, 0, ctypes.byref(out)) == _lib.ERR_BAD_ALPHABET
    assert _lib.lib.frb_pack_key(b"acgt", 4, 1, ctypes.byref(out)) == 0          # sheet mode folds case
    assert _lib.lib.frb_pack_key(b"A" * 22, 22, 0, ctypes.byref(out)) == _lib.ERR_KEY_TOO_LONG
    from frender_b200.engine import pack_keys, unpack_keys
    assert unpack_keys(pack_keys(["ACGTN+TTGCA", "", "NNNNNNNNNN+ACGTACGTAC"])) == ["ACGTN+TTGCA", "", "NNNNNNNNNN+ACGTACGTAC"]
    assert int(pack_keys(["ACGTN+TTGCA"])[0]) == sum(
        c << (3 * i) for i, c in enumerate([1, 2, 3, 4, 5, 6, 4, 4, 3, 2, 1]))


def test_no_gpu_means_loud_failure():
    from frender_b200 import _lib
    n = ctypes.c_int()
    _lib.lib.frb_device_count(ctypes.byref(n))
    if n.value > 0:
        pytest.skip("a GPU is present")
    from frender_b200.engine import Context, FrbError
    with pytest.raises(FrbError, match="no CPU fallback"):
        Context(0)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under frender_b200/ or frender.py may reference it."""
    offenders = []
    for base, _, files in os.walk(os.path.join(ROOT, "frender_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".cpp", ".h")):
                text = open(os.path.join(base, f), errors="replace").read()
                if re.search(r"frender_oracle|c_oracle|liboracle|/oracle/", text):
                    offenders.append(os.path.join(base, f))
    text = open(os.path.join(ROOT, "frender.py")).read()
    if "oracle" in text:
        offenders.append("frender.py")
    assert not offenders, offenders


def test_native_scan_csv_writer_matches_csv_module(tmp_path):
    """frb_write_scan_csv is host-only: its bytes must equal csv.writer's (default dialect, F:499) for keys
    with zero, one and two '+' parts, empty matches, 2^40 counts and sheet strings that need quoting."""
    import csv
    import ctypes as C
    import io

    import numpy as np
    from frender_b200 import _lib
    from frender_b200.engine import pack_keys, unpack_keys
    keys = np.array(pack_keys(["ACGT+TTTT", "AAAA+CCCC+GG", "NNNN", "", "ACGTACGTAC+ACGTACGTAC"]), np.uint64)
    counts = np.array([1, 22, 333, 4444, 2 ** 40], np.uint64)
    m1, m2 = np.array([0, -1, 2, 1, 0], np.int32), np.array([1, -1, 0, 2, 0], np.int32)
    kind, srow = np.array([2, 0, 3, 1, 2], np.uint8), np.array([0, -1, -1, -1, 2], np.int32)
    ok = np.array([1, 0, 1, 1, 0], np.uint8)
    idx1, idx2, ids = ["ACGT", "GG,TT", 'A"B'], ["TTTT", "CCCC", "x\ny"], ["S 1", 'we"ird,name', "plain"]

    def strings(items):
        arr = (C.c_char_p * len(items))()
        for i, item in enumerate(items):
            arr[i] = item.encode()
        return arr

    path = tmp_path / "out.csv"
    rc = _lib.lib.frb_write_scan_csv(str(path).encode(),
                                     *[a.ctypes.data_as(C.c_void_p) for a in (keys, counts, m1, m2, kind, srow, ok)],
                                     len(keys), strings(idx1), strings(idx2), strings(ids), 3, 0)
    assert rc == 0
    buf = io.StringIO(newline="")
    w = csv.writer(buf)
    w.writerow(["idx1", "idx2", "matched_idx1", "matched_idx2", "read_type", "sample_name", "reads", "demux_ok"])
    kinds = ("undetermined", "index_hop", "demuxable", "ambiguous")
    for k, c, a, b, t, s, o in zip(unpack_keys(keys), counts.tolist(), m1.tolist(), m2.tolist(), kind.tolist(),
                                   srow.tolist(), ok.tolist()):
        parts = k.split("+")
        w.writerow([parts[0], parts[1] if len(parts) > 1 else "", idx1[a] if a >= 0 else "",
                    idx2[b] if b >= 0 else "", kinds[t], ids[s] if s >= 0 else "", c, bool(o)])
    assert path.read_bytes() == buf.getvalue().encode()
    assert _lib.lib.frb_write_scan_csv(b"/nonexistent-dir/x.csv", None, None, None, None, None, None, None, 0,
                                       strings(idx1), strings(idx2), strings(ids), 3, 0) == _lib.ERR_IO
