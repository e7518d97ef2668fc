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
