"""ctypes view of oracle/_build/liboracle.so (TEST INFRASTRUCTURE ONLY, see frender_oracle.c)."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "liboracle.so")


def load():
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(os.path.join(HERE, "frender_oracle.c")):
        subprocess.run(["make", "-s", "-C", HERE], check=True)
    lib = C.CDLL(SO)
    lib.oracle_tally.restype = C.c_void_p
    lib.oracle_tally.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.POINTER(C.c_int)]
    for name, res in (("oracle_tally_size", C.c_uint64), ("oracle_tally_reads", C.c_uint64)):
        getattr(lib, name).restype = res
        getattr(lib, name).argtypes = [C.c_void_p]
    lib.oracle_tally_key.restype = C.c_char_p
    lib.oracle_tally_key.argtypes = [C.c_void_p, C.c_uint64]
    lib.oracle_tally_count.restype = C.c_uint64
    lib.oracle_tally_count.argtypes = [C.c_void_p, C.c_uint64]
    lib.oracle_tally_free.argtypes = [C.c_void_p]
    lib.oracle_classify.restype = C.c_int
    lib.oracle_classify.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.POINTER(C.c_int)]
    return lib


def tally(data, rule=0, sample=0):
    """(ordered {key: count}, reads) of decompressed FASTQ bytes; IndexError like the reference."""
    lib = load()
    buf = (C.c_char * len(data)).from_buffer_copy(data) if len(data) else None
    err = C.c_int()
    h = lib.oracle_tally(buf, len(data), rule, sample, C.byref(err))
    try:
        if err.value == -1:
            raise IndexError("list index out of range")
        if err.value:
            raise ValueError("key too long for the C oracle")
        n = lib.oracle_tally_size(h)
        out = {lib.oracle_tally_key(h, i).decode("latin1"): lib.oracle_tally_count(h, i) for i in range(n)}
        return out, lib.oracle_tally_reads(h)
    finally:
        lib.oracle_tally_free(h)


def classify_all(keys, indexes, max_subs):
    """[(m1_row, m2_row, type, sample_row)] for every key (forward sheet only)."""
    lib = load()
    l1, l2 = len(indexes["idx1"][0]), len(indexes["idx2"][0])
    a = "".join(indexes["idx1"]).encode()
    b = "".join(indexes["idx2"]).encode()
    out = (C.c_int * 4)()
    res = []
    for k in keys:
        if lib.oracle_classify(k.encode(), a, b, len(indexes["id"]), l1, l2, max_subs, out):
            raise AssertionError(f"Barcode {k} doesn't match length of supplied barcode")
        res.append(tuple(out))
    return res
