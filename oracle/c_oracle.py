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
    lib.oracle_classify_rc.restype = C.c_int
    lib.oracle_classify_rc.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_int), C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.oracle_revcomp_sheet.restype = None
    lib.oracle_revcomp_sheet.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p]
    lib.oracle_route.restype = C.c_int
    lib.oracle_route.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_char_p, C.c_void_p, C.c_void_p,
                                 C.c_uint64, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint64)]
    lib.oracle_fnv1a_segments.restype = None
    lib.oracle_fnv1a_segments.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    return lib


FNV_BASIS = 14695981039346656037


def route_sums(r1, n1, r2, n2, keys, sinks, n_sinks):
    """The demux loop (F:774-810) over R1/R2 at host addresses r1 / r2 (ints, or bytes objects): per sink the byte
    counts, record count and order-sensitive FNV-1a digests of both mates, as a structured numpy array with the
    fields bytes1, bytes2, hash1, hash2, records.  keys: list of str, sinks: sink id per key.  SystemExit like the
    reference for a key that is not in the table."""
    import numpy as np
    lib = load()
    blob = "".join(keys).encode()
    off = np.zeros(len(keys) + 1, np.uint32)
    np.cumsum([len(k) for k in keys], out=off[1:])
    sk = np.ascontiguousarray(sinks, np.uint32)
    out = np.zeros(n_sinks, np.dtype([("bytes1", "u8"), ("bytes2", "u8"), ("hash1", "u8"), ("hash2", "u8"), ("records", "u8")]))
    keep = []

    def addr(x, n):
        if isinstance(x, int):
            return C.c_void_p(x), n
        buf = np.frombuffer(x, np.uint8)
        keep.append(buf)
        return C.c_void_p(buf.ctypes.data), buf.size

    a1, n1 = addr(r1, n1)
    a2, n2 = addr(r2, n2)
    bad = C.c_uint64()
    rc = lib.oracle_route(a1, n1, a2, n2, blob, off.ctypes.data_as(C.c_void_p), sk.ctypes.data_as(C.c_void_p), len(keys),
                          n_sinks, out.ctypes.data_as(C.c_void_p), C.byref(bad))
    if rc == -1:
        raise SystemExit(f"Couldn't find barcode of record {bad.value} in supplied frender result file!")
    assert rc == 0
    return out


def fnv1a_segments(hashes, base_addr, off):
    """hashes[s] continued over the bytes base[off[s]:off[s + 1]] for every sink (hashes: uint64 array, updated)."""
    import numpy as np
    lib = load()
    off = np.ascontiguousarray(off, np.uint64)
    lib.oracle_fnv1a_segments(hashes.ctypes.data_as(C.c_void_p), C.c_void_p(base_addr), off.ctypes.data_as(C.c_void_p),
                              len(off) - 1)


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


def classify_all_rc(keys, indexes, max_subs):
    """First pass of `-rc` (F:294-351) for every key:
    [(m1_row, m2_row, type, sample_row, m2rc_row, rc_type, rc_sample_row)]."""
    lib = load()
    rows = len(indexes["id"])
    l1, l2 = len(indexes["idx1"][0]), len(indexes["idx2"][0])
    a = "".join(indexes["idx1"]).encode()
    b = "".join(indexes["idx2"]).encode()
    brc = C.create_string_buffer(len(b) + 1)
    lib.oracle_revcomp_sheet(b, rows, l2, brc)
    first = {}
    group = (C.c_int * rows)(*[first.setdefault(name, len(first)) for name in indexes["id"]])
    out = (C.c_int * 7)()
    res = []
    for k in keys:
        if lib.oracle_classify_rc(k.encode(), a, b, brc.raw[:len(b)], group, rows, l1, l2, max_subs, out):
            raise AssertionError(f"Barcode {k} doesn't match length of supplied barcode")
        res.append(tuple(out))
    return res


def rc_calls(keys, counts, first_pass, indexes):
    """call_rc_mode_per_id (F:354-388) from classify_all_rc's rows: {name: (use_rc, reads_f, reads_rc)}."""
    acc = {name: [0, 0] for name in indexes["id"]}
    for n, rec in zip(counts, first_pass):
        if rec[3] >= 0:
            acc[indexes["id"][rec[3]]][0] += int(n)
        if rec[6] >= 0:
            acc[indexes["id"][rec[6]]][1] += int(n)
    return {name: (f < r, f, r) for name, (f, r) in acc.items()}
