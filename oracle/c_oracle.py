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
