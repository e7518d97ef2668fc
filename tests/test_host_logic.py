"""Host-side logic of the CLI that needs no GPU: text-mode newline translation of the demux reader (F:776), the
class x file matrix of demux_ok (F:521-550) against the reference's own loop restated by the oracle, grouping of small
files into runs, the -c stream policy, table sizing."""
import gzip
import os
import re

import numpy as np
import pytest


def test_text_chunks_translate_newlines_like_text_mode(tmp_path):
    """Any read size gives the bytes gzip.open(..., "rt") gives (universal newlines): "\\r\\n" and a lone "\\r" become
    "\\n", also when a chunk ends between the two."""
    from frender_b200.cli import TextChunks
    raw = b"@a 1:N:0:AC+GT\r\nACGT\r+\r\nFFFF\n@b\r\r\nAC\n\r"
    p = tmp_path / "x.fastq.gz"
    p.write_bytes(gzip.compress(raw))
    want = gzip.open(p, "rt").read().encode()
    for size in (1, 2, 3, 5, 7, 64, 1 << 20):
        t = TextChunks(p)
        got = b""
        while not t.eof:
            piece = t.read(size)
            assert len(piece) <= max(size, 2)
            got += piece
        assert got == want, size


def test_class_file_matrix_matches_the_references_regexes():
    """Row = read type 0/1/3 or 4 + sample row, column = file: what re.search says in F:521-550 (sample name as a
    pattern, prefix removed, case-insensitive); a name that is not a valid pattern is marked, not raised."""
    from frender_b200.cli import class_file_matrix
    files = ["S1_S1_L001_R1_001.fastq.gz", "Undetermined_S0_L001_R1_001.fastq.gz", "x-s2_ambiguous.fastq.gz", "Index-hop.fq.gz"]
    ids = ["pre_S1", "pre_s2", "pre_S[", "S.*"]
    m = class_file_matrix(files, ids, "pre_")
    assert m.shape == (8, 4) and m.dtype == np.uint8
    for f, name in enumerate(files):
        assert m[0, f] == bool(re.search("undetermined", name, re.I))
        assert m[1, f] == bool(re.search("undetermined|index-hop", name, re.I))
        assert m[3, f] == bool(re.search("undetermined|ambiguous", name, re.I))
        assert m[4, f] == bool(re.search(re.compile("S1", re.I), name))
        assert m[5, f] == bool(re.search(re.compile("s2", re.I), name))
        assert m[6, f] == 2                                   # "S[" does not compile
        assert m[7, f] == bool(re.search(re.compile("S.*", re.I), name))
    assert (m[2] == 0).all()                                  # demuxable keys use the sample rows


def test_small_file_run_and_stream_policy(tmp_path, monkeypatch):
    from frender_b200 import cli
    sizes = [5 << 20, 6 << 20, 40 << 20, 1 << 20, 10, 2 << 20, 2 << 20]
    files = []
    for i, n in enumerate(sizes):
        p = tmp_path / f"f{i}.fastq.gz"
        with open(p, "wb") as fh:
            fh.truncate(n)
        files.append(p)
    assert cli.small_file_run(files, 0) == 2          # stops in front of the 40 MB file
    assert cli.small_file_run(files, 2) == 1          # a large file goes on its own
    assert cli.small_file_run(files, 3) == 1          # ... and so does a file in front of one that cannot be gzip
    assert cli.small_file_run(files, 4) == 1
    assert cli.small_file_run(files, 5) == 2
    monkeypatch.setattr(cli, "SMALL_RUN_BYTES", 8 << 20)
    assert cli.small_file_run(files, 0) == 1          # 5 + 6 MB exceed the run
    monkeypatch.delenv("FRENDER_MAX_STREAMS", raising=False)
    monkeypatch.delenv("FRB_GZ_DEVICE", raising=False)
    assert cli.concurrent_streams(8, files) == 1      # inflate on the device: one after the other
    monkeypatch.setenv("FRB_GZ_DEVICE", "0")
    assert cli.concurrent_streams(8, files) == 7      # host zlib: one thread per file
    assert cli.concurrent_streams(3, files) == 3
    monkeypatch.setenv("FRENDER_MAX_STREAMS", "2")
    assert cli.concurrent_streams(8, files) == 2


def test_initial_table_size_follows_the_input(tmp_path, monkeypatch):
    from frender_b200 import cli
    monkeypatch.delenv("FRENDER_TABLE_LOG2", raising=False)
    small, big = tmp_path / "a.gz", tmp_path / "b.gz"
    with open(small, "wb") as fh:
        fh.truncate(1 << 20)
    with open(big, "wb") as fh:
        fh.truncate(30 << 30)                          # sparse: a lane's worth of compressed bytes
    lo, hi = cli.initial_table_log2([small]), cli.initial_table_log2([big])
    assert 16 <= lo < hi <= 32
    monkeypatch.setenv("FRENDER_TABLE_LOG2", "19")
    assert cli.initial_table_log2([big]) == 19


def test_text_chunks_random_newline_mixes(tmp_path):
    """Random mixes of "\\r", "\\n", "\\r\\n" and text, random read sizes: always the bytes of text mode."""
    import random
    from frender_b200.cli import TextChunks
    rng = random.Random(20)
    for case in range(40):
        raw = b"".join(rng.choice([b"\r", b"\n", b"\r\n", b"A", b"@x:1", b"+", b"FF"]) for _ in range(rng.randint(0, 200)))
        p = tmp_path / f"m{case}.gz"
        p.write_bytes(gzip.compress(raw))
        want = gzip.open(p, "rt").read().encode()
        t, got = TextChunks(p), b""
        while not t.eof:
            got += t.read(rng.randint(1, 17))
        assert got == want, (case, raw)
