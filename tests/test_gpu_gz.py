"""Device-side gzip inflate (csrc/gz_kernels.cuh) against zlib, through the C-ABI."""
import gzip
import os
import random
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from frender_b200.engine import Context
    c = Context(0, table_log2=16)
    yield c
    c.close()


def fastq(reads, seed=3):
    from frender_b200 import synth
    return synth.generate_big(synth.make_spec("C2"), seed * 1000, seed * 1000 + reads)


def gz_members(data, parts, level):
    step = (len(data) + parts - 1) // parts
    return b"".join(gzip.compress(data[o:o + step], level, mtime=0) for o in range(0, max(len(data), 1), max(step, 1)))


@pytest.mark.parametrize("level", [1, 6, 9])
def test_inflate_single_member_fastq(ctx, tmp_path, level, monkeypatch):
    """One member, many deflate blocks, small chunk stride: hundreds of chunks start at found block starts and
    resolve their markers through the windows of the chunks in front of them."""
    monkeypatch.setenv("FRB_GZ_STRIDE_KB", "32")
    data = fastq(60_000)                              # 22 MB of text
    p = tmp_path / "a.fastq.gz"
    p.write_bytes(gzip.compress(data, level, mtime=0))
    got = ctx.gz_inflate(p, len(data) + 1024)
    assert got is not None, "the device path declined an ordinary gzip stream"
    assert got == data


def test_inflate_across_pieces(ctx, tmp_path, monkeypatch):
    """Pieces of 1 MiB of compressed bytes: block starts found in the overlap, windows and an unfinished line
    carried from piece to piece."""
    monkeypatch.setenv("FRB_GZ_PIECE_MB", "1")
    monkeypatch.setenv("FRB_GZ_STRIDE_KB", "32")
    data = fastq(50_000, seed=5)
    p = tmp_path / "b.fastq.gz"
    p.write_bytes(gzip.compress(data, 6, mtime=0))
    assert os.path.getsize(p) > 3 << 20
    assert ctx.gz_inflate(p, len(data) + 1024) == data


def test_inflate_file_between_one_and_one_and_a_half_pieces(ctx, tmp_path, monkeypatch):
    """Such a file is ONE piece, larger than the nominal piece size (the buffers must follow the plan)."""
    monkeypatch.setenv("FRB_GZ_PIECE_MB", "1")
    n = 20_000
    while True:
        data = fastq(n, seed=9)
        blob = gzip.compress(data, 6, mtime=0)
        if len(blob) <= (1 << 20) + 65536:
            n = n * 5 // 4
        elif len(blob) >= (3 << 19) - 65536:
            n = n * 9 // 10
        else:
            break
    p = tmp_path / "c.fastq.gz"
    p.write_bytes(blob)
    assert ctx.gz_inflate(p, len(data) + 1024) == data


def test_inflate_members_and_odd_streams(ctx, tmp_path):
    rnd = random.Random(7)
    text = fastq(4000, seed=9)
    cases = {
        "multi_member": gz_members(text, 7, 6),
        "bgzf_like": gz_members(text, 40, 6),                    # many small members
        "empty": gzip.compress(b"", 6, mtime=0),
        "tiny": gzip.compress(b"@r 1:N:0:ACGT+ACGT\nA\n+\nF\n", 9, mtime=0),
        "stored": gzip.compress(text[:200_000], 0, mtime=0),     # stored blocks only
        "zero_padded": gzip.compress(text[:50_000], 6, mtime=0) + b"\0" * 37,
        "fname_extra": None,
        "incompressible": gzip.compress(bytes(rnd.getrandbits(8) for _ in range(300_000)).replace(b"\r", b"x"), 6, mtime=0),
        "runs": gzip.compress((b"F" * 5000 + b"\n") * 300, 6, mtime=0),
    }
    plain = {"multi_member": text, "bgzf_like": text, "empty": b"", "tiny": b"@r 1:N:0:ACGT+ACGT\nA\n+\nF\n",
             "stored": text[:200_000], "zero_padded": text[:50_000], "runs": (b"F" * 5000 + b"\n") * 300}
    # a member with FNAME and FEXTRA in its header, as gzip.open("wb") writes it (F:672)
    named = tmp_path / "named.gz"
    with gzip.GzipFile(named, "wb", compresslevel=9, mtime=0) as fh:
        fh.write(text[:70_000])
    cases["fname_extra"] = named.read_bytes()
    plain["fname_extra"] = text[:70_000]
    plain["incompressible"] = zlib.decompress(cases["incompressible"], 31)
    for name, blob in cases.items():
        p = tmp_path / f"{name}.gz"
        p.write_bytes(blob)
        got = ctx.gz_inflate(p, len(plain[name]) + 1024)
        assert got is None or got == plain[name], name
        if name in ("multi_member", "bgzf_like", "tiny", "zero_padded", "fname_extra", "runs"):
            assert got is not None, name


def test_scan_gz_uses_the_device_and_matches_oracle(ctx, tmp_path, monkeypatch):
    """frb_scan_gz end to end on the device path against the oracle's tally, small pieces."""
    import frender_oracle as O
    monkeypatch.setenv("FRB_GZ_PIECE_MB", "1")
    data = fastq(30_000, seed=11)
    p = tmp_path / "Undetermined_S0_L001_R1_001.fastq.gz"
    p.write_bytes(gzip.compress(data, 6, mtime=0))
    want = O.tally_barcodes(1, [p])
    ctx.reset()
    reads, uniq, raw = ctx.scan_gz(p, 0)
    assert reads == 30_000 and raw == len(data) and uniq == len(want["total"])
    assert list(ctx.counter()["total"].items()) == list(want["total"].items())


def test_truncated_and_foreign_files_fail(ctx, tmp_path):
    """A .gz cut mid-stream raises in the reference (EOFError from gzip, F:159), a file that is not gzip raises
    BadGzipFile: both are FRB_ERR_IO here, on either inflate path."""
    from frender_b200 import _lib
    from frender_b200.engine import FrbError
    data = fastq(20_000, seed=13)
    blob = gzip.compress(data, 6, mtime=0)
    for name, content in (("cut.fastq.gz", blob[:len(blob) * 2 // 3]), ("cut_trailer.fastq.gz", blob[:-5]),
                          ("plain.fastq.gz", data[:100_000])):
        p = tmp_path / name
        p.write_bytes(content)
        ctx.reset()
        with pytest.raises(FrbError) as info:
            ctx.scan_gz(p, 0)
        assert info.value.code == _lib.ERR_IO, name
    ctx.reset()


@pytest.mark.parametrize("piece_mb", ["128", "1"])
def test_member_trailers_are_checked(ctx, tmp_path, monkeypatch, piece_mb):
    """CRC32 and ISIZE of every gzip member (RFC 1952) are verified on the device: a wrong trailer is what
    gzip.open reports as BadGzipFile in the reference (F:159) -- FRB_ERR_IO here; members that span pieces and
    chunks, several members per file and BGZF-sized members pass."""
    from frender_b200 import _lib
    from frender_b200.engine import FrbError
    monkeypatch.setenv("FRB_GZ_PIECE_MB", piece_mb)
    data = fastq(40_000, seed=21)
    one = gzip.compress(data, 6, mtime=0)
    three = gz_members(data, 3, 6)
    tiny = gz_members(data, 300, 6)                       # ~ 50 KB of text per member, like BGZF
    first_len = len(gzip.compress(data[:(len(data) + 2) // 3], 6, mtime=0))
    good = {"one": one, "three": three, "tiny": tiny}
    for name, blob in good.items():
        p = tmp_path / f"{name}.fastq.gz"
        p.write_bytes(blob)
        assert ctx.gz_inflate(p, len(data) + 1024) == data, name
    flip = lambda blob, at, bit=1: blob[:at] + bytes([blob[at] ^ bit]) + blob[at + 1:]
    n = len(one)
    bad = {"crc_last": flip(one, n - 8), "isize_last": flip(one, n - 2, 0x10),
           "crc_first_member": flip(three, first_len - 7), "isize_first_member": flip(three, first_len - 4),
           "crc_tiny_member": flip(tiny, len(tiny) - 6)}
    for name, blob in bad.items():
        p = tmp_path / f"{name}.fastq.gz"
        p.write_bytes(blob)
        with pytest.raises(FrbError) as info:
            ctx.gz_inflate(p, len(data) + 1024)
        assert info.value.code == _lib.ERR_IO, name
        with pytest.raises((gzip.BadGzipFile, EOFError, zlib.error)):
            gzip.open(p, "rb").read()


def test_crlf_file_larger_than_the_staging_buffer(ctx, tmp_path, monkeypatch):
    """Universal newlines on the host path (the device path hands '\\r' files over): "\\r\\n", lone "\\r" and a
    "\\r\\n" split across reads, with staging buffers far smaller than the file."""
    import frender_oracle as O
    monkeypatch.setenv("FRB_STAGE_MB", "1")
    from frender_b200.engine import Context
    text = fastq(12_000, seed=17).decode()
    lines = text.split("\n")
    rnd = random.Random(1)
    crlf = "".join(l + rnd.choice(["\r\n", "\r\n", "\r", "\n"]) for l in lines[:-1])
    p = tmp_path / "crlf_R1.fastq.gz"
    p.write_bytes(gzip.compress(crlf.encode(), 1, mtime=0))
    want = O.tally_barcodes(1, [p])
    c2 = Context(0, table_log2=16)       # its own context: staging buffers are sized at first use
    reads, uniq, _ = c2.scan_gz(p, 0)
    assert reads == 12_000 and list(c2.counter()["total"].items()) == list(want["total"].items())
    c2.close()
