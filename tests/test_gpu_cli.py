"""The drop-in CLI against files the reference itself wrote (tests/golden)."""
import gzip
import hashlib
import os

import pytest

from conftest import unb64

pytestmark = pytest.mark.gpu


def run_cli(argv, cwd):
    from frender_b200.cli import main
    old = os.getcwd()
    os.chdir(cwd)
    try:
        main(argv)
    finally:
        os.chdir(old)


def test_cli_three_file_kat(golden, tmp_path, capsys):
    case = golden["cli3"]
    (tmp_path / "sheet.csv").write_text(case["sheet"])
    files = []
    for name, data in case["files"].items():
        with open(tmp_path / name, "wb") as fh:
            fh.write(gzip.compress(unb64(data)))
        files.append(str(tmp_path / name))
    run_cli(["scan", "-n", "1", "-rc", "-o", "kat", "-b", str(tmp_path / "sheet.csv")] + files, tmp_path)
    produced = sorted(f for f in os.listdir(tmp_path) if f.startswith("frender-"))
    scan_csv = [f for f in produced if "scan-results" in f][0]
    calls_csv = [f for f in produced if "index-2-calls" in f][0]
    assert scan_csv.startswith("frender-scan-results_1-mismatches_kat_") and scan_csv.endswith("_UTC.csv")
    assert (tmp_path / scan_csv).read_bytes() == unb64(case["scan_csv"])
    assert (tmp_path / calls_csv).read_bytes() == unb64(case["calls_csv"])
    out = capsys.readouterr().out
    assert "Incorrectly demultiplexed barcodes found! Affected files:" in out
    assert "S1\tCCCC\t3\tGGGG\t0\tforward" in out


@pytest.mark.parametrize("name", ["c1", "c2", "c2n0", "c3", "sampled", "c2_384"])
def test_cli_scan_csv_bytes(golden, golden_dir, tmp_path, name):
    """Byte-identical scan-results CSV (and index-2-calls CSV) on the single-file cases."""
    case = golden["scan"][name]
    (fname, _), = case["files"].items()
    src = os.path.join(golden_dir, f"{name}__{fname}")
    dst = tmp_path / fname
    dst.write_bytes(open(src, "rb").read())
    (tmp_path / "SampleSheet.csv").write_text(case["sheet_csv"])
    argv = ["scan", "-n", str(case["n"]), "-b", str(tmp_path / "SampleSheet.csv")]
    if case["rc"]:
        argv.append("-rc")
    if case["sample"]:
        argv += ["-s", str(case["sample"])]
    run_cli(argv + [str(dst)], tmp_path)
    out_name = f"frender-scan-results_{case['n']}-mismatches_{fname}.csv"
    assert (tmp_path / out_name).read_bytes() == unb64(case["scan_csv"])
    if case["rc"]:
        calls = out_name.replace("frender-scan-results_", "frender-index-2-calls_")
        assert (tmp_path / calls).read_bytes() == unb64(case["rc_calls_csv"])


@pytest.mark.parametrize("cores,mode", [("1", "device"), ("3", "device"), ("3", "streams"), ("3", "zlib")])
def test_cli_scan_multi_file_prefix(golden, golden_dir, tmp_path, cores, mode, monkeypatch):
    """Three files, `-p` prefix.  `-c 3`: one after the other with the inflate on the device (default), on three
    contexts of the same GPU at the same time when asked for (FRENDER_MAX_STREAMS) or with host zlib."""
    if mode == "streams":
        monkeypatch.setenv("FRENDER_MAX_STREAMS", "3")
    if mode == "zlib":
        monkeypatch.setenv("FRB_GZ_DEVICE", "0")       # the policy follows it (the library reads it once per process)
    case = golden["scan"]["multi"]
    files = []
    for fname in case["files"]:
        dst = tmp_path / fname
        dst.write_bytes(open(os.path.join(golden_dir, f"multi__{fname}"), "rb").read())
        files.append(str(dst))
    (tmp_path / "SampleSheet.csv").write_text(case["sheet_csv"])
    run_cli(["scan", "-n", "1", "-rc", "-c", cores, "-p", case["prefix"], "-o", "m", "-b",
             str(tmp_path / "SampleSheet.csv")] + files, tmp_path)
    out = [f for f in os.listdir(tmp_path) if f.startswith("frender-scan-results_")][0]
    assert (tmp_path / out).read_bytes() == unb64(case["scan_csv"])


@pytest.mark.parametrize("name", ["c1", "c1_ia", "c1_short_r2", "c4"])
@pytest.mark.parametrize("chunk_mb", ["64", "0"])
def test_cli_demux_streams(golden, tmp_path, name, chunk_mb, monkeypatch):
    """Same sink files, same decompressed bytes as the reference's demux (F:733-814)."""
    from frender_b200 import synth
    case = golden["demux"][name]
    scan = golden["scan"][case["scan_case"]]
    spec = synth.make_spec(scan["config"], n_samples=scan["n_samples"])
    r1 = synth.generate_big(spec, 0, case["reads"], 1)
    r2 = synth.generate_big(spec, 0, case["reads"], 2)
    if case["truncate_r2"]:
        pos = -1
        for _ in range(4 * case["truncate_r2"] + 1):
            pos = r2.index(b"\n", pos + 1)
        r2 = r2[:pos + 11]
    p1 = tmp_path / "Undetermined_S0_L001_R1_001.fastq.gz"
    p2 = tmp_path / "Undetermined_S0_L001_R2_001.fastq.gz"
    p1.write_bytes(gzip.compress(r1, 1))
    p2.write_bytes(gzip.compress(r2, 1))
    (tmp_path / "results.csv").write_bytes(unb64(case["results_csv"]))
    if chunk_mb == "0":
        monkeypatch.setenv("FRENDER_DEMUX_CHUNK_MB", "0")      # forces the tiny-window path (many carries)
        import frender_b200.cli as cli
        monkeypatch.setattr(cli, "MIN_CHUNK", 50_000, raising=False)
    flags = case["flags"]
    argv = ["demux", "-r", str(tmp_path / "results.csv"), "-d", str(tmp_path / "out")]
    argv += ["-i"] if flags.get("i") else []
    argv += ["-a"] if flags.get("a") else []
    argv += ["-o", flags["o"]] if flags.get("o") else []
    run_cli(argv + [str(p1), str(p2)], tmp_path)
    got = sorted(os.listdir(tmp_path / "out"))
    assert got == sorted(case["sinks"])
    for fname, want in case["sinks"].items():
        raw = gzip.open(tmp_path / "out" / fname, "rb").read()
        assert len(raw) == want["bytes"], fname
        assert hashlib.sha256(raw).hexdigest() == want["sha256"], fname


def test_cli_demux_unknown_key(golden, tmp_path):
    rows = "idx1,idx2,reads,matched_idx1,matched_idx2,read_type,sample_name,demux_ok\r\nAAAA,CCCC,1,AAAA,CCCC,demuxable,S1,True\r\n"
    (tmp_path / "r.csv").write_text(rows)
    rec = lambda k, n: f"@r:1 {n}:N:0:{k}\nACGT\n+\nFFFF\n".encode()
    (tmp_path / "x_R1_001.fastq.gz").write_bytes(gzip.compress(rec("AAAA+CCCC", 1) + rec("GGGG+TTTT", 1)))
    (tmp_path / "x_R2_001.fastq.gz").write_bytes(gzip.compress(rec("AAAA+CCCC", 2) + rec("GGGG+TTTT", 2)))
    with pytest.raises(SystemExit, match="Couldn't find barcode GGGG\\+TTTT in supplied frender result file!"):
        run_cli(["demux", "-r", str(tmp_path / "r.csv"), "-d", str(tmp_path / "o"),
                 str(tmp_path / "x_R1_001.fastq.gz"), str(tmp_path / "x_R2_001.fastq.gz")], tmp_path)


def test_cli_scan_grows_a_full_table(golden, golden_dir, tmp_path, monkeypatch, capsys):
    """More unique keys than table slots: the reference's dict never refuses a key (F:172-177); the CLI re-creates
    the tables four times as large and tallies again.  Same CSV bytes as the reference."""
    case = golden["scan"]["c2_384"]                      # 2614 unique keys
    (fname, _), = case["files"].items()
    dst = tmp_path / fname
    dst.write_bytes(open(os.path.join(golden_dir, f"c2_384__{fname}"), "rb").read())
    (tmp_path / "SampleSheet.csv").write_text(case["sheet_csv"])
    monkeypatch.setenv("FRENDER_TABLE_LOG2", "10")       # 1024 slots
    run_cli(["scan", "-n", "1", "-rc", "-b", str(tmp_path / "SampleSheet.csv"), str(dst)], tmp_path)
    assert "tallying again with 2^12" in capsys.readouterr().out
    assert (tmp_path / f"frender-scan-results_1-mismatches_{fname}.csv").read_bytes() == unb64(case["scan_csv"])


def test_cli_demux_accepts_the_scan_layout(golden, tmp_path):
    """Extension (SURVEY finding 1): `demux -r` takes the CSV exactly as `scan` wrote it; same sinks, same bytes as
    with the reordered file the reference needs.  Any other header still fails with the reference's assertion."""
    from frender_b200 import synth
    case = golden["demux"]["c1"]
    scan = golden["scan"]["c1"]
    spec = synth.make_spec(scan["config"], n_samples=scan["n_samples"])
    p1 = tmp_path / "Undetermined_S0_L001_R1_001.fastq.gz"
    p2 = tmp_path / "Undetermined_S0_L001_R2_001.fastq.gz"
    p1.write_bytes(gzip.compress(synth.generate_big(spec, 0, case["reads"], 1), 1))
    p2.write_bytes(gzip.compress(synth.generate_big(spec, 0, case["reads"], 2), 1))
    (tmp_path / "scan.csv").write_bytes(unb64(scan["scan_csv"]))
    run_cli(["demux", "-r", str(tmp_path / "scan.csv"), "-d", str(tmp_path / "out"), str(p1), str(p2)], tmp_path)
    assert sorted(os.listdir(tmp_path / "out")) == sorted(case["sinks"])
    for fname, want in case["sinks"].items():
        raw = gzip.open(tmp_path / "out" / fname, "rb").read()
        assert hashlib.sha256(raw).hexdigest() == want["sha256"], fname
    (tmp_path / "bad.csv").write_text("a,b,c\r\n1,2,3\r\n")
    with pytest.raises(AssertionError, match="does not appear to be a valid frender result file"):
        run_cli(["demux", "-r", str(tmp_path / "bad.csv"), "-d", str(tmp_path / "out2"), str(p1), str(p2)], tmp_path)
