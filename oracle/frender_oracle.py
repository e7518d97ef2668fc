"""CPU oracle for the frender scan/demux hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a plain-Python restatement of the algorithms in the reference
(`/root/reference/frender.py`, abbreviated F: below).  It exists so that the
CUDA path can be checked bit-for-bit; it is NOT part of the product.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import it.  The product (`frender_b200/`) never does and
fails loudly when its CUDA library is missing.

Parity status: PINNED.  The reference is pure Python and was imported in the
build container by `tests/golden/make_golden.py`; the fixtures it wrote under
`tests/golden/` hold reference outputs (tally dicts, `process()` results of both
passes, orientation calls, demux_ok flags, CSV bytes, demux streams) and
`tests/test_oracle_golden.py` checks every function here against them.  The one
exception is single-index matching, which the reference cannot run at all
(F:104-107, F:306) -- see `match_single_index` ("parity unpinned").

Every function cites the reference lines it follows.  The structure is
deliberately that of the reference (text-mode gzip, str.split, dict counters,
per-character Hamming loops) so that timing it is a fair stand-in for timing the
reference on a box where `/root/reference` does not exist.
"""
import csv
import gzip
import os
import re
from multiprocessing import Pool

READ_TYPES = ("undetermined", "index_hop", "demuxable", "ambiguous")


# ----------------------------------------------------------------------------
# Hot path A: read-name parse + unique-combination counter
# ----------------------------------------------------------------------------
def scan_key(header_line):
    """Index string of one R1 header line, scan rule (F:168-170): strip trailing
    newlines, take the 2nd space-separated token, then its last ':' field."""
    return header_line.rstrip("\n").split(" ")[1].split(":")[-1]


def demux_key(header_line):
    """Index string of one R2 header line, demux rule (F:778): last ':' field of
    the whole line, trailing newlines stripped afterwards."""
    return header_line.split(":")[-1].rstrip("\n")


def tally_text(lines, sample=None):
    """Count keys of every 4th line of an iterable of text lines (F:160-177).
    Returns (counter dict in first-seen order, reads visited)."""
    counter = {}
    visited = 0
    for lineno, line in enumerate(lines):
        if lineno & 3:
            continue
        if sample and visited >= sample:
            break
        visited += 1
        key = scan_key(line)
        counter[key] = counter.get(key, 0) + 1
    return counter, visited


def scan_file(path, sample=None):
    """(basename, per-file counter, reads) for one fastq.gz (F:154-181).  Text
    mode, so universal newlines apply exactly as in the reference (F:159)."""
    with gzip.open(path, "rt") as handle:
        counter, visited = tally_text(handle, sample)
    return os.path.basename(str(path)), counter, visited


def tally_barcodes(cores, files, sample=None):
    """{"total": merged counter, basename: per-file counter, ...} (F:183-207).
    File-level process pool when cores > 1 (F:189-193); the merge walks files in
    list order so "total" is in first-appearance order (F:199-205); files that
    share a basename overwrite each other's per-file entry (F:204-205)."""
    if sample:
        assert sample >= 1
    jobs = [(f, sample) for f in files]
    if cores > 1 and len(jobs) > 1:
        with Pool(processes=cores) as pool:
            per_file = pool.starmap(scan_file, jobs)
    else:
        per_file = [scan_file(*j) for j in jobs]
    merged = {}
    for _, counter, _ in per_file:
        for key, n in counter.items():
            merged[key] = merged.get(key, 0) + n
    out = {"total": merged}
    for name, counter, _ in per_file:
        out[name] = counter
    return out


# ----------------------------------------------------------------------------
# Hot path B: mismatch matcher, classifier, RC wrapper, orientation call
# ----------------------------------------------------------------------------
_RC_TABLE = str.maketrans("ATGCNatgcn", "TACGNtacgn")


def reverse_complement(seq):
    """F:210-211.  Characters outside ATGCNatgcn pass through unchanged."""
    return seq.translate(_RC_TABLE)[::-1]


def approx_match_rows(query, sheet_column, max_subs):
    """Ascending sheet row numbers within `max_subs` substitutions of `query`
    (F:214-234).  Case-insensitive; unequal lengths are an AssertionError with
    the reference's message (F:227-229)."""
    rows = []
    q = query.lower()
    for row, cand in enumerate(sheet_column):
        c = cand.lower()
        assert len(q) == len(c), f"Barcode {q} doesn't match length of supplied barcode {c}"
        if sum(1 for x, y in zip(q, c) if x != y) <= max_subs:
            rows.append(row)
    return rows


def classify(idx1, idx2, sheet_idx1, sheet_idx2, sheet_ids, max_subs):
    """One index pair against the sheet (F:237-291): first matching row supplies
    the reported strings (F:261-262); the size of the row-set intersection picks
    index_hop / demuxable / ambiguous (F:264-278); a miss on either side blanks
    everything (F:280-284)."""
    rows1 = approx_match_rows(idx1, sheet_idx1, max_subs)
    rows2 = approx_match_rows(idx2, sheet_idx2, max_subs)
    out = {"matched_idx1": "", "matched_idx2": "", "read_type": "undetermined", "sample_name": ""}
    if rows1 and rows2:
        out["matched_idx1"] = sheet_idx1[rows1[0]]
        out["matched_idx2"] = sheet_idx2[rows2[0]]
        both = set(rows1) & set(rows2)
        if not both:
            out["read_type"] = "index_hop"
        elif len(both) == 1:
            out["read_type"] = "demuxable"
            out["sample_name"] = sheet_ids[both.pop()]
        else:
            out["read_type"] = "ambiguous"
    return out


def classify_with_rc(key, reads, sheet_idx1, sheet_idx2, sheet_ids, max_subs, rc_mode):
    """F:294-351.  Key order of the returned dict follows the reference:
    matched_idx1, matched_idx2, read_type, sample_name, reads[, matched_rc_idx2,
    rc_read_type, rc_sample_name]."""
    idx1, idx2 = key.split("+")[0:2]
    fwd = classify(idx1, idx2, sheet_idx1, sheet_idx2, sheet_ids, max_subs)
    fwd["reads"] = reads
    if not rc_mode:
        return fwd
    flipped = [reverse_complement(s) for s in sheet_idx2]          # F:315
    rc = classify(idx1, idx2, sheet_idx1, flipped, sheet_ids, max_subs)
    if fwd["matched_idx1"] == "":                                  # F:319-323
        fwd["matched_idx1"] = rc["matched_idx1"]
    fwd["matched_rc_idx2"] = rc["matched_idx2"]
    fwd["rc_read_type"] = rc["read_type"]
    fwd["rc_sample_name"] = rc["sample_name"]
    if fwd["read_type"] == "demuxable" and rc["read_type"] == "demuxable":
        if fwd["sample_name"] != rc["sample_name"]:                # F:336-349
            fwd["read_type"] = fwd["rc_read_type"] = "ambiguous"
            fwd["sample_name"] = fwd["rc_sample_name"] = ""
    return fwd


def _classify_job(args):
    return classify_with_rc(*args)


def process(cores, counter, indexes, max_subs, rc_mode):
    """{key: classification} over the unique keys of `counter` (F:391-426)."""
    jobs = [
        (k, n, indexes["idx1"], indexes["idx2"], indexes["id"], max_subs, rc_mode)
        for k, n in counter.items()
    ]
    if cores > 1:
        with Pool(processes=cores) as pool:
            res = pool.map(_classify_job, jobs, chunksize=max(1, len(jobs) // (cores * 8) or 1))
    else:
        res = [classify_with_rc(*j) for j in jobs]
    return dict(zip(counter.keys(), res))


def call_rc_mode_per_id(records, ids):
    """Per sample *name*: reads supporting forward vs reverse-complement i5;
    use_rc only when forward < rc (F:354-388)."""
    assert "rc_read_type" in records[0]
    acc = {name: [0, 0] for name in ids}
    for rec in records:
        if rec["sample_name"] != "":
            acc[rec["sample_name"]][0] += int(rec["reads"])
        if rec["rc_sample_name"] != "":
            acc[rec["rc_sample_name"]][1] += int(rec["reads"])
    return {name: {"call": f < r, "reads_f": f, "reads_rc": r} for name, (f, r) in acc.items()}


def orient_sheet(indexes, rc_calls):
    """Second-pass sheet: idx2 replaced row by row according to the call for the
    row's sample name (F:618-623)."""
    return {
        "id": list(indexes["id"]),
        "idx1": list(indexes["idx1"]),
        "idx2": [
            reverse_complement(s) if rc_calls[name]["call"] else s
            for s, name in zip(indexes["idx2"], indexes["id"])
        ],
    }


def flatten(results):
    """List of CSV rows, idx1/idx2 first (F:482-492)."""
    rows = []
    for key, rec in results.items():
        parts = key.split("+")
        row = {"idx1": parts[0], "idx2": parts[1]}
        row.update(rec)
        rows.append(row)
    return rows


def scan_analysis(cores, counter, indexes, max_subs, rc_mode):
    """The matcher part of frender_scan (F:610-630): first pass, optional
    orientation call and oriented second pass.  Returns (results, rc_calls or
    None, sheet used for the final pass)."""
    results = process(cores, counter["total"], indexes, max_subs, rc_mode)
    rc_calls = None
    if rc_mode:
        rc_calls = call_rc_mode_per_id(flatten(results), indexes["id"])
        indexes = orient_sheet(indexes, rc_calls)
        results = process(cores, counter["total"], indexes, max_subs, False)
    return results, rc_calls, indexes


def demux_ok(counter, results, prefix=""):
    """Adds results[key]["demux_ok"] and returns the set of file names that hold
    a key they should not (F:504-564).  The sample name is used as a regex."""
    files = [f for f in counter if f != "total"]
    bad_files = set()
    fixed = {
        "undetermined": re.compile("undetermined", re.I),
        "index_hop": re.compile("undetermined|index-hop", re.I),
        "ambiguous": re.compile("undetermined|ambiguous", re.I),
    }
    for key, rec in results.items():
        kind = rec["read_type"]
        if kind in fixed:
            pattern = fixed[kind]
        else:
            assert kind == "demuxable", f"Strange read type ('{kind}') found"
            pattern = re.compile(rec["sample_name"].removeprefix(prefix), re.I)
        verdicts = [(counter[f].get(key, 0) == 0) or bool(pattern.search(f)) for f in files]
        if files:
            rec["demux_ok"] = all(verdicts)
        bad_files.update(f for f, ok in zip(files, verdicts) if not ok)
    return results, bad_files


def scan_csv_bytes(results):
    """The scan-results CSV exactly as csv.DictWriter(newline="") writes it
    (F:495-501): header from the first row's key order, CRLF line ends."""
    import io

    rows = flatten(results)
    buf = io.StringIO(newline="")
    w = csv.DictWriter(buf, rows[0].keys())
    w.writeheader()
    w.writerows(rows)
    return buf.getvalue().encode()


def rc_calls_csv_bytes(rc_calls, indexes):
    """The index-2-calls CSV (F:456-479); `indexes` is the sheet as supplied."""
    import io

    buf = io.StringIO(newline="")
    w = csv.writer(buf)
    w.writerow(["sample_name", "supplied_index_2", "reads_supplied_index_2", "rc_index_2",
                "reads_rc_index_2", "use_rc"])
    for name, call in rc_calls.items():
        row = indexes["id"].index(name)
        supplied = indexes["idx2"][row]
        w.writerow([name, supplied, call["reads_f"], reverse_complement(supplied), call["reads_rc"],
                    "TRUE" if call["call"] else "FALSE"])
    return buf.getvalue().encode()


def match_single_index(key, sheet_idx1, sheet_ids, max_subs):
    """EXTENSION, parity unpinned: the reference cannot classify single-index
    keys (F:104-107 SystemExit, F:306 ValueError).  Defined as the idx1-only
    half of `classify`: 0 rows -> undetermined, 1 -> demuxable, >1 -> ambiguous."""
    rows = approx_match_rows(key, sheet_idx1, max_subs)
    if not rows:
        return {"matched_idx1": "", "read_type": "undetermined", "sample_name": ""}
    if len(rows) == 1:
        return {"matched_idx1": sheet_idx1[rows[0]], "read_type": "demuxable",
                "sample_name": sheet_ids[rows[0]]}
    return {"matched_idx1": sheet_idx1[rows[0]], "read_type": "ambiguous", "sample_name": ""}


# ----------------------------------------------------------------------------
# Hot path C: demux record router
# ----------------------------------------------------------------------------
DEMUX_HEADER = ["idx1", "idx2", "reads", "matched_idx1", "matched_idx2", "read_type", "sample_name"]


def parse_results_file(path):
    """{key: (read_type, sample_id)} by column POSITION 0,1,5,6 (F:645-664)."""
    with open(path, newline="") as handle:
        rows = csv.reader(handle)
        header = next(rows)
        assert header[0:7] == DEMUX_HEADER, f"${path} does not appear to be a valid frender result file!"
        return {r[0] + "+" + r[1]: (r[5], r[6]) for r in rows}


def sink_names(table, index_hop=True, ambiguous=True, undeter=True, samples=True):
    """Names of the sinks frender_demux opens (F:736-759), as a dict
    role -> name (role is a sample id or one of '#hop', '#amb', '#und')."""
    und = "Undetermined" + ("-ambiguous" if ambiguous else "") + ("-index-hop" if index_hop else "")
    roles = {}
    if samples:
        for sid in set(v[1] for v in table.values()) - {""}:
            roles[sid] = sid
    roles["#und"] = und if undeter else None
    roles["#hop"] = "Index-hop" if index_hop else roles["#und"]
    roles["#amb"] = "Ambiguous" if ambiguous else roles["#und"]
    return roles


def route_pairs(r1_lines, r2_lines, table, roles):
    """Route 4-line record pairs (F:774-810).  Returns {sink name: (R1 bytes,
    R2 bytes)} of the decompressed streams.  The key comes from the R2 header
    (F:778); the shorter file ends the loop and a trailing partial record is
    padded with "" (F:719-723, F:777).  Unknown key -> SystemExit (F:807-810)."""
    sinks = {name: ([], []) for name in roles.values() if name}

    def groups(lines):
        block = []
        for line in lines:
            block.append(line)
            if len(block) == 4:
                yield block
                block = []
        if block:
            yield block + [""] * (4 - len(block))

    have_samples = any(not r.startswith("#") for r in roles)
    for rec1, rec2 in zip(groups(r1_lines), groups(r2_lines)):
        key = demux_key(rec2[0])
        if key not in table:
            raise SystemExit(f"Couldn't find barcode {key} in supplied frender result file!")
        kind, sid = table[key]
        if kind == "demuxable" and have_samples:
            name = roles[sid]
        elif kind == "index_hop" and roles["#hop"]:
            name = roles["#hop"]
        elif kind == "ambiguous" and roles["#amb"]:
            name = roles["#amb"]
        elif kind == "undetermined" and roles["#und"]:
            name = roles["#und"]
        else:
            raise SystemExit("Unrecognized read type found in supplied frender result file!")
        sinks[name][0].extend(rec1)
        sinks[name][1].extend(rec2)
    return {n: ("".join(a).encode(), "".join(b).encode()) for n, (a, b) in sinks.items()}


def demux_files(r1_path, r2_path, table, roles):
    """route_pairs over two fastq.gz files opened in text mode (F:776)."""
    with gzip.open(r1_path, "rt") as a, gzip.open(r2_path, "rt") as b:
        return route_pairs(a, b, table, roles)


# ----------------------------------------------------------------------------
# Sample sheet (host, cold) -- restated only so the oracle is self-contained
# ----------------------------------------------------------------------------
def read_sheet(path):
    """{"id": [...], "idx1": [...], "idx2": [...]} (F:52-116): skip an Illumina
    [Header]..[Data] preamble, find the id / index / index2 columns by regex."""
    with open(path, "r") as handle:
        rows = list(csv.reader(handle))
    start = 0
    if re.search(r"\[Header\]", rows[0][0]):
        start = 1
        while not re.search(r"\[Data\]", rows[start][0]):
            start += 1
        start += 1
    header = rows[start]

    def col(pattern, veto=None):
        for i, name in enumerate(header):
            if re.search(pattern, name, re.I) and not (veto and re.search(veto, name, re.I)):
                return i
        raise SystemExit(f'Couldn\'t find column matching "{pattern}" in csv header {header}')

    c_id, c_1, c_2 = col("id|name"), col("index", "id|2"), col("index.*2")
    body = rows[start + 1:]
    return {"id": [r[c_id] for r in body], "idx1": [r[c_1] for r in body], "idx2": [r[c_2] for r in body]}
