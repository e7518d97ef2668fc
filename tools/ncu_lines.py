"""Join an ncu SASS source dump with nvdisasm -g line info: samples / instructions per source line.
usage: ncu_lines.py ncu_source.csv disasm_with_lineinfo.txt [N]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
base = int(data[0][ix["Address"]], 16)
line_of = {}
cur = None
for ln in open(sys.argv[2]):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_s = tot_i = 0.0
for r in data:
    off = int(r[ix["Address"]], 16) - base
    key = line_of.get(off, ("?", 0))
    s, i = float(r[ix["# Samples"]] or 0), float(r[ix["Instructions Executed"]] or 0)
    agg[key][0] += s
    agg[key][1] += i
    for h in st:
        agg[key][2][h[6:]] += float(r[ix[h]] or 0)
    tot_s += s
    tot_i += i
src = {}
for key in agg:
    if key[0] != "?" and key[0] not in src:
        try:
            src[key[0]] = open(f"/root/repo/frender_b200/csrc/{key[0]}").read().split("\n")
        except OSError:
            src[key[0]] = []
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
print(f"samples {tot_s:.0f} warp-instr {tot_i:.0f}")
for key, (s, i, c) in sorted(agg.items(), key=lambda x: -x[1][0])[:n]:
    text = src.get(key[0], [])
    code = text[key[1] - 1].strip()[:80] if 0 < key[1] <= len(text) else ""
    top = ",".join(f"{k}:{100 * v / max(s, 1):.0f}" for k, v in c.most_common(2))
    print(f"{key[0]}:{key[1]:<4d} {100 * s / tot_s:5.1f}% smp {100 * i / tot_i:5.1f}% ins  [{top}]  {code}")
