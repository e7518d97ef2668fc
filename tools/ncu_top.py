"""Summarise an ncu --page source --csv dump: stall mix, hottest SASS lines.  usage: ncu_top.py file.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, h):
    try:
        return float(r[ix[h]])
    except (ValueError, KeyError):
        return 0.0


tot_s = sum(f(r, "# Samples") for r in data)
tot_i = sum(f(r, "Instructions Executed") for r in data)
print(f"kernel {rows[0][1]}  samples {tot_s:.0f}  warp-instr {tot_i:.0f}")
st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(f(r, h) for r in data) for h in st}
for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
    print(f"  {h:26s} {100 * v / max(sum(agg.values()), 1):5.1f}%")
base = int(data[0][ix["Address"]], 16)
print("hottest lines (offset, samples%, instr%, top stall, sass)")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:top]:
    stall = max(st, key=lambda h: f(r, h))
    print(f"  {int(r[ix['Address']], 16) - base:#7x} {100 * f(r, '# Samples') / tot_s:5.1f}% "
          f"{100 * f(r, 'Instructions Executed') / tot_i:5.2f}% {stall[6:]:14s} {r[ix['Source']][:70]}")
