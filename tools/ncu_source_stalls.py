"""Summarises the per-instruction stall samples of an `ncu --page source --csv` dump.

usage: ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_source_stalls.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: sum(int(r[ix[s]]) for r in body) for s in stalls}
n = sum(int(r[ix["# Samples"]]) for r in body)
print("samples", n, {s[6:]: round(v / n, 3) for s, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v})
for s in ("stall_long_sb", "stall_math", "stall_wait", "stall_dispatch", "stall_short_sb"):
    print("--", s)
    order = sorted(range(len(body)), key=lambda i: -int(body[i][ix[s]]))[:top]
    for i in sorted(order):
        r = body[i]
        print(f"  {i:5d} {int(r[ix[s]]):6d}  {r[ix['Source']].strip()}")
