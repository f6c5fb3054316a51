"""Per-region view of an ncu source page: groups consecutive SASS instructions with the same
execution count and prints instruction share, sample share and the stall mix of each group.
usage: ncu -i rep --page source --csv > src.csv; python tools/ncu_regions.py src.csv [min_share%]"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_")]
body = [r for r in rows[2:] if len(r) > 6 and r[col["Instructions Executed"]].isdigit()]
regions = []
for i, r in enumerate(body):
    e = int(r[col["Instructions Executed"]]); s = int(r[col["# Samples"]]); t = r[col["Source"]].strip()
    op = (t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0]
    if not regions or regions[-1]["e"] != e:
        regions.append({"e": e, "start": i, "n": 0, "s": 0, "ops": defaultdict(int), "st": defaultdict(int)})
    g = regions[-1]
    g["n"] += 1; g["s"] += s; g["ops"][op] += 1; g["end"] = i
    for h in stalls:
        v = r[col[h]]
        if v.isdigit(): g["st"][h[6:]] += int(v)
tot = sum(g["e"] * g["n"] for g in regions); stot = sum(g["s"] for g in regions)
print("total warp instructions", tot, "samples", stot)
for g in regions:
    w = g["e"] * g["n"]
    if 100 * w / tot >= minshare or 100 * g["s"] / stot >= minshare:
        st = sorted(g["st"].items(), key=lambda kv: -kv[1])[:5]
        print(f"{g['start']:5d}-{g['end']:5d} exec {g['e']:9d} n {g['n']:4d} instr {100*w/tot:5.1f}% samples {100*g['s']/stot:5.1f}%",
              dict(sorted(g["ops"].items(), key=lambda kv: -kv[1])[:6]), [(k, round(100 * v / max(1, g['s']))) for k, v in st])
