"""Aggregate ONE step of an `ncu --metrics gpu__time_duration.sum --csv` launch list (the launches between the last two
AdamW launches): python tools/agg_step.py launches.csv > summary.txt"""
import collections
import csv
import re
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
L = [(r["Kernel Name"], float(r["Metric Value"].replace(",", "")) / 1000, r["Stream"]) for r in rows
     if r.get("Metric Name") == "gpu__time_duration.sum"]
ad = [i for i, x in enumerate(L) if "adamw_multi" in x[0]]
a, b = ad[-2] + 1, ad[-1] + 1


def short(n):
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("morec::", "")
    return n[:100]


agg = collections.defaultdict(lambda: [0, 0.0])
for n, t, s in L[a:b]:
    k = short(n)
    agg[k][0] += 1
    agg[k][1] += t
tot = sum(v[1] for v in agg.values())
print(f"one step: {b - a} launches, {tot / 1000:.3f} ms of serialised kernel time (cold caches, under ncu)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:9.1f} us {v[0]:5d} {100 * v[1] / tot:5.1f}%  {k}")
