"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==") and not l.startswith("#")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u in ("nsecond", "ns") else v * 1e3 if u in ("msecond", "ms") else v
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:80]
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f"total {tot/1e3:.2f} ms over {sum(n for n, _ in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f"{t:10.1f} us {100*t/tot:5.1f}%  n={n:4d}  {k}")
