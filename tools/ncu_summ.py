"""Summarise an ncu --csv launch list (gpu__time_duration.sum) per kernel name."""
import csv
import sys
from collections import defaultdict

with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    name = r["Kernel Name"][:90]
    tot[name][0] += 1
    tot[name][1] += us
allus = sum(v[1] for v in tot.values())
print(f"total {allus/1000:.3f} ms over {sum(v[0] for v in tot.values())} launches")
for k, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{us/1000:10.3f} ms {100*us/allus:5.1f}%  {c:6d} x  avg {us/c:9.1f} us  {k}")
