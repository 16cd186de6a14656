"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel + grid."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.OrderedDict()
unit = "?"
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    key = (name[:60], row.get("Grid Size", ""), row.get("Block Size", ""))
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += v
scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3}.get(unit, 1e-3)
tot = sum(a[1] for a in agg.values())
print(f"total {tot * scale:.1f} us over {sum(a[0] for a in agg.values())} launches (unit {unit})")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1] / tot * 100:5.1f}%  n={a[0]:3d}  avg={a[1] / a[0] * scale:8.1f} us  {k[0]} grid={k[1]} block={k[2]}")
