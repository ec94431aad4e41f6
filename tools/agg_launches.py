"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import re
import sys

path = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else None
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.OrderedDict()
tot = 0.0
seq = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(row["Metric Unit"], 1.0)
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
    seq.append((name, v))
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{v:8.3f} ms {c:4d}  {100 * v / tot:5.1f}%  {k}")
print(f"total {tot:.3f} ms, {len(seq)} launches")
if pat:
    for n, v in seq:
        if re.search(pat, n):
            print(f"{v:7.3f} {n}")
