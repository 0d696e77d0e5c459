"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    name = re.sub(r"\(.*", "", r[ki])
    tot[name][0] += 1; tot[name][1] += v
allt = sum(v[1] for v in tot.values())
print(f"total {allt / 1e3:.2f} ms over {sum(v[0] for v in tot.values())} launches")
for name, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{t / 1e3:10.3f} ms {100 * t / allt:5.1f}%  {c:6d} x  {name}")
