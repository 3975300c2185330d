#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` log: launches, total time and share per kernel."""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
name_i, val_i, unit_i = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[val_i].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[unit_i].replace("second", "s").replace("nsecond", "ns"), None)
    if scale is None:
        scale = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(r[unit_i], 1e-3)
    name = r[name_i].split("(")[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v * scale
total = sum(a[1] for a in agg.values())
print("# kernel | launches | total us | share | mean us")
for k, (n, t) in agg.items():
    print("%s | %d | %.1f | %.1f%% | %.1f" % (k, n, t, 100 * t / total, t / n))
