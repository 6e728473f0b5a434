#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name the
number of launches, mean duration and share.  summarize_launches.py launches.csv [last_n]"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, gi, vi, ui = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value"), hdr.index("Metric Unit")
data = rows[1:]
if len(sys.argv) > 2:
    data = data[-int(sys.argv[2]):]
agg = OrderedDict()
for r in data:
    v = float(r[vi].replace(",", ""))
    v = v / 1e6 if r[ui] in ("ns", "nsecond") else v / 1e3 if r[ui] in ("us", "usecond") else v
    name = r[ki].split("(")[0]
    k = (name, r[gi])
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for (name, grid), (n, t) in agg.items():
    print("%-44s grid=%-22s n=%4d mean %9.4f ms  share %5.1f%%" % (name[:44], grid, n, t / n, 100 * t / tot))
print("total %.3f ms over %d launches" % (tot, len(data)))
