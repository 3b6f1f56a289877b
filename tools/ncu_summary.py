#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total, mean, share."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = r["Kernel Name"].split("(")[0].replace("void ", "")
    v = float(r["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r["Metric Unit"]]
    agg.setdefault(k, []).append(v)
tot = sum(sum(v) for v in agg.values())
print("%-34s %5s %11s %10s %7s" % ("kernel", "n", "total ms", "mean ms", "share"))
for k, v in agg.items():
    print("%-34s %5d %11.3f %10.3f %6.1f%%" % (k[:34], len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))
print("%-34s %5s %11.3f" % ("all", "", tot))
