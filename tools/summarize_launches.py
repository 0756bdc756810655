#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count / average / total / share."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        agg.setdefault(r[ki][:70], []).append(float(r[vi].replace(',', '')))
    except ValueError:
        pass
tot = sum(sum(v) for v in agg.values())
print("%-72s %5s %10s %11s %6s" % ("kernel", "n", "avg_us", "total_us", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-72s %5d %10.1f %11.1f %5.1f%%" % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3, 100 * sum(v) / tot))
