#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step)."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith('==')))
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    agg.setdefault(r[ki].split('(')[0], []).append(v)
tot = sum(sum(v) for v in agg.values())
print('| kernel | launches | mean us | total us | share |')
print('|---|---|---|---|---|')
for k, v in agg.items():
    print('| %s | %d | %.1f | %.1f | %.1f%% |' % (k[-60:], len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3, 100 * sum(v) / tot))
