#!/usr/bin/env python3
"""Aggregate an ncu launch-list CSV (gpu__time_duration + dram bytes) per kernel.
usage: python tools/launch_summary.py gpurun_out/launches_c3.csv [first_n]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
h = rows[0]
ki, mi, vi, ii = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('ID')
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[ii], {'k': r[ki].split('(')[0].replace('void ', '')[:44]})[r[mi]] = float(r[vi].replace(',', ''))
agg = collections.OrderedDict()
tot = sum(d['gpu__time_duration.sum'] for d in per.values())
for d in per.values():
    a = agg.setdefault(d['k'], [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d['gpu__time_duration.sum']
    a[2] += d.get('dram__bytes_read.sum', 0)
    a[3] += d.get('dram__bytes_write.sum', 0)
print(f"total {tot / 1e6:.3f} ms over {len(per)} launches")
print(f"{'kernel':46s} {'n':>4s} {'ms':>9s} {'share':>6s} {'dram rd GB':>11s} {'dram wr GB':>11s}")
for k, a in agg.items():
    print(f"{k:46s} {a[0]:4d} {a[1] / 1e6:9.3f} {a[1] / tot * 100:5.1f}% {a[2] / 1e9:11.3f} {a[3] / 1e9:11.3f}")
for i, d in enumerate(per.values()):
    if i < first:
        print("   ", d['k'], round(d['gpu__time_duration.sum'] / 1e6, 3), 'ms', round(d.get('dram__bytes_read.sum', 0) / 1e9, 3),
              round(d.get('dram__bytes_write.sum', 0) / 1e9, 3))
