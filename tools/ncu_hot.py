#!/usr/bin/env python3
"""Summarise an ncu `--page source --csv` dump: top stall-sample instructions with their dominant
stall reason.  usage: ncu -i X.ncu-rep --page source --csv > f.csv; python tools/ncu_hot.py f.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = None
cur = None
out = {}
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = r[1][:70] + "#%d" % len(out)
        out[cur] = []
        continue
    if r and r[0] == "Address":
        h = r
        continue
    if h is None or len(r) < len(h):
        continue
    out[cur].append(r)
for k, rs in out.items():
    si = h.index("# Samples")
    total = sum(int(r[si] or 0) for r in rs)
    print("==", k, "total samples", total)
    stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    agg = {h[i]: sum(int(r[i] or 0) for r in rs) for i in stall_cols}
    print("  stall totals:", ", ".join(f"{a}={b * 100 // max(total, 1)}%" for a, b in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    for idx, r in sorted(enumerate(rs), key=lambda x: -int(x[1][si] or 0))[:top]:
        dom = max(stall_cols, key=lambda i: int(r[i] or 0))
        print(f"  {int(r[si]) * 100.0 / max(total, 1):5.1f}%  #{idx:4d} {r[1].strip()[:70]:70s} {h[dom]}")
