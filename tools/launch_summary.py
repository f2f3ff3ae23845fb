#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
time and share.  usage: launch_summary.py launches.csv [first_id last_id]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, vi, ii = h.index("Kernel Name"), h.index("Metric Value"), h.index("ID")
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 2:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", "")); i = int(r[ii])
    except ValueError:
        continue
    if lo <= i <= hi:
        n = r[ki][:90]
        agg[n][0] += 1; agg[n][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot / 1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{t / 1e6:9.3f} ms {c:5d} {100 * t / tot:5.1f}%  {n}")
