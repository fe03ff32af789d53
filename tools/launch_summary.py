#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (markdown)."""
import collections, csv, sys

rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
agg = collections.OrderedDict()
for r in rows:
    nm = r["Kernel Name"]
    v = float(r["Metric Value"])
    us = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r["Metric Unit"], v)
    a = agg.setdefault((nm, r["Grid Size"], r["Block Size"]), [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print("# %s\n" % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("%d launches, %.1f ms of GPU time in total (cold-cache, serialised by ncu: compare shares).\n" % (len(rows), tot / 1e3))
print("| kernel | grid | block | launches | total us | avg us | share |\n|---|---|---|---|---|---|---|")
for (nm, g, b), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %s | %s | %d | %.1f | %.1f | %.1f%% |" % (nm[:120].replace("|", "/"), g, b, c, t, t / c, 100 * t / tot))
