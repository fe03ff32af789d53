#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into a markdown table:
per kernel the launches, total and average duration and the share of the captured GPU time.
    python tools/launch_list_summary.py gpurun_out/launches.csv "title" > profiles/rNN_launches.md
Also handles multi-metric CSVs: --metrics mode prints per launch the named metrics."""
import csv
import sys
from collections import OrderedDict


def rows(path):
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    return list(csv.DictReader(lines))


def main():
    path, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
    rs = rows(path)
    if len(sys.argv) > 3 and sys.argv[3] == "--metrics":
        per = OrderedDict()
        for r in rs:
            per.setdefault(r["ID"], {"kernel": r["Kernel Name"]})[r["Metric Name"]] = (r["Metric Value"], r["Metric Unit"])
        print("# %s\n" % title)
        names = [k for k in next(iter(per.values())) if k != "kernel"]
        print("| launch | kernel | " + " | ".join(names) + " |")
        print("|---|---|" + "---|" * len(names))
        for i, d in per.items():
            print("| %s | `%s` | " % (i, d["kernel"][:60]) + " | ".join("%s %s" % d[n] for n in names) + " |")
        return
    agg = OrderedDict()
    tot = 0.0
    for r in rs:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        key = (r["Kernel Name"], r["Grid Size"], r["Block Size"])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += us
        tot += us
    print("# %s\n" % title)
    print("%d launches, %.1f ms of GPU time in total (cold-cache, serialised by ncu: compare shares).\n" % (
        sum(a[0] for a in agg.values()), tot / 1e3))
    print("| kernel | grid | block | launches | total us | avg us | share |\n|---|---|---|---|---|---|---|")
    for (k, g, b), (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %s | %s | %d | %.1f | %.1f | %.1f%% |" % (k[:120], g, b, c, us, us / c, 100 * us / tot))


if __name__ == "__main__":
    main()
