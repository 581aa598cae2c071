#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean, share."""
import csv
import sys
from collections import defaultdict


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = None
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        unit = d["Metric Unit"]
        us = v * (1e-3 if unit.startswith("n") else 1.0 if unit.startswith("u") else 1e3)
        k = d["Kernel Name"].split("(")[0][:60]
        agg[k][0] += 1
        agg[k][1] += us
    total = sum(t for _, t in agg.values())
    print(f"{'launches':>8} {'mean us':>10} {'total us':>12} {'share':>7}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{n:8d} {t / n:10.1f} {t:12.1f} {100 * t / total:6.1f}%  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
