#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import sys


def main(path, skip=0):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr, agg, n_seen = None, collections.defaultdict(lambda: [0, 0.0]), 0
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        n_seen += 1
        if n_seen <= skip:
            continue
        u = d["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "nsecond": 1e-3}.get(u, 1.0)
        name = d["Kernel Name"].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:34s} n={v[0]:5d} total={v[1]:10.1f}us avg={v[1] / v[0]:8.1f}us {100 * v[1] / tot:5.1f}%")
    print(f"{'TOTAL':34s} {tot:10.1f}us")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
