#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name.
usage: tools/ncu_summary.py gpurun_out/<tag>_launches.csv > profiles/<tag>_launches_summary.txt"""
import collections
import csv
import sys


def main(path):
    agg = collections.OrderedDict()
    hdr = None
    n_rows = 0
    for r in csv.reader(open(path, errors="replace")):
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if not hdr or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        name = d["Kernel Name"]
        name = name.replace("capdec::<unnamed>::", "").replace("void ", "")
        key = (name[:90], d["Grid Size"], d["Block Size"])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
        n_rows += 1
    tot = sum(v[1] for v in agg.values())
    print("# %s: %d launches, %.1f us total device time (ncu: serialised, cold cache -- compare SHARES)"
          % (path, n_rows, tot))
    print("%7s %11s %6s %9s  %-22s %s" % ("launches", "total_us", "share", "avg_us", "grid/block", "kernel"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%7d %11.1f %5.1f%% %9.2f  %-22s %s" % (v[0], v[1], 100 * v[1] / tot, v[1] / v[0],
                                                       k[1] + k[2], k[0]))


if __name__ == "__main__":
    main(sys.argv[1])
