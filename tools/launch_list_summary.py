#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file x.csv): per kernel / grid the number of
launches, mean duration and share of the summed kernel time.   usage: launch_list_summary.py x.csv > x_summary.txt"""
import collections, csv, re, sys


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    iK, iB, iG, iV, iU = hdr.index("Kernel Name"), hdr.index("Block Size"), hdr.index("Grid Size"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[iK]).replace("void unnamed>::", "").strip()
        us = float(r[iV].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iU], 1e-3)
        agg.setdefault((name, r[iG], r[iB]), []).append(us)
    total = sum(sum(v) for v in agg.values())
    fam = collections.Counter()
    for (name, grid, block), v in agg.items():
        print("%-34s grid %-18s block %-14s launches %3d  mean %10.1f us  share %5.1f%%" % (name, grid, block, len(v), sum(v) / len(v), 100 * sum(v) / total))
        fam[name.split("<")[0]] += sum(v)
    print(", ".join("%s total share %.1f%%" % (k, 100 * v / total) for k, v in fam.most_common()))


if __name__ == "__main__":
    main(sys.argv[1])
