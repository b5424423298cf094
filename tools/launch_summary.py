#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / mean / total."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    d = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi or not r[vi].replace(",", "").replace(".", "").isdigit():
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000.0 if r[ui] == "ns" else (v * 1000.0 if r[ui] == "ms" else v)
        d.setdefault(r[ki], []).append(v)
    tot = sum(sum(v) for v in d.values())
    for n, v in d.items():
        print("%-78s n=%3d mean=%9.1f us  total=%9.1f us (%4.1f%%)" % (n[:78], len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))


if __name__ == "__main__":
    main(sys.argv[1])
