#!/usr/bin/env python3
"""Per-SASS-instruction executed counts from an `ncu --page source --csv` export.
usage: ncu_src_summary.py file.csv [min_share_percent]   prints address offset, executed count (M), share, samples, SASS"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
ia, isrc, iex, ismp, ith = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Avg. Predicated-On Threads Executed")
data = [(int(r[ia], 16), r[isrc].strip(), int(r[iex]), int(r[ismp]), r[ith]) for r in rows[hi + 1:] if len(r) > iex and r[iex].isdigit()]
base = data[0][0]
tot = sum(d[2] for d in data); tots = sum(d[3] for d in data)
print("total executed %.1f M warp-instr, %d samples" % (tot / 1e6, tots))
for a, s, e, sm, th in data:
    print("%05x %8.2fM %5.2f%% smp %5.2f%% thr %5s  %s" % (a - base, e / 1e6, 100.0 * e / tot, 100.0 * sm / max(tots, 1), th, s))
