#!/usr/bin/env python
"""Stall samples per CUDA source line from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: python scripts/ncu_lines.py X_cs.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
fn = fp = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, ""])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fp = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        fn = r[1][:60]; continue
    if r[0] == "Line No":
        hdr = r; iS = r.index("# Samples"); iI = r.index("Instructions Executed"); continue
    if hdr and r[0].isdigit():
        key = (fn, fp, int(r[0]))
        try:
            agg[key][0] += int(r[iS] or 0); agg[key][1] += int(r[iI] or 0)
        except ValueError:
            pass
        agg[key][2] = r[1][:120]
for kern in sorted(set(k[0] for k in agg)):
    tot = sum(v[0] for k, v in agg.items() if k[0] == kern)
    print("==", kern, "samples", tot)
    for s, k, v in sorted([(v[0], k, v) for k, v in agg.items() if k[0] == kern], reverse=True)[:top]:
        print("  %5.1f%% %-18s:%4d  %s" % (100 * s / max(tot, 1), k[1], k[2], v[2].strip()))
