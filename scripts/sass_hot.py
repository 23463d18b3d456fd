#!/usr/bin/env python
"""Hot-loop view of `ncu --page source --csv --print-source sass`: per kernel, the instructions executed at least
`frac` x the most-executed count, with executed counts (thousands of warp instructions) and stall samples.
usage: python scripts/sass_hot.py X_sass.csv KERNEL_SUBSTRING [--all | --nonfp | --summary]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]
mode = sys.argv[3] if len(sys.argv) > 3 else "--summary"
kern = None; hdr = None; data = {}
for r in rows:
    if r and r[0] == "Kernel Name":
        kern = r[1][:90]
        if kern in data: kern = None
        else: data[kern] = []
        hdr = None; continue
    if r and r[0] == "Address":
        hdr = {h: i for i, h in enumerate(r)}; continue
    if kern and hdr and len(r) > 5:
        data[kern].append((r[hdr["Source"]].strip(), int(r[hdr["Instructions Executed"]] or 0), int(r[hdr["# Samples"]] or 0)))
FP = ("DFMA", "DMUL", "DADD", "DSETP", "MUFU", "FSEL")
for k, v in data.items():
    if want not in k: continue
    mx = max(x[1] for x in v); tot = sum(x[1] for x in v); ts = sum(x[2] for x in v)
    step = sorted(x[1] for x in v if x[1] > 0.5 * mx)[0]
    print("==", k, "| warp inst", tot, "| per warp-step %.1f" % (tot / step), "| samples", ts)
    c = collections.Counter(); cs = collections.Counter()
    for i, (src, n, sm) in enumerate(v):
        t = src.split(); op = (t[1] if t[0].startswith("@") else t[0]).rstrip(";")
        base = op.split(".")[0]
        if base == "IMAD" and "RZ, RZ" in src: base = "MOV(IMAD)"
        c[base] += n; cs[base] += sm
        hot = n >= 0.5 * mx
        if mode == "--all" and hot: print(i, src[:100], n // 1000, sm)
        if mode == "--nonfp" and hot and base not in FP: print(i, src[:100], n // 1000, sm)
    if mode == "--summary":
        for op, n in c.most_common(28):
            print("  %-10s %7.1f per warp-step  %5.1f%%   samples %5.1f%%" % (op, n / step, 100.0 * n / tot, 100.0 * cs[op] / max(ts, 1)))
