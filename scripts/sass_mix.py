#!/usr/bin/env python
"""Opcode mix per kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass` (executed warp instructions)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
kern = None; mix = None; out = []
for r in rows:
    if r and r[0] == "Kernel Name":
        kern = r[1][:80]; mix = collections.Counter(); samp = collections.Counter(); out.append((kern, mix, samp)); hdr = None; continue
    if r and r[0] == "Address":
        hdr = {h: i for i, h in enumerate(r)}; continue
    if kern and hdr and len(r) > 5:
        src = r[hdr["Source"]].strip()
        toks = src.split()
        if not toks: continue
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.rstrip(";")
        base = op.split(".")[0]
        n = int(r[hdr["Instructions Executed"]] or 0)
        mix[base] += n
        samp[base] += int(r[hdr["# Samples"]] or 0)
for kern, mix, samp in out:
    tot = sum(mix.values()); ts = sum(samp.values())
    print("==", kern, "total warp inst", tot)
    for op, n in mix.most_common(top):
        print("  %-10s %12d  %5.1f%%   samples %5.1f%%" % (op, n, 100.0 * n / tot, 100.0 * samp[op] / max(ts, 1)))
