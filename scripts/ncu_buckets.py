#!/usr/bin/env python
"""Stall samples of one kernel of `ncu -i X.ncu-rep --page source --csv --print-source sass`, bucketed by windows of W
SASS instructions (program regions), with the leading stall reasons of each window.
usage: python scripts/ncu_buckets.py X_sass.csv <kernel substring> [W]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want, W = sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 64
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
for b in blocks:
    if want not in b["name"] or not b["rows"]:
        continue
    h = b["hdr"]; iS = h.index("# Samples"); iI = h.index("Instructions Executed")
    st = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    tot = sum(int(r[iS] or 0) for r in b["rows"]); toti = sum(int(r[iI] or 0) for r in b["rows"])
    print(b["name"][:70], "| SASS instrs", len(b["rows"]), "samples", tot)
    for s in range(0, len(b["rows"]), W):
        seg = b["rows"][s:s + W]
        ss = sum(int(r[iS] or 0) for r in seg); ie = sum(int(r[iI] or 0) for r in seg)
        if ss / max(tot, 1) > 0.004:
            reasons = {h[i]: sum(int(r[i] or 0) for r in seg) for i in st}
            top = sorted(reasons.items(), key=lambda x: -x[1])[:4]
            ops = {}
            for r in seg:
                o = r[1].split()[0] if r[1].split() else "?"
                if o.startswith("@"):
                    o = r[1].split()[1]
                ops[o.split(".")[0]] = ops.get(o.split(".")[0], 0) + 1
            print("%5d-%5d %5.1f%% samples %5.1f%% exec | %s | %s" % (
                s, s + W, 100 * ss / tot, 100 * ie / toti, " ".join("%s=%.0f%%" % (k[6:], 100 * v / max(ss, 1)) for k, v in top),
                " ".join("%s:%d" % kv for kv in sorted(ops.items(), key=lambda x: -x[1])[:5])))
