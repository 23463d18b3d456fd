#!/usr/bin/env python
"""Static opcode histogram of one kernel's SASS (cuobjdump), no GPU needed: python scripts/sass_static.py OBJ KERNEL_SUBSTRING"""
import collections, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
cur = None; c = collections.Counter(); n = 0
for line in out.splitlines():
    if "Function :" in line:
        cur = line.split("Function :")[1].strip(); continue
    if cur and sys.argv[2] in cur and "/*" in line and ";" in line:
        body = line.split("*/")[1].strip() if line.strip().startswith("/*") else line
        t = body.split()
        if not t: continue
        op = t[1] if t[0].startswith("@") else t[0]
        base = op.split(".")[0].rstrip(";")
        if base == "IMAD" and "RZ, RZ" in body: base = "MOV(IMAD)"
        c[base] += 1; n += 1
print(n, "instructions")
print(", ".join("%s %d" % kv for kv in c.most_common(24)))
