#!/usr/bin/env python
"""SASS instruction count per kernel of a built library (cuobjdump -sass).  usage: sass_sizes.py [lib.so]"""
import re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "gofr_b200/libgofr_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, n, res = None, 0, []
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        if name: res.append((name, n))
        name, n = m.group(1), 0
    elif re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        n += 1
if name: res.append((name, n))
for nm, k in res:
    short = subprocess.run(["c++filt", nm], capture_output=True, text=True).stdout.strip().split("(")[0]
    print(f"{short:40s} {k:7d} instructions  {k * 16 / 1024:7.1f} KB")
