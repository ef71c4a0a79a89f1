#!/usr/bin/env python
"""Joins an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) with nvdisasm --print-line-info output of the
same cubin, and aggregates executed instructions / stall samples per CUDA source line.

usage: sass_by_line.py src.csv kernel.sass [top_n]
"""
import os, csv
import collections
import re
import sys

src_csv, sass_file = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40

# ---- address → (file, line) from nvdisasm ----
addr_line = {}
cur = None
in_kernel = False
for ln in open(sass_file, errors="replace"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if ".text." in ln:
        in_kernel = os.environ.get("KERNEL", "serve_kernel") in ln
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and in_kernel:
        addr_line[int(m.group(1), 16)] = cur

rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
base = None
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        a = int(r[ci["Address"]], 16)
    except ValueError:
        continue
    if base is None:
        base = a
    key = addr_line.get(a - base, ("?", 0))
    ex = int(r[ci["Instructions Executed"]] or 0)
    smp = int(r[ci["# Samples"]] or 0)
    thr = int(r[ci["Thread Instructions Executed"]] or 0)
    for t, v in zip((agg[key], tot), ((ex, smp, thr),) * 2):
        t[0] += v[0]; t[1] += v[1]; t[2] += v[2]
print(f"total warp-instructions {tot[0]}  samples {tot[1]}  thread-instructions {tot[2]}")
print("%-28s %12s %7s %9s %7s" % ("file:line", "warp-inst", "%", "samples", "%"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-28s %12d %6.1f%% %9d %6.1f%%" % (f"{k[0]}:{k[1]}", v[0], 100.0 * v[0] / max(tot[0], 1), v[1], 100.0 * v[1] / max(tot[1], 1)))
