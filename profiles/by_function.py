#!/usr/bin/env python
"""Aggregates an ncu SASS source page by device function (serve_device.cuh) using nvdisasm line info.
usage: by_function.py src.csv kernel.sass [device_header]"""
import bisect, collections, csv, os, re, sys
src_csv, sass_file = sys.argv[1], sys.argv[2]
header = sys.argv[3] if len(sys.argv) > 3 else "gofr_b200/csrc/serve_device.cuh"
addr_line, cur, in_k = {}, None, False
for ln in open(sass_file, errors="replace"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if ".text." in ln:
        in_k = os.environ.get("KERNEL", "serve_kernel") in ln
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and in_k:
        addr_line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
funcs = []
for i, l in enumerate(open(header), 1):
    m = re.match(r"\s*(?:template <[^>]*>\s*)?GOFR_HD\s+(?:static\s+)?[\w:<>\*& ]+?\s+(\w+)\(", l)
    if m: funcs.append((i, m.group(1)))
    m = re.match(r"struct (\w+)", l)
    if m: funcs.append((i, "struct " + m.group(1)))
starts = [f[0] for f in funcs]
agg, smp = collections.Counter(), collections.Counter()
base = None; tot = tots = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    try: a = int(r[ci["Address"]], 16)
    except ValueError: continue
    if base is None: base = a
    k = addr_line.get(a - base, ("?", 0))
    ex = int(r[ci["Instructions Executed"]] or 0); s = int(r[ci["# Samples"]] or 0)
    if k[0] == header.split("/")[-1]:
        j = bisect.bisect_right(starts, k[1]) - 1
        name = funcs[j][1] if j >= 0 else "?"
    else:
        name = k[0]
    agg[name] += ex; smp[name] += s; tot += ex; tots += s
print(f"total warp-instructions {tot}, stall samples {tots}")
for n, v in agg.most_common(28):
    print("%-28s %10d %5.1f%%  samples %5.1f%%" % (n, v, 100 * v / tot, 100 * smp[n] / max(tots, 1)))
