#!/usr/bin/env python
"""Per source line: samples of one stall reason (ncu source page).  usage: stall_by_line.py src.csv kernel.sass stall_long_sb [top]"""
import os, sys, re, csv, collections
src_csv, sass_file, col = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
addr_line = {}; cur = None; in_k = False; text = {}
for ln in open(sass_file, errors="replace"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if ".text." in ln: in_k = os.environ.get("KERNEL", "serve_kernel") in ln
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and in_k: addr_line[int(m.group(1), 16)] = cur; text[int(m.group(1), 16)] = m.group(2)
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
agg = collections.Counter(); ex = collections.defaultdict(list); base = None; tot = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    try: a = int(r[ci["Address"]], 16)
    except ValueError: continue
    if base is None: base = a
    k = addr_line.get(a - base, ("?", 0))
    s = int(r[ci[col]] or 0)
    agg[k] += s; tot += s
    if s: ex[k].append((s, text.get(a - base, "?")))
print(col, "total", tot)
for k, v in agg.most_common(top):
    best = max(ex[k])[1] if ex[k] else ""
    print("%-28s %6d %5.1f%%   %s" % (f"{k[0]}:{k[1]}", v, 100 * v / max(tot, 1), best[:70]))
