#!/usr/bin/env python
"""Build-time check of a -DGOFR_STORE256 library: every evict-first global store in the SASS must be a 256-bit one.
CUDA 12.9's ptxas lowers `st.global.cs.v8.b32` to a SCALAR `STG.E.EF` (first word only, the other seven source registers
dropped) in some clones of an out-of-line device function (seen: gofr::flush_out as inlined into serve_kernel, fine in
serve_slots_kernel of the same translation unit).  In such a build no other store uses .cs, so any `STG.E.EF` that is not
`.256` is that miscompilation.  usage: check_sector_stores.py lib.so   (exit status 1 when a bad store is found)"""
import re, subprocess, sys
lib = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, bad, good = None, [], 0
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        fn = m.group(1)
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if not m or "STG" not in m.group(2) or ".EF" not in m.group(2):
        continue
    if ".256" in m.group(2):
        good += 1
    else:
        bad.append((fn, m.group(1), m.group(2).strip()))
print(f"{lib}: {good} 256-bit evict-first stores, {len(bad)} narrower ones")
for b in bad[:10]:
    print("  BAD", b)
sys.exit(1 if bad or not good else 0)
