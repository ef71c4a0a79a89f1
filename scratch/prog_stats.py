"""Response programs of a route table as the serve kernel walks them: ops, appended segments, and the number of RUNS the
program collapses to when generated scalars are folded into their neighbouring literals.

    python scratch/prog_stats.py            # configs 1-4 of BASELINE.json
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from gofr_b200 import synth  # noqa: E402
from gofr_b200.table import Table  # noqa: E402
from tests.emu import emu  # noqa: E402

NAMES = {0: "LIT", 1: "HEXID", 3: "CLEN", 4: "I64", 5: "I32", 6: "BOOL", 7: "STR", 8: "PARAM", 9: "LOCATION", 10: "ERRMSG", 11: "BLOB",
         12: "KEY", 13: "BSTR"}
VERBATIM = {7, 8, 9, 10, 11, 13}      # bytes copied from request / cold memory: a run of their own
OP = np.dtype([("code", "u1"), ("arg", "u1"), ("flags", "u1"), ("kind", "u1"), ("len", "<u4"), ("off", "<u4"), ("aux", "<u4")])


def stats(image: bytes):
    L = emu.lib()
    L.emu_prog_ops.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    img = np.frombuffer(image, dtype=np.uint8).copy()
    rows = {}
    p = 0
    while True:
        ops = np.zeros(256, dtype=OP)
        hf, bf, sc = C.c_uint32(), C.c_uint32(), C.c_uint32()
        n = L.emu_prog_ops(img.ctypes.data, p, ops.ctypes.data, 256, C.byref(hf), C.byref(bf), C.byref(sc))
        if n < 0:
            break
        ops = ops[:n]
        appends = sum(1 + (1 if (o["code"] not in (0, 12, 11) and o["len"]) else 0) for o in ops)
        # Runs: pieces in program order — literal (its length), generated scalar (upper bound of its
        # length), verbatim copy.  Literals of >= 32 bytes stay template runs; shorter literals and scalars that touch
        # each other are written into ONE scratch run by phase A; every verbatim copy is a run of its own.
        GEN = {1: 32, 3: 10, 4: 20, 5: 11, 6: 5, 12: 0}
        pieces = []
        for o in ops:
            c = int(o["code"])
            if c == 0:
                pieces.append(("lit", int(o["len"])))
                continue
            if c == 12:
                pieces.append(("lit", int(o["len"]) + 1))        # key with its comma
                continue
            if o["len"] and c != 11:
                pieces.append(("lit", int(o["len"])))            # literal prefix folded into the value op
            pieces.append(("copy", 0) if c in VERBATIM else ("gen", GEN.get(c, 8)))
        runs, scratch_bytes, in_scratch = 0, 0, False
        for kind, ln in pieces:
            if kind == "copy" or (kind == "lit" and ln >= 32):
                runs += 1
                in_scratch = False
            else:
                if not in_scratch:
                    runs += 1
                in_scratch = True
                scratch_bytes += ln
        key = (int(sc.value), n, appends, runs, scratch_bytes, hf.value + bf.value, " ".join(NAMES.get(int(o["code"]), "?") for o in ops))
        rows[key] = rows.get(key, 0) + 1
        p += 1
    return rows


if __name__ == "__main__":
    for name, spec in (("config1", synth.config1_spec()), ("config2", synth.config2_spec()), ("config3", synth.config3_spec()),
                       ("config4", synth.config4_spec())):
        print(f"== {name}")
        for (sc, n, app, runs, scr, fixed, seq), cnt in sorted(stats(Table(spec).serialize()).items()):
            print(f"  class {sc:2d}  x{cnt:<4d} ops {n:2d}  appends {app:2d}  runs {runs:2d}  scratch <= {scr:3d} B  fixed bytes {fixed:4d}   {seq}")
