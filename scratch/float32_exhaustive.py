#!/usr/bin/env python
"""Every finite non-zero float32 (sign aside: 2^31 bit patterns) through the device code's float32 text on the CPU
(tests/emu emu_float32_check): parses back to the same float32, is the shortest such decimal, and its digits are the
correctly rounded ones of that length.  ~5 minutes on 8 cores.   python scratch/float32_exhaustive.py [threads]"""
import ctypes as C
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.emu import emu  # noqa: E402

emu._build()
L = emu.lib()
L.emu_float32_check.restype = C.c_uint64
L.emu_float32_check.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint32)]
T = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 4)
total, res = 1 << 31, {}


def work(k):
    lo, hi = total * k // T, total * (k + 1) // T
    bad = C.c_uint32(0)
    res[k] = (L.emu_float32_check(lo, 1, hi - lo, C.byref(bad)), bad.value)


t0 = time.time()
th = [threading.Thread(target=work, args=(k,)) for k in range(T)]
[t.start() for t in th]
[t.join() for t in th]
fails = sum(v[0] for v in res.values())
print(f"float32 text, all {total} bit patterns with the sign bit clear: {fails} failures"
      + ("" if not fails else " first " + ", ".join(hex(v[1]) for v in res.values() if v[0])) + f" ({time.time() - t0:.0f} s, {T} threads)")
