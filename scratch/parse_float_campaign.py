#!/usr/bin/env python
"""Bind's float64 literal parser (bind_device.cuh bd_parse_float: exact cases, Eisel-Lemire, sure over- / underflows) on the CPU
against glibc's correctly rounded strtod, on generated literals (tests/emu emu_parse_float_check: random mantissas and
exponents, the 15 .. 25-digit texts of random doubles, integers at and around the exact half-way points between doubles,
19-digit mantissas with a tail of dropped digits).   python scratch/parse_float_campaign.py [millions per mode and thread] [threads]"""
import ctypes as C
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.emu import emu  # noqa: E402

emu._build()
L = emu.lib()
L.emu_parse_float_check.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_uint64), C.c_char_p, C.c_uint32]
M = int(sys.argv[1]) if len(sys.argv) > 1 else 10
T = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 4)
res = {}


def work(k):
    tot = [0, 0, 0, 0]
    first = b""
    for mode in range(4):
        out = (C.c_uint64 * 4)()
        bad = C.create_string_buffer(600)
        L.emu_parse_float_check(1000 * k + mode + int(time.time()) % 100000 * 7919, M * 1_000_000, mode, out, bad, 600)
        tot = [a + int(b) for a, b in zip(tot, out)]
        first = first or bad.value
    res[k] = (tot, first)


t0 = time.time()
th = [threading.Thread(target=work, args=(k,)) for k in range(T)]
[t.start() for t in th]
[t.join() for t in th]
tot = [sum(v[0][i] for v in res.values()) for i in range(4)]
print(f"{sum(tot[:2])} literals: {tot[0]} decided on the device code ({tot[3]} of them as overflows), {tot[1]} left to the host, "
      f"{tot[2]} MISMATCHES against strtod" + ("" if not tot[2] else " e.g. " + repr([v[1] for v in res.values() if v[1]][:3])) +
      f" ({time.time() - t0:.0f} s, {T} threads)")
