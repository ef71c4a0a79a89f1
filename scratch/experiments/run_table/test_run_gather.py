"""CPU check of scratch/experiments/run_table/run_gather.cuh against plain concatenation (run: python this_file)."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = r'''
#include "run_gather.cuh"
extern "C" void windows(const uint8_t* pool, const uint32_t* offs, const uint32_t* lens, uint32_t n, uint8_t* out, uint32_t n_windows) {
    RunTable T; T.n = n; uint32_t e = 0;
    for (uint32_t k = 0; k < n; k++) { e += lens[k]; T.end[k] = e; T.src[k] = pool + offs[k]; }
    for (uint32_t w = 0; w < n_windows; w++) { uint32_t v[4]; rg_window(T, w * 16, v); memcpy(out + 16 * w, v, 16); }
}
'''
with tempfile.TemporaryDirectory() as d:
    open(os.path.join(d, "t.cpp"), "w").write(SRC)
    so = os.path.join(d, "t.so")
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-fsanitize=address,undefined", "-static-libasan", "-I" + HERE, os.path.join(d, "t.cpp"), "-o", so])
    env_preload = subprocess.check_output(["g++", "-print-file-name=libasan.so"]).decode().strip()
    code = f'''
import ctypes as C, numpy as np
L = C.CDLL({so!r})
L.windows.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
rng = np.random.default_rng(3)
for trial in range(3000):
    n = int(rng.integers(1, 13))
    lens = rng.integers(0, 80, n).astype(np.uint32)
    if trial % 7 == 0: lens[:] = rng.integers(0, 4, n)
    pool = rng.integers(0, 256, int(lens.sum()) + 8 * n + 64, dtype=np.uint8)
    offs = np.zeros(n, dtype=np.uint32); pos = int(rng.integers(0, 4))
    for k in range(n):
        offs[k] = pos; pos += int(lens[k]) + int(rng.integers(0, 5))
    want = b"".join(pool[int(o):int(o) + int(l)].tobytes() for o, l in zip(offs, lens))
    nw = (len(want) + 15) // 16 + 1
    out = np.zeros(16 * nw, dtype=np.uint8)
    L.windows(pool.ctypes.data, offs.ctypes.data, lens.ctypes.data, n, out.ctypes.data, nw)
    assert out.tobytes() == want + bytes(16 * nw - len(want)), (trial, n, lens)
print("run_gather: 3000 random run tables ok")
'''
    subprocess.check_call(["python", "-c", code], env=dict(os.environ, LD_PRELOAD=env_preload, ASAN_OPTIONS="detect_leaks=0"))
