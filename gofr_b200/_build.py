"""Builds libgofr_b200.so (CUDA, sm_100a) in-tree with nvcc.  No torch involvement: the library is a plain C ABI."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgofr_b200.so")
SOURCES = ["serve_kernel.cu", "grpc_kernel.cu", "reqlog_kernel.cu", "route_kernel.cu", "http_kernel.cu", "egress_kernel.cu", "engine.cu", "table_build.cpp", "frontend.cpp"]
HEADERS = ["serve_device.cuh", "bind_device.cuh", "grpc_device.cuh", "reqlog_device.cuh", "http_device.cuh", "tile_common.cuh", "table_format.h", "engine_internal.h",
           "../../include/gofr_b200.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-Wall", "--expt-relaxed-constexpr"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    m = os.path.getmtime(LIB)
    files = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(f) and os.path.getmtime(f) > m for f in files)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, f) for f in SOURCES if os.path.exists(os.path.join(CSRC, f))]
    extra = os.environ.get("GOFR_EXTRA_NVCC", "").split()  # experiments only, e.g. -DGOFR_SERVE_T=96
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", LIB] + srcs + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libgofr_b200.so")
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
