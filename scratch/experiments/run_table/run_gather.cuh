// run_gather.cuh — EXPERIMENT for DESIGN.md §9a (not part of libgofr_b200.so): phase B of the run-table serve kernel.
//
// A response is described by a short table of runs (source address in shared-memory space, length); lane w of a warp
// produces bytes [16w, 16w+16) of the response: find the run that holds byte 16w, gather 16 bytes that may continue
// into the following runs, return them as one uint4 for a coalesced st.global.v4.
// __host__ __device__ so that scratch/experiments/run_table/test_run_gather.py can check it against a plain
// concatenation on the CPU; `nvcc -cubin` + cuobjdump gives the static instruction count of the window routine.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RG_HD __host__ __device__ __forceinline__
#else
#define RG_HD inline
#endif

constexpr int kMaxRuns = 12;

struct RunTable {
    uint32_t n;                    // runs in use
    uint32_t end[kMaxRuns];        // exclusive prefix ends: run k covers [end[k-1], end[k])
    const uint8_t* src[kMaxRuns];  // first byte of run k (any alignment; readable up to the next aligned word after its end)
};

// unaligned 32-bit load as two aligned loads + funnel shift (sources are only guaranteed readable to the end of the
// aligned word that holds their last byte)
RG_HD uint32_t rg_load32(const uint8_t* p) {
    const uintptr_t a = (uintptr_t)p;
    const uint32_t* q = (const uint32_t*)(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    const uint32_t lo = q[0];
    if (sh == 0) return lo;
    const uint32_t hi = q[1];
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return lo >> sh | hi << (32 - sh);
#endif
}

// bytes [pos, pos+16) of the concatenation of the runs; bytes past the end of the response are zero
RG_HD void rg_window(const RunTable& T, uint32_t pos, uint32_t out[4]) {
    // branch-free search: k = number of runs that end at or before pos
    uint32_t k = 0;
#pragma unroll
    for (int j = 0; j < kMaxRuns; j++) k += (j < (int)T.n && T.end[j] <= pos) ? 1u : 0u;
    out[0] = out[1] = out[2] = out[3] = 0;
    if (k >= T.n) return;
    uint32_t start = k ? T.end[k - 1] : 0u;
    if (T.end[k] - pos >= 16) {  // interior window: the common case, 16 bytes from one run
        const uint8_t* p = T.src[k] + (pos - start);
        out[0] = rg_load32(p); out[1] = rg_load32(p + 4); out[2] = rg_load32(p + 8); out[3] = rg_load32(p + 12);
        return;
    }
    // boundary window: byte by byte across runs (at most a handful of windows per response)
    uint8_t b[16];
    for (uint32_t i = 0; i < 16; i++) {
        const uint32_t at = pos + i;
        while (k < T.n && T.end[k] <= at) { start = T.end[k]; k++; }
        b[i] = k < T.n ? T.src[k][at - start] : 0;
    }
    memcpy(out, b, 16);
}
