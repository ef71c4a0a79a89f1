// serve_body.cuh — the fused, persistent serve kernel for sm_100a (body shared by the two kernels: serve_kernel.cu
// instantiates the packed layout, serve_slots_kernel.cu the slot layout; separate translation units so that each can be
// built with the store shape that compiles correctly for it, see gofr_b200/_build.py).
//
// One launch replaces, for a whole batch, what the reference does per request on a goroutine:
// mux.Router.ServeHTTP → Tracer/Logging/CORS → handler.ServeHTTP → Responder.Respond → net/http framing
// (pkg/gofr/http/router.go:14, middleware/{tracer,logger,cors}.go, pkg/gofr/handler.go:32-36, pkg/gofr/http/responder.go:19-41).
//
// Execution model (HBM-bound integer/byte work; no tensor cores):
//   * grid = co-resident CTAs only (SMs × occupancy); CTA b walks tiles b, b+grid, … of 128 requests, one thread per
//     request;
//   * the tile's descriptors and trace ids are read with coalesced 16-byte loads; its contiguous arena byte range is
//     pulled into shared memory with ONE TMA bulk copy (cp.async.bulk.shared::cluster.global + mbarrier) when it
//     fits, so all per-request byte walking hits shared memory, not HBM;
//   * responses are packed back-to-back in request order: sizes are scanned inside the CTA and chained across CTAs
//     with a decoupled look-back (single pass — inputs are read from HBM exactly once);
//   * each thread streams its response through a funnel-shift word writer (serve_device.cuh) whose completed words
//     collect in a conflict-free shared-memory ring and leave for HBM as aligned 16-byte st.global.cs.v4 chunks;
//     L2 merges the two halves of each sector, so HBM sees full-sector writes.  No output tile lives in shared
//     memory, which keeps 5 CTAs (20 warps) resident per SM for this latency-bound byte work.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "engine_internal.h"
#include "serve_device.cuh"
#include "tile_common.cuh"

namespace gofr {

// ---------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------
constexpr int T = kServeT;
constexpr int NW = T / 32;

constexpr uint32_t kDeferred = 0xFFFFFFFFu;  // length-column mark of a request left for the general pass (never a length)

struct TileShared {
    uint64_t bar;  // mbarrier for the arena bulk load
    uint32_t warp_sum[NW];
    uint32_t warp_lo[NW], warp_hi[NW];
    uint32_t warp_cls[NW];  // slot layout: program shape class of each warp's lane 0
    unsigned long long tile_base;
    uint32_t in_lo, in_hi;
    uint32_t defer_n;         // slot layout: requests waiting for the general pass; their indices (at most 2 T: one tile's
                              // worth is added between two checks) live in the request staging area, which is idle once
                              // the tile loop is over — as static shared memory the list cost every instance its fifth CTA
    uint32_t ring[GOFR_STAGE_WORDS * T];  // word-major staging buffer of the Writer (serve_device.cuh)
};

// Slot layout, general pass.  The tile loop of serve_slots_kernel runs only the fast path (size_fast / emit_fast); a request
// it cannot take — a program outside the fast path, a value that needs escaping or decoding, HEAD, an unstaged tile — is
// marked in the length column and served here after the loop, by the general interpreter, with ALL threads of the CTA
// working on such requests at once (128 at a time).  Slots make this legal: a response depends on nothing but its own request.
// Mixed traffic used to make every warp run both emitters, or — when the slow requests were sorted into one warp — made
// the other three wait for it at the next barrier (barrier stall 5.5 cycles per issue on the 64-route workload).  The
// request bytes are read from HBM here (the tile that staged them is gone); the code is out of line so that the tile loop
// keeps its registers and its instruction-cache footprint.
static __device__ __noinline__ void general_pass(const ServeParams* pp, const TableView tv, const BatchRefs br, TileShared* sh,
                                          const uint32_t* defer, uint32_t first, uint32_t count) {
    const ServeParams& p = *pp;
    const uint32_t tid = threadIdx.x;
    if (tid >= count) return;
    const uint32_t r = defer[first + tid];
    const uint4 d = __ldg((const uint4*)p.desc + r);
    ReqCtx c;
    c.set(p.arena, d.x, d.y & 0xFFFFu, d.y >> 16, d.z, d.w & 0xFFu, (d.w >> 8) & 0xFFu, false, r);
    route_request(tv, br, c);
    size_routed_general(tv, br, c);
    p.out_off[r] = c.total_len;
    p.meta[r] = request_status(tv, c) | (c.route << 16);
    if (c.total_len <= p.slot_bytes && c.total_len)
        emit_request_general<true>(tv, br, c, p.out + (size_t)r * p.slot_bytes, &sh->ring[tid]);
}

template <bool SLOTS>
__device__ __forceinline__ void serve_body(const ServeParams& p) {
    // (p is a __grid_constant__ kernel parameter: its address can be handed to the out-of-line general pass)
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(16) TileShared sh;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* tbl = smem;
    uint8_t* in_stage = smem + ((p.hot_bytes + 127u) & ~127u);

    // table → shared memory (once per CTA), then the batch's Date is patched into the literal pool
    {
        const uint4* src = (const uint4*)p.image;
        uint4* dst = (uint4*)tbl;
        for (uint32_t i = tid; i < p.hot_bytes / 16; i += T) dst[i] = src[i];
        if (tid == 0) {
            mbar_init(&sh.bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            sh.defer_n = 0;
        }
    }
    __syncthreads();
    patch_dates(tbl, (const uint8_t*)p.date, tid, T);
    __syncthreads();
    tbl = launder_after_sync(tbl);
    TableView tv;
    tv.bind(tbl, p.image);
    BatchRefs br;
    br.ids = (const uint8_t*)p.ids; br.bind_scratch = p.bind_scratch; br.bind_row_words = p.bind_row_words;

    uint32_t parity = 0;
    const unsigned long long chain0 = p.chain_pos ? (*p.chain_pos & 15ull) : 0ull;

    // Static round-robin tile assignment over the co-resident grid: tile t only ever waits on tiles < t, all of which
    // belong to resident CTAs that process their tiles in increasing order, so the look-back cannot deadlock.
    // (A dynamic ticket counter was measured slower: the atomic's round trip sits on every tile's critical path.)
    for (uint32_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const uint32_t i = tile * T + tid;
        const bool valid = i < p.n;
        uint4 d = make_uint4(0, 0, 0, 0);
        if (valid) d = __ldg((const uint4*)p.desc + i);
        // the trace ids are read only when the response is written: have the line in L1 by then (one lane per 128-byte line)
        if (valid && (lane & 7u) == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"((const uint4*)p.ids + i));
        const uint32_t arena_off = d.x, path_len = d.y & 0xFFFFu, query_len = d.y >> 16, data_len = d.z;
        const uint32_t data_off = (arena_off + path_len + query_len + 3u) & ~3u;
        const uint32_t end = data_off + data_len;

        // ---- the tile's arena byte range ----
        uint32_t lo = valid ? arena_off : 0xFFFFFFFFu, hi = valid ? end : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, o));
        }
        if (lane == 0) { sh.warp_lo[warp] = lo; sh.warp_hi[warp] = hi; }
        __syncthreads();  // also: every thread is done reading the previous tile's in_stage
        if (tid == 0) {
            uint32_t l = sh.warp_lo[0], h = sh.warp_hi[0];
#pragma unroll
            for (int w = 1; w < NW; w++) { l = min(l, sh.warp_lo[w]); h = max(h, sh.warp_hi[w]); }
            l &= ~15u;
            h = (h + 15u) & ~15u;
            sh.in_lo = l;
            sh.in_hi = h;
            if (h > l && h - l <= p.in_cap) {
                mbar_expect_tx(&sh.bar, h - l);
                bulk_g2s(in_stage, p.arena + l, h - l, &sh.bar);
            }
#if defined(GOFR_PREFETCH_NEXT)
            // the tile this CTA takes next: ask for its bytes in L2 now (a hint — the range is guessed from its first and
            // last descriptor, which is exact whenever requests lie in the arena in index order)
            const uint32_t nt = tile + gridDim.x;
            if (nt < p.n_tiles) {
                const uint32_t i0 = nt * T, i1 = min(p.n, i0 + T) - 1;
                const uint4 d0 = __ldg((const uint4*)p.desc + i0), d1 = __ldg((const uint4*)p.desc + i1);
                const uint32_t plo = d0.x & ~15u;
                const uint32_t phi = (((d1.x + (d1.y & 0xFFFFu) + (d1.y >> 16) + 3u) & ~3u) + d1.z + 15u) & ~15u;
                if (phi > plo && phi - plo <= 2u * p.in_cap)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.arena + plo), "r"(phi - plo) : "memory");
            }
#endif
        }
#if defined(GOFR_PREFETCH_NEXT)
        {   // next tile's descriptors and trace ids: one lane per 128-byte line
            const uint32_t ni = (tile + gridDim.x) * T + tid;
            if (ni < p.n && (lane & 7u) == 0) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"((const uint4*)p.desc + ni));
                asm volatile("prefetch.global.L2 [%0];" ::"l"((const uint4*)p.ids + ni));
            }
        }
#endif
        __syncthreads();
        const uint32_t in_lo = sh.in_lo, in_hi = sh.in_hi;
        const bool in_staged = in_hi > in_lo && in_hi - in_lo <= p.in_cap;
        const uint8_t* abase = p.arena;
        if (in_staged) {
            mbar_wait(&sh.bar, parity);
            parity ^= 1;
            abase = launder_after_sync((const uint8_t*)in_stage) - in_lo;  // abase + arena_off lands in the staged copy
        }

        // ---- stage 1: route ----
        ReqCtx c;
        c.set(abase, arena_off, path_len, query_len, data_len, d.w & 0xFFu, (d.w >> 8) & 0xFFu, in_staged, valid ? i : 0xFFFFFFFFu);
        if (valid) route_request(tv, br, c);

        if (SLOTS) {
            // Slot layout: response i owns out + i * slot_bytes.  No scan, no look-back, no dependence between tiles —
            // and every response starts 16-byte aligned.  Nothing ties a request to a particular thread either, so a
            // tile with mixed traffic is first regrouped by program shape: lanes of a warp then walk the same op
            // sequence instead of serialising over every shape present (the interpreter's only divergence).
            uint32_t cls = 31u;  // idle lanes sort last
            if (valid) {
                cls = 0u;
                if (c.prog != 0xFFFF) {
                    const ProgRec& PR = tv.progs()[c.prog];
                    // requests that are known to take the general interpreter (a program outside the fast path, a query
                    // value that needs decoding, HEAD) form one class of their own: they end up in the same warp instead
                    // of making every warp of the tile run both emitters
                    const bool general = !(PR.flags & PF_FAST) || (c.pv_flags & 2u) || c.method() == GOFR_M_HEAD || !c.staged();
                    cls = general ? 30u : PR.shape_class;
                }
            }
            const uint32_t cls0 = __shfl_sync(0xFFFFFFFFu, cls, 0);
            if (lane == 0) sh.warp_cls[warp] = cls0;
            bool mixed = __syncthreads_or(cls != cls0);
#pragma unroll
            for (int w = 1; w < NW; w++) mixed |= sh.warp_cls[w] != sh.warp_cls[0];
            if (mixed) {
                uint32_t* hist = sh.ring;  // the staging ring is idle until the first response is written
                if (tid < 32) hist[tid] = 0;
                __syncthreads();
                const uint32_t rank = atomicAdd(&hist[cls], 1u);
                __syncthreads();
                if (warp == 0) {
                    const uint32_t v = hist[lane];
                    uint32_t incl = v;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                        if (lane >= (uint32_t)o) incl += u;
                    }
                    hist[lane] = incl - v;
                }
                __syncthreads();
                const uint32_t pos = hist[cls] + rank;
                __syncthreads();
                uint32_t* x = sh.ring + pos;  // word-major exchange record of the request now owned by thread `pos`
                x[0 * T] = (uint32_t)(c.path - abase); x[1 * T] = c.path_len | c.query_len << 16; x[2 * T] = c.data_len;
                x[3 * T] = c.data_off; x[4 * T] = c.mflags; x[5 * T] = c.index; x[6 * T] = c.prog | c.route << 16;
                x[7 * T] = c.pv_off; x[8 * T] = c.pv_len; x[9 * T] = c.pv_flags; x[10 * T] = c.def_off; x[11 * T] = c.def_len;
                __syncthreads();
                const uint32_t* y = sh.ring + tid;
                c.path = abase + y[0 * T]; c.path_len = y[1 * T] & 0xFFFFu; c.query_len = y[1 * T] >> 16; c.data_len = y[2 * T];
                c.data_off = y[3 * T]; c.mflags = y[4 * T]; c.index = y[5 * T]; c.prog = y[6 * T] & 0xFFFFu; c.route = y[6 * T] >> 16;
                c.pv_off = y[7 * T]; c.pv_len = y[8 * T]; c.pv_flags = y[9 * T]; c.def_off = y[10 * T]; c.def_len = y[11 * T];
                __syncthreads();  // the ring is free again before any Writer stages into it
            }
            const uint32_t r = c.index;  // the request this thread serves now
            if (r != 0xFFFFFFFFu) {
                if (c.prog == 0xFFFFu) {  // GOFR_H_HOST: nothing to emit, status 0 = pending on the host
                    p.out_off[r] = 0;
                    p.meta[r] = c.route << 16;
                } else if (size_fast(tv, br, c)) {
                    p.out_off[r] = c.total_len;  // the length column; > slot_bytes tells the host the slot was too small
                    p.meta[r] = request_status(tv, c) | (c.route << 16);
                    if (c.total_len <= p.slot_bytes && c.total_len) emit_fast(tv, br, c, p.out + (size_t)r * p.slot_bytes, &sh.ring[tid]);
                } else {
                    p.out_off[r] = kDeferred;  // left for the general pass after the tile loop
                }
            }
            continue;
        }
        if (valid) size_routed(tv, br, c);
        // ---- block scan of response sizes ----
        uint32_t incl = c.total_len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += v;
        }
        if (lane == 31) sh.warp_sum[warp] = incl;
        __syncthreads();
        uint32_t warp_excl = 0, tile_total = 0;
#pragma unroll
        for (int w = 0; w < NW; w++) {
            uint32_t s = sh.warp_sum[w];
            if ((uint32_t)w < warp) warp_excl += s;
            tile_total += s;
        }
        const uint32_t excl = warp_excl + incl - c.total_len;

        // ---- chain tiles (warp 0) ----
        if (warp == 0) {
            unsigned long long base = (p.debug_flags & 1u) ? (unsigned long long)tile * tile_total
                                                           : lookback(p.tile_state, p.epoch, tile, tile_total, lane);
            if (lane == 0) sh.tile_base = base;
        }
        __syncthreads();
        // host-batch path: the chunk starts at the same offset mod 16 as its destination in the caller's buffer
        const unsigned long long tile_base = sh.tile_base + chain0;
        const bool fits = tile_base + tile_total <= p.out_cap && tile_base + tile_total <= 0xFFFFFFFFull;
        if (!fits && tid == 0) atomicExch(p.overflow, 1u);
        if (valid) {
            p.out_off[i] = (uint32_t)(tile_base + excl);
            p.meta[i] = request_status(tv, c) | (c.route << 16);
            if (i == p.n - 1) p.out_off[p.n] = (uint32_t)(tile_base + excl + c.total_len);
        }

        // ---- stage 3: emit straight to HBM in 16-byte chunks ----
        if (fits && valid && c.total_len) emit_request(tv, br, c, p.out + tile_base + excl, &sh.ring[tid]);
        // the next iteration's first __syncthreads orders this tile's shared-memory reads before any overwrite
    }
    if (SLOTS) {
        // General pass: the requests the tile loop marked in the length column.  Every CTA rescans the columns of its own
        // tiles (one coalesced 4-byte load per request; the marks were stored by this CTA, so a barrier makes them
        // visible — read past L1, which may hold nothing newer than the launch), gathers the marked indices and serves
        // them 128 at a time.  Nothing of this is in the tile loop: no list, no atomics, no call site.
        __syncthreads();
        uint32_t* const defer = (uint32_t*)in_stage;  // >= 2 T words: the engine never stages less than 16 bytes per request
        for (uint32_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const uint32_t i = tile * T + tid;
            const bool marked = i < p.n && __ldcg(p.out_off + i) == kDeferred;
            if (__syncthreads_or(marked)) {
                if (marked) defer[atomicAdd(&sh.defer_n, 1u)] = i;
                __syncthreads();
                const uint32_t dn = sh.defer_n;
                if (dn >= (uint32_t)T) {
                    general_pass(&p, tv, br, &sh, defer, dn - T, T);
                    __syncthreads();
                    if (tid == 0) sh.defer_n = dn - T;
                    __syncthreads();
                }
            }
        }
        const uint32_t dn = sh.defer_n;
        if (dn) general_pass(&p, tv, br, &sh, defer, 0, dn);
    }
}

}  // namespace gofr
