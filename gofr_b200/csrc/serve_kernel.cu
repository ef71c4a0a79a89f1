// serve_kernel.cu — the fused, persistent serve kernel for sm_100a.
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

struct TileShared {
    uint64_t bar;  // mbarrier for the arena bulk load
    uint32_t warp_sum[NW];
    uint32_t warp_lo[NW], warp_hi[NW];
    uint32_t warp_cls[NW];  // slot layout: program shape class of each warp's lane 0
    unsigned long long tile_base;
    uint32_t in_lo, in_hi;
    uint32_t ring[GOFR_STAGE_WORDS * T];  // word-major staging buffer of the Writer (serve_device.cuh)
};

template <bool SLOTS>
__device__ __forceinline__ void serve_body(const ServeParams& p) {
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
        }
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
            if (valid) cls = c.prog == 0xFFFF ? 0u : tv.progs()[c.prog].shape_class;
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
                size_routed<true>(tv, br, c);
                p.out_off[r] = c.total_len;  // the length column; > slot_bytes tells the host the slot was too small
                p.meta[r] = request_status(tv, c) | (c.route << 16);
                if (c.total_len <= p.slot_bytes && c.total_len)
                    emit_request<true>(tv, br, c, p.out + (size_t)r * p.slot_bytes, &sh.ring[tid]);
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
}

__global__ void __launch_bounds__(T, kServeCtas) serve_kernel(const ServeParams p) { serve_body<false>(p); }
__global__ void __launch_bounds__(T, kServeCtas) serve_slots_kernel(const ServeParams p) { serve_body<true>(p); }

uint32_t serve_smem_bytes(uint32_t hot_bytes, uint32_t in_cap) {
    return ((hot_bytes + 127u) & ~127u) + in_cap + 64;
}

int serve_max_grid(uint32_t smem_bytes, int device, int* blocks_per_sm) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(serve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(serve_slots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) return -1;
    int nb = 0, nb2 = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, serve_kernel, T, smem_bytes) != cudaSuccess) return -1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, serve_slots_kernel, T, smem_bytes) != cudaSuccess) return -1;
    if (nb2 < nb) nb = nb2;
    if (blocks_per_sm) *blocks_per_sm = nb;
    return nb * prop.multiProcessorCount;
}

int launch_serve(const ServeParams& p, int grid, uint32_t smem_bytes, void* stream) {
    if (p.slot_bytes) serve_slots_kernel<<<grid, T, smem_bytes, (cudaStream_t)stream>>>(p);
    else serve_kernel<<<grid, T, smem_bytes, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}

}  // namespace gofr
