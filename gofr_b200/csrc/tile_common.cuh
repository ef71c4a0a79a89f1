// tile_common.cuh — machinery shared by the tile kernels (serve_kernel.cu, grpc_kernel.cu): TMA bulk-copy and mbarrier
// wrappers, and the decoupled look-back that chains per-tile output sizes into packed offsets in a single pass.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "engine_internal.h"

namespace gofr {

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
// global → shared bulk copy (TMA, 1-D): dst/src 16-byte aligned, bytes % 16 == 0
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_global, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst_smem)),
                 "l"(src_global), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
// The copy sources read staged bytes through NON-volatile `ld.shared` inline asm (the compiler may schedule them freely
// among themselves).  Such an asm has no memory operand, so neither a barrier nor a "memory" clobber orders it: only a
// data dependence does.  Every pointer to freshly staged shared memory (the TMA-filled tile after the mbarrier wait, the
// table copy after its __syncthreads) is therefore passed through this empty asm right after the synchronisation: all
// addresses derived from the result depend on it, so no load of the staged bytes can be hoisted above the wait.
// (Round 2: a build with a different store instruction hoisted literal-pool loads above the table barrier in the first
// tile of a CTA — zeros in the output, visible only on the GPU.)
template <typename T>
__device__ __forceinline__ T* launder_after_sync(T* p) {
    asm volatile("" : "+l"(p) : : "memory");
    return p;
}
__device__ __forceinline__ unsigned long long ld_state(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// decoupled look-back over tile totals.  state = epoch(24) | flag(2) | value(38); flag 1 = tile total, 2 = inclusive
// prefix.  Words written by earlier launches carry an older epoch and read as "not ready"; the host clears every state
// buffer when the epoch wraps (engine.cu, next_epoch).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kValBits = 38;
constexpr unsigned long long kValMask = (1ull << kValBits) - 1;
__device__ __forceinline__ unsigned long long pack_state(uint32_t epoch, uint32_t flag, unsigned long long v) {
    return ((unsigned long long)(epoch & kEpochMask) << (kValBits + 2)) | ((unsigned long long)flag << kValBits) | (v & kValMask);
}

// Called by warp 0.  Returns the exclusive prefix of `tile` (valid in every lane).
__device__ __forceinline__ unsigned long long lookback(unsigned long long* state, uint32_t epoch, uint32_t tile,
                                                       unsigned long long total, uint32_t lane) {
    if (lane == 0) st_state(&state[tile], pack_state(epoch, tile == 0 ? 2 : 1, total));
    unsigned long long excl = 0;
    if (tile > 0) {
        long long base = (long long)tile - 1;
        for (;;) {
            long long t = base - (long long)lane;
            unsigned long long s = t >= 0 ? ld_state(&state[t]) : pack_state(epoch, 2, 0);
            uint32_t flag = (uint32_t)(s >> kValBits) & 3u;
            bool ready = (uint32_t)(s >> (kValBits + 2)) == (epoch & kEpochMask) && flag != 0;
            uint32_t not_ready = __ballot_sync(0xFFFFFFFFu, !ready);
            uint32_t is_p = __ballot_sync(0xFFFFFFFFu, ready && flag == 2);
            uint32_t upto = is_p ? (uint32_t)__ffs((int)is_p) - 1 : 31u;  // lanes 0..upto contribute
            uint32_t need = upto == 31 ? 0xFFFFFFFFu : ((1u << (upto + 1)) - 1);
            if (not_ready & need) { __nanosleep(40); continue; }
            unsigned long long v = (lane <= upto) ? (s & kValMask) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
            excl += v;
            if (is_p) break;
            base -= 32;
        }
        if (lane == 0) st_state(&state[tile], pack_state(epoch, 2, excl + total));
    }
    return excl;
}


}  // namespace gofr
