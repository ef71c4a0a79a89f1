// egress_kernel.cu — device-driven egress of the host-batch path.
//
// The packed size of a chunk is only known on the device.  Instead of stalling the host once per chunk to learn it
// (sync → exact-size cudaMemcpy), the chunk's bytes are streamed to the caller's pinned buffers by a small kernel:
//   advance_kernel  (1 thread, compute stream, after the serve kernel): chunk total → ChunkInfo, bump the batch-wide
//                   packed position `chain_pos` that the next chunk's serve kernel starts from;
//   egress_kernel   (egress stream): 16-byte vector copy HBM → pinned host memory (the serve kernel already placed
//                   the chunk at the same offset mod 16 as its destination), plus the offset column rebased to the
//                   batch and the meta column.
// The host enqueues every chunk without blocking and synchronises once at the end.
#include <cuda_runtime.h>
#include <stdint.h>

#include "engine_internal.h"

namespace gofr {

__global__ void advance_kernel(unsigned long long* chain_pos, const uint32_t* d_off, uint32_t n, uint32_t* d_overflow,
                               ChunkInfo* info, unsigned long long host_cap) {
    const unsigned long long base = *chain_pos;
    const uint32_t base0 = (uint32_t)(base & 15ull);
    const uint32_t total = d_off[n] - base0;
    const bool ovf = *d_overflow != 0 || base + total > host_cap || base + total > 0xFFFFFFFFull;
    info->host_base = base;
    info->base0 = base0;
    info->total = ovf ? 0u : total;
    info->overflow = ovf ? 1u : 0u;
    if (*d_overflow) *d_overflow = 0;
    if (!ovf) *chain_pos = base + total;
}

__global__ void __launch_bounds__(256) egress_kernel(const ChunkInfo* info, const uint8_t* d_out, const uint32_t* d_off,
                                                     const uint32_t* d_meta, uint32_t n, uint8_t* h_out, uint32_t* h_off,
                                                     uint32_t* h_meta, volatile unsigned long long* h_status) {
    const ChunkInfo ci = *info;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    // columns: offsets rebased to the whole batch, status/route copied
    const uint32_t rebase = (uint32_t)ci.host_base - ci.base0;
    for (uint32_t i = tid; i < n; i += nthreads) {
        h_off[i] = ci.overflow ? (uint32_t)ci.host_base : d_off[i] + rebase;
        h_meta[i] = d_meta[i];
    }
    if (ci.overflow) {
        if (tid == 0) h_status[1] = 1ull;
        return;
    }
    const uint8_t* src = d_out + ci.base0;
    uint8_t* dst = h_out + ci.host_base;  // same alignment mod 16 as src
    const uint32_t total = ci.total;
    uint32_t head = (16u - ci.base0) & 15u;
    if (head > total) head = total;
    if (tid < head) dst[tid] = src[tid];
    const uint32_t nvec = (total - head) >> 4;
    const uint4* s4 = (const uint4*)(src + head);
    uint4* d4 = (uint4*)(dst + head);
    for (uint32_t v = tid; v < nvec; v += nthreads) d4[v] = __ldcs(s4 + v);
    const uint32_t tail0 = head + (nvec << 4);
    if (tid < total - tail0) dst[tail0 + tid] = src[tail0 + tid];
}

int launch_advance(unsigned long long* chain_pos, const uint32_t* d_off, uint32_t n, uint32_t* d_overflow, ChunkInfo* info,
                   unsigned long long host_cap, void* stream) {
    advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(chain_pos, d_off, n, d_overflow, info, host_cap);
    return (int)cudaGetLastError();
}

int launch_egress(const ChunkInfo* info, const uint8_t* d_out, const uint32_t* d_off, const uint32_t* d_meta, uint32_t n,
                  uint8_t* h_out, uint32_t* h_off, uint32_t* h_meta, unsigned long long* h_status, int grid, void* stream) {
    egress_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(info, d_out, d_off, d_meta, n, h_out, h_off, h_meta, h_status);
    return (int)cudaGetLastError();
}

}  // namespace gofr
