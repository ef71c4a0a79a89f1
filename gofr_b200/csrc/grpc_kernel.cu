// grpc_kernel.cu — the gRPC message path over batches of length-prefixed frames, sm_100a: the unary Hello of config 5
// (request frame → response frame) and the proto3 encoder / decoder for any flat message type (rows ↔ frames).
//
// Same execution model as serve_kernel.cu: persistent co-resident CTAs, one thread per frame, the tile's contiguous
// input byte range pulled into shared memory with one TMA bulk copy, exact output sizes scanned in the CTA and chained
// across CTAs by the decoupled look-back, responses packed in request order and written through the aligned staging
// Writer (16-byte st.global.cs.v4).  Per-frame logic: grpc_device.cuh.
#include <cuda_runtime.h>
#include <stdint.h>

#include "engine_internal.h"
#include "grpc_device.cuh"
#include "tile_common.cuh"

namespace gofr {

constexpr int GT = kServeThreads;
constexpr int GNW = GT / 32;
constexpr uint32_t kGrpcStage = 8 * 1024;  // bytes of input frames staged per tile (128 frames × ≤ 64 B)

struct GrpcShared {
    uint64_t bar;
    uint32_t warp_sum[GNW];
    unsigned long long tile_base;
    uint32_t stage[GOFR_STAGE_WORDS * GT];
    __align__(16) uint8_t in[kGrpcStage + 32];
};

// The tile loop shared by the two frame codecs (Hello request → response, row → proto3 message).
template <class Codec>
__device__ __forceinline__ void frame_tiles(const GrpcParams& p, const Codec& cd, GrpcShared& sh) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&sh.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity = 0;
    for (uint32_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const uint32_t i = tile * GT + tid;
        const bool valid = i < p.n;
        const uint32_t t0 = tile * GT, t1 = min(p.n, t0 + GT);
        const uint32_t lo = __ldg(p.in_off + t0) & ~15u, hi = (__ldg(p.in_off + t1) + 15u) & ~15u;
        const bool staged = hi > lo && hi - lo <= kGrpcStage;
        __syncthreads();  // previous tile's reads of sh.in are done
        if (staged && tid == 0) {
            mbar_expect_tx(&sh.bar, hi - lo);
            bulk_g2s(sh.in, p.in + lo, hi - lo, &sh.bar);
        }
        uint32_t fo = 0, fn = 0;
        if (valid) { fo = __ldg(p.in_off + i); fn = __ldg(p.in_off + i + 1) - fo; }
        const uint8_t* base = p.in;
        if (staged) {
            mbar_wait(&sh.bar, parity);
            parity ^= 1;
            base = launder_after_sync((const uint8_t*)sh.in) - lo;
        }
        typename Codec::R r = cd.none();
        if (valid) r = cd.parse(base + fo, fn, fo);

        uint32_t incl = r.out_len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += v;
        }
        if (lane == 31) sh.warp_sum[warp] = incl;
        __syncthreads();
        uint32_t warp_excl = 0, tile_total = 0;
#pragma unroll
        for (int w = 0; w < GNW; w++) {
            uint32_t s = sh.warp_sum[w];
            if ((uint32_t)w < warp) warp_excl += s;
            tile_total += s;
        }
        const uint32_t excl = warp_excl + incl - r.out_len;
        if (warp == 0) {
            unsigned long long b = lookback(p.tile_state, p.epoch, tile, tile_total, lane);
            if (lane == 0) sh.tile_base = b;
        }
        __syncthreads();
        const unsigned long long tile_base = sh.tile_base;
        const bool fits = tile_base + tile_total <= p.out_cap && tile_base + tile_total <= 0xFFFFFFFFull;
        if (!fits && tid == 0) atomicExch(p.overflow, 1u);
        if (valid) {
            p.out_off[i] = (uint32_t)(tile_base + excl);
            p.meta[i] = r.status;
            if (i == p.n - 1) p.out_off[p.n] = (uint32_t)(tile_base + excl + r.out_len);
        }
        if (fits && valid && r.out_len) cd.emit(base + fo, r, p.out + tile_base + excl, &sh.stage[tid]);
    }
}

struct HelloCodec {
    typedef HelloReq R;
    __device__ R none() const { return HelloReq{GOFR_GRPC_OK, 0, 0, 0}; }
    __device__ R parse(const uint8_t* f, uint32_t fn, uint32_t) const { return hello_parse(f, fn); }
    __device__ void emit(const uint8_t* f, const R& r, uint8_t* dst, uint32_t* col) const { hello_emit(f, r, dst, col); }
};

struct ProtoCodec {
    typedef ProtoMsg R;
    const ProtoSchema& S;
    __device__ R none() const { return ProtoMsg{GOFR_GRPC_OK, 0}; }
    __device__ R parse(const uint8_t* row, uint32_t rn, uint32_t off) const { return proto_size(S, row, rn, (off & 3u) == 0); }
    __device__ void emit(const uint8_t* row, const R& r, uint8_t* dst, uint32_t* col) const { proto_emit(S, row, r, dst, col); }
};

struct ProtoDecodeCodec {
    typedef ProtoRow R;
    const ProtoSchema& S;
    __device__ R none() const { ProtoRow r; r.status = GOFR_GRPC_OK; r.out_len = 0; return r; }
    __device__ R parse(const uint8_t* f, uint32_t fn, uint32_t) const { ProtoRow r; proto_decode_scan(S, f, fn, r); return r; }
    __device__ void emit(const uint8_t* f, const R& r, uint8_t* dst, uint32_t* col) const { proto_decode_emit(S, f, r, dst, col); }
};

__global__ void __launch_bounds__(GT, 12) grpc_hello_kernel(const GrpcParams p) {
    __shared__ __align__(16) GrpcShared sh;
    frame_tiles(p, HelloCodec{}, sh);
}

// rows → proto3 messages (gofr_proto_encode_device): the same pipeline, the row in place of the request frame
__global__ void __launch_bounds__(GT, 8) proto_encode_kernel(const GrpcParams p, const __grid_constant__ ProtoSchema S) {
    __shared__ __align__(16) GrpcShared sh;
    frame_tiles(p, ProtoCodec{S}, sh);
}

int launch_grpc_hello(const GrpcParams& p, int grid, void* stream) {
    grpc_hello_kernel<<<grid, GT, 0, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}

// frames → rows (gofr_proto_decode_device)
__global__ void __launch_bounds__(GT, 8) proto_decode_kernel(const GrpcParams p, const __grid_constant__ ProtoSchema S) {
    __shared__ __align__(16) GrpcShared sh;
    frame_tiles(p, ProtoDecodeCodec{S}, sh);
}

int launch_proto_decode(const GrpcParams& p, const ProtoSchema& S, int grid, void* stream) {
    proto_decode_kernel<<<grid, GT, 0, (cudaStream_t)stream>>>(p, S);
    return (int)cudaGetLastError();
}

int proto_decode_max_grid(int device) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, proto_decode_kernel, GT, 0) != cudaSuccess) return -1;
    return nb * prop.multiProcessorCount;
}

int launch_proto_encode(const GrpcParams& p, const ProtoSchema& S, int grid, void* stream) {
    proto_encode_kernel<<<grid, GT, 0, (cudaStream_t)stream>>>(p, S);
    return (int)cudaGetLastError();
}

int proto_max_grid(int device) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, proto_encode_kernel, GT, 0) != cudaSuccess) return -1;
    return nb * prop.multiProcessorCount;
}

int grpc_max_grid(int device) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, grpc_hello_kernel, GT, 0) != cudaSuccess) return -1;
    return nb * prop.multiProcessorCount;
}

}  // namespace gofr
