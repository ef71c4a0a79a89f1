// Microbenchmark: how fast can an SM push many SMALL shared->global bulk copies (cp.async.bulk.global.shared::cta)?
// Question behind it: a response's fixed-position prefix (status line + constant headers, ~208-304 bytes, 16-byte
// aligned in its slot) could leave as one TMA store per request issued by the request's own thread, instead of 13-19
// st.global.v4 per thread.  Variants:
//   0: one bulk store of S bytes per request, issued by the request's thread (128 threads/CTA, persistent tiles)
//   1: thread-per-request st.global.cs.v4 loop (what the Writer's long copy does today: ld.shared x4 + st.v4)
//   2: warp-cooperative coalesced copy: a warp writes 32 consecutive 16-byte windows per instruction
//   3: variant 0 but one elected lane per warp issues the 32 bulk stores of its warp (serialised issue)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_store_bench bulk_store_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int VARIANT>
__global__ void __launch_bounds__(128) k(uint8_t* out, uint32_t n, uint32_t slot, uint32_t S, uint32_t split) {
    __shared__ __align__(128) uint8_t tmpl[1024];
    for (uint32_t i = threadIdx.x; i < 1024; i += 128) tmpl[i] = (uint8_t)(i * 7 + 1);
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes of the template visible to the async proxy
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t tile = blockIdx.x; tile * 128 < n; tile += gridDim.x) {
        const uint32_t i = tile * 128 + threadIdx.x;
        if (VARIANT == 0) {
            if (i < n) {
                uint8_t* d = out + (size_t)i * slot;
                if (split && split < S) { bulk_s2g(d, tmpl, split); bulk_s2g(d + split + 48, tmpl + split + 48, S - split); }  // two pieces around a 48-byte hole
                else bulk_s2g(d, tmpl, S);
                bulk_commit();
            }
        } else if (VARIANT == 3) {
            const uint32_t w0 = tile * 128 + (threadIdx.x & ~31u);
            if (lane == 0) {
                for (uint32_t j = 0; j < 32 && w0 + j < n; j++) bulk_s2g(out + (size_t)(w0 + j) * slot, tmpl, S);
                bulk_commit();
            }
        } else if (VARIANT == 1) {
            if (i < n) {
                uint8_t* d = out + (size_t)i * slot;
                for (uint32_t o = 0; o < S; o += 16) __stcs((uint4*)(d + o), *(const uint4*)(tmpl + o));
            }
        } else {
            const uint32_t w0 = tile * 128 + (threadIdx.x & ~31u);
            const uint32_t wins = S / 16;  // windows per response
            const uint32_t total = 32 * wins;
            for (uint32_t g = lane; g < total; g += 32) {
                const uint32_t r = g / wins, j = g - r * wins;
                if (w0 + r < n) __stcs((uint4*)(out + (size_t)(w0 + r) * slot + j * 16), *(const uint4*)(tmpl + j * 16));
            }
        }
    }
    if (VARIANT == 0 || VARIANT == 3) bulk_wait0();
}

int main(int argc, char** argv) {
    const uint32_t n = 1u << 20, slot = argc > 1 ? atoi(argv[1]) : 528;
    uint8_t* out;
    CK(cudaMalloc(&out, (size_t)n * slot));
    CK(cudaMemset(out, 0, (size_t)n * slot));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const uint32_t sizes[] = {64, 208, 304, 512};
    for (int ctas = 4; ctas <= 16; ctas *= 2) {
        for (uint32_t S : sizes) {
            if (S > slot) continue;
            for (int v = 0; v < 5; v++) {
                const uint32_t split = v == 4 ? 208 : 0;
                if (v == 4 && S != 304) continue;
                float best = 1e9f;
                for (int rep = 0; rep < 5; rep++) {
                    CK(cudaEventRecord(e0));
                    const int grid = 148 * ctas;
                    if (v == 0 || v == 4) k<0><<<grid, 128>>>(out, n, slot, S, split);
                    else if (v == 1) k<1><<<grid, 128>>>(out, n, slot, S, 0);
                    else if (v == 2) k<2><<<grid, 128>>>(out, n, slot, S, 0);
                    else k<3><<<grid, 128>>>(out, n, slot, S, 0);
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (rep && ms < best) best = ms;
                }
                printf("ctas/SM=%2d S=%3u variant=%d%s: %.4f ms  %.1f GB/s  %.2f G stores/s\n", ctas, S, v == 4 ? 0 : v, v == 4 ? "(split 208+48)" : "", best,
                       (double)n * (v == 4 ? S - 48 : S) / best / 1e6, (double)n / best / 1e6);
            }
        }
    }
    // correctness spot check of variant 0
    CK(cudaMemset(out, 0, (size_t)n * slot));
    k<0><<<148 * 8, 128>>>(out, n, slot, 208, 0);
    CK(cudaDeviceSynchronize());
    std::vector<uint8_t> h(slot * 4);
    CK(cudaMemcpy(h.data(), out + (size_t)(n - 4) * slot, slot * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int r = 0; r < 4; r++) for (uint32_t b = 0; b < slot; b++) { uint8_t want = b < 208 ? (uint8_t)(b * 7 + 1) : 0; if (h[r * slot + b] != want) bad++; }
    printf("check variant 0: %s\n", bad ? "MISMATCH" : "ok");
    return 0;
}
