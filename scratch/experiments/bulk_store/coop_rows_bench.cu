// Microbenchmark 2: coalescing granularity.  32 responses per warp, each lane owns one response (slot = 528 bytes).  The
// warp writes the responses in row groups: R consecutive 16-byte chunks of one response are written by R adjacent lanes
// of one st.global.v4 instruction (32/R responses per instruction).  R = 1 is the thread-per-request pattern of the
// round-1 kernel (every lane its own 128-byte line), R = 8 writes whole 128-byte lines.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o coop_rows_bench coop_rows_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// aligned variant: the row groups follow the 128-byte lines of the GLOBAL address space (a response starts at chunk
// (33 * i) % 8 of a line when slot = 528), so every instruction writes whole aligned lines except at response edges
template <int R>
__global__ void __launch_bounds__(128) ka(uint8_t* out, uint32_t n, uint32_t slot, uint32_t chunks) {
    __shared__ __align__(16) uint4 tmpl[64];
    if (threadIdx.x < 64) tmpl[threadIdx.x] = make_uint4(threadIdx.x, 1, 2, 3);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t tile = blockIdx.x; tile * 128 < n; tile += gridDim.x) {
        const uint32_t w0 = tile * 128 + (threadIdx.x & ~31u);
        for (uint32_t g = 0; g * R < chunks + R; g++) {
#pragma unroll
            for (uint32_t i = 0; i < R; i++) {
                const uint32_t r = i * (32 / R) + lane / R;
                const uint32_t first = (uint32_t)(((size_t)(w0 + r) * slot / 16) % R);  // chunk phase of the response start
                const int c = (int)(g * R + lane % R) - (int)first;                     // group g = global line group
                if (w0 + r < n && c >= 0 && c < (int)chunks) __stcs((uint4*)(out + (size_t)(w0 + r) * slot + c * 16), tmpl[c & 63]);
            }
        }
    }
}

template <int R>
__global__ void __launch_bounds__(128) k(uint8_t* out, uint32_t n, uint32_t slot, uint32_t chunks) {
    __shared__ __align__(16) uint4 tmpl[64];
    if (threadIdx.x < 64) tmpl[threadIdx.x] = make_uint4(threadIdx.x, 1, 2, 3);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t tile = blockIdx.x; tile * 128 < n; tile += gridDim.x) {
        const uint32_t w0 = tile * 128 + (threadIdx.x & ~31u);
        // chunk groups of R: group g covers chunks [g*R, g*R+R) of every response of the warp
        for (uint32_t g = 0; g * R < chunks; g++) {
#pragma unroll
            for (uint32_t i = 0; i < R; i++) {  // R instructions write the group for all 32 responses
                const uint32_t r = i * (32 / R) + lane / R, c = g * R + lane % R;
                if (w0 + r < n && c < chunks) __stcs((uint4*)(out + (size_t)(w0 + r) * slot + c * 16), tmpl[c & 63]);
            }
        }
    }
}

int main() {
    const uint32_t n = 1u << 20, slot = 528, chunks = 33;
    uint8_t* out;
    CK(cudaMalloc(&out, (size_t)n * slot));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int ctas = 4; ctas <= 8; ctas *= 2)
        for (int R = 1; R <= 32; R *= 2) {
            float best = 1e9f;
            for (int rep = 0; rep < 5; rep++) {
                CK(cudaEventRecord(e0));
                const int grid = 148 * ctas;
                switch (R) {
                    case 1: k<1><<<grid, 128>>>(out, n, slot, chunks); break;
                    case 2: k<2><<<grid, 128>>>(out, n, slot, chunks); break;
                    case 4: k<4><<<grid, 128>>>(out, n, slot, chunks); break;
                    case 8: k<8><<<grid, 128>>>(out, n, slot, chunks); break;
                    case 16: k<16><<<grid, 128>>>(out, n, slot, chunks); break;
                    default: k<32><<<grid, 128>>>(out, n, slot, chunks); break;
                }
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (rep && ms < best) best = ms;
            }
            printf("ctas/SM=%d R=%2d (%3d contiguous bytes per response per instruction): %.4f ms  %.1f GB/s\n", ctas, R, R * 16, best, (double)n * chunks * 16 / best / 1e6);
            if (R == 4 || R == 8 || R == 16) {
                best = 1e9f;
                for (int rep = 0; rep < 5; rep++) {
                    CK(cudaEventRecord(e0));
                    const int grid = 148 * ctas;
                    if (R == 4) ka<4><<<grid, 128>>>(out, n, slot, chunks); else if (R == 8) ka<8><<<grid, 128>>>(out, n, slot, chunks); else ka<16><<<grid, 128>>>(out, n, slot, chunks);
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (rep && ms < best) best = ms;
                }
                printf("ctas/SM=%d R=%2d ALIGNED to %3d-byte global blocks: %.4f ms  %.1f GB/s\n", ctas, R, R * 16, best, (double)n * chunks * 16 / best / 1e6);
            }
        }
    return 0;
}
