// Microbenchmark 3: can ONE lane write a whole aligned 32-byte sector (st.global.v8.b32, new with sm_100) and does that
// remove the partial-sector penalty of the thread-per-request pattern?  Each thread owns one 528-byte slot; it writes its
// 33 chunks as: [one 16-byte store if the slot starts mid-sector] + 32-byte aligned v8 stores + [a trailing 16-byte store].
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o v8_bench v8_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ void st_v8(void* p, uint4 a, uint4 b) {
    asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

template <int MODE>  // 0: 16-byte stores; 1: aligned 32-byte stores
__global__ void __launch_bounds__(128) k(uint8_t* out, uint32_t n, uint32_t slot, uint32_t chunks) {
    __shared__ __align__(16) uint4 tmpl[64];
    if (threadIdx.x < 64) tmpl[threadIdx.x] = make_uint4(threadIdx.x, 1, 2, 3);
    __syncthreads();
    for (uint32_t tile = blockIdx.x; tile * 128 < n; tile += gridDim.x) {
        const uint32_t i = tile * 128 + threadIdx.x;
        if (i >= n) continue;
        uint8_t* d = out + (size_t)i * slot;
        uint32_t c = 0;
        if (MODE == 1) {
            if ((uintptr_t)d & 16) { __stcs((uint4*)d, tmpl[0]); c = 1; }
            for (; c + 2 <= chunks; c += 2) st_v8(d + c * 16, tmpl[c & 63], tmpl[(c + 1) & 63]);
        }
        for (; c < chunks; c++) __stcs((uint4*)(d + c * 16), tmpl[c & 63]);
    }
}

int main() {
    const uint32_t n = 1u << 20, slot = 528, chunks = 33;
    uint8_t* out;
    CK(cudaMalloc(&out, (size_t)n * slot));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int ctas = 4; ctas <= 8; ctas *= 2)
        for (int mode = 0; mode < 2; mode++) {
            float best = 1e9f;
            for (int rep = 0; rep < 5; rep++) {
                CK(cudaEventRecord(e0));
                if (mode) k<1><<<148 * ctas, 128>>>(out, n, slot, chunks); else k<0><<<148 * ctas, 128>>>(out, n, slot, chunks);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (rep && ms < best) best = ms;
            }
            printf("ctas/SM=%d %s: %.4f ms  %.1f GB/s\n", ctas, mode ? "thread-per-request, aligned 32-byte st.global.v8" : "thread-per-request, 16-byte st.global.v4", best, (double)n * chunks * 16 / best / 1e6);
        }
    return 0;
}
