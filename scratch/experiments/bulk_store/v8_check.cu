// Correctness probe for 256-bit global stores (st.global.v8.b32 -> STG.E.256) when several lanes of ONE instruction hit
// sectors of the same 128-byte line, with full and partial warps, with and without the .cs hint.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
__device__ __forceinline__ void st_v8_cs(void* p, uint32_t v) {
    asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v), "r"(v + 1), "r"(v + 2), "r"(v + 3), "r"(v + 4), "r"(v + 5), "r"(v + 6), "r"(v + 7) : "memory");
}
__device__ __forceinline__ void st_v8(void* p, uint32_t v) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v), "r"(v + 1), "r"(v + 2), "r"(v + 3), "r"(v + 4), "r"(v + 5), "r"(v + 6), "r"(v + 7) : "memory");
}
// mode bit0: .cs ; mask: lanes that store; stride: bytes between the sectors of consecutive lanes
__global__ void k(uint8_t* out, uint32_t mask, uint32_t stride, int cs) {
    const uint32_t lane = threadIdx.x & 31, g = blockIdx.x * blockDim.x + threadIdx.x;
    if ((mask >> lane) & 1u) {
        uint8_t* p = out + (size_t)g * stride;
        if (cs) st_v8_cs(p, g * 8); else st_v8(p, g * 8);
    }
}
int main() {
    const uint32_t threads = 1 << 16;
    uint8_t* out;
    const size_t bytes = (size_t)threads * 128 + 256;
    CK(cudaMalloc(&out, bytes));
    std::vector<uint32_t> h(bytes / 4);
    const uint32_t masks[] = {0xFFFFFFFFu, 0xAAAAAAAAu, 0x0000FFFFu, 0x80000001u, 0x00000006u, 0xF0F0F0F0u};
    const uint32_t strides[] = {32, 64, 96, 128};
    for (int cs = 0; cs < 2; cs++)
        for (uint32_t stride : strides)
            for (uint32_t m : masks) {
                CK(cudaMemset(out, 0xEE, bytes));
                k<<<threads / 128, 128>>>(out, m, stride, cs);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(h.data(), out, bytes, cudaMemcpyDeviceToHost));
                size_t bad = 0, missing = 0;
                for (uint32_t g = 0; g < threads; g++) {
                    const bool on = (m >> (g & 31)) & 1u;
                    for (int j = 0; j < 8; j++) {
                        const uint32_t got = h[((size_t)g * stride) / 4 + j];
                        if (on && got != g * 8 + j) { bad++; if (got == 0xEEEEEEEEu) missing++; }
                    }
                }
                printf("cs=%d stride=%3u mask=%08x: %s (%zu wrong words, %zu never written)\n", cs, stride, m, bad ? "WRONG" : "ok", bad, missing);
            }
    return 0;
}
