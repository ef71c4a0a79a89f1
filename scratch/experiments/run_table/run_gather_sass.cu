// compile-only harness: nvcc -cubin, then count the SASS instructions of the interior-window path
#include "run_gather.cuh"
__global__ void gather_kernel(const RunTable* tables, uint4* out, uint32_t windows_per_response) {
    const uint32_t r = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
    __shared__ RunTable T[4];
    if (lane == 0) T[threadIdx.x / 32] = tables[r];
    __syncwarp();
    for (uint32_t w = lane; w < windows_per_response; w += 32) {
        uint32_t v[4];
        rg_window(T[threadIdx.x / 32], w * 16, v);
        out[(size_t)r * windows_per_response + w] = make_uint4(v[0], v[1], v[2], v[3]);
    }
}
