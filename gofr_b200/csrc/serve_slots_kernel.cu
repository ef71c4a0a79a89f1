// serve_slots_kernel.cu — slot-layout instance of the serve kernel (serve_body.cuh): the resident headline kernel.
#include "serve_body.cuh"

namespace gofr {

__global__ void __launch_bounds__(T, kServeCtas) serve_slots_kernel(const __grid_constant__ ServeParams p) { serve_body<true>(p); }

int serve_slots_blocks_per_sm(uint32_t smem_bytes) {
    if (cudaFuncSetAttribute(serve_slots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) return -1;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, serve_slots_kernel, T, smem_bytes) != cudaSuccess) return -1;
    return nb;
}

int launch_serve_slots(const ServeParams& p, int grid, uint32_t smem_bytes, void* stream) {
    serve_slots_kernel<<<grid, T, smem_bytes, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}

}  // namespace gofr
