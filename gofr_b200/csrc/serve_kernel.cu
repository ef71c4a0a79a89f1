// serve_kernel.cu — packed-layout instance of the serve kernel (serve_body.cuh) and the launch helpers.
//
// This instance does not know the programs of the wider data model (PF_VALUES: float64, pointers, slices, maps ...): the
// packed layout has the interpreter inlined in its tile loop, and the two extra call sites cost every table registers
// (serve_device.cuh GOFR_TU_VALUES).  Tables that contain such a program are served by serve_values_kernel.cu.
#define GOFR_TU_VALUES 0
#include "serve_body.cuh"

namespace gofr {

__global__ void __launch_bounds__(T, kServeCtas) serve_kernel(const __grid_constant__ ServeParams p) { serve_body<false>(p); }

// slot-layout instance: serve_slots_kernel.cu
int serve_slots_blocks_per_sm(uint32_t smem_bytes, bool wide);
int launch_serve_slots(const ServeParams& p, int grid, uint32_t smem_bytes, void* stream, bool wide, bool values);
// packed layout, tables with PF_VALUES programs: serve_values_kernel.cu
int serve_values_blocks_per_sm(uint32_t smem_bytes);
int launch_serve_values(const ServeParams& p, int grid, uint32_t smem_bytes, void* stream);

uint32_t serve_smem_bytes(uint32_t hot_bytes, uint32_t in_cap) {
    return ((hot_bytes + 127u) & ~127u) + in_cap + 64;
}

int serve_max_grid(uint32_t smem_bytes, int device, int* blocks_per_sm, int* wide_grid) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(serve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) return -1;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, serve_kernel, T, smem_bytes) != cudaSuccess) return -1;
    const int nb2 = serve_slots_blocks_per_sm(smem_bytes, false);
    const int nbw = serve_slots_blocks_per_sm(smem_bytes, true);
    const int nbv = serve_values_blocks_per_sm(smem_bytes);
    if (nb2 < 0 || nbw < 0 || nbv < 0) return -1;
    if (nb2 < nb) nb = nb2;
    if (nbv < nb) nb = nbv;
    if (wide_grid) *wide_grid = nbw * prop.multiProcessorCount;
    if (blocks_per_sm) *blocks_per_sm = nb;
    return nb * prop.multiProcessorCount;
}

int launch_serve(const ServeParams& p, int grid, uint32_t smem_bytes, void* stream, bool wide_slots, bool values) {
    if (p.slot_bytes) return launch_serve_slots(p, grid, smem_bytes, stream, wide_slots && !values, values);
    if (values) return launch_serve_values(p, grid, smem_bytes, stream);
    serve_kernel<<<grid, T, smem_bytes, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}

}  // namespace gofr
