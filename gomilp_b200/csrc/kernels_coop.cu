// kernels_coop.cu — tier 6, the cooperative kernel: G CTAs of one cooperative launch share one LP (simplex_cta.cuh,
// "Cooperative tier"). Used when a launch has fewer LPs than SMs: one large LP (BASELINE config 4), the narrow
// first waves of a branch-and-bound. Pricing by column tiles, FTRAN / ratio test / fused rank-1 update by row blocks
// of the basis inverse (HBM / L2 resident), two group barriers per pivot; blocked Gauss-Jordan inversion with the
// trailing updates on the FP64 tensor cores (mma.sync m8n8k4 f64, SASS DMMA).
#include "kernels.h"

namespace {
__global__ void __launch_bounds__(gm_kernels::kHbmThreads, 1) simplex_wave_coop(gm::BatchParams P) {
    extern __shared__ __align__(128) double smem[];
    __shared__ int slot;
    gm::coop_cta_main(P, smem, &slot);
}
}  // namespace

namespace gm_kernels {
cudaError_t coop_prepare(size_t smem_max) {
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, simplex_wave_coop);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(simplex_wave_coop, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(smem_max - a.sharedSizeBytes));
}
cudaError_t coop_occupancy(int block, size_t smem, int* ctas_per_sm) {
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, simplex_wave_coop, block, smem);
}
// Cooperative launch: all CTAs are guaranteed co-resident (they wait on each other through group barriers); the
// runtime refuses the launch instead of deadlocking if the grid cannot be resident at once.
cudaError_t coop_launch(const gm::BatchParams& P, int grid, int block, size_t smem, cudaStream_t st) {
    gm::BatchParams Pc = P;
    void* args[] = {&Pc};
    return cudaLaunchCooperativeKernel((const void*)simplex_wave_coop, dim3(grid), dim3(block), args, smem, st);
}
}  // namespace gm_kernels
