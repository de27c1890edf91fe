// kernels_reg.cu — tier 1 (m <= 64): basis inverse in registers, W + staging tile + vectors in shared memory.
#include "kernels.h"

namespace {
// cold waves / batches (the bench path): no warm-start code in the kernel
__global__ void __launch_bounds__(gm_kernels::kSmemThreads, 2) simplex_wave_reg(gm::BatchParams P) {
    extern __shared__ double smem[];
    __shared__ int slot;
    const gm::WsLayout w = gm::ws_layout(P.m0 + P.L, P.n0 + P.L, gm_kernels::kSmemThreads, true);
    gm::cta_main<true, false>(P, smem + w.W, smem + w.Bi, smem + w.big_doubles, &slot);
}
// gm_solve_wave_warm: children start from the parent's inverse and every node writes its own back to HBM
__global__ void __launch_bounds__(gm_kernels::kSmemThreads, 2) simplex_wave_reg_warm(gm::BatchParams P) {
    extern __shared__ double smem[];
    __shared__ int slot;
    const gm::WsLayout w = gm::ws_layout(P.m0 + P.L, P.n0 + P.L, gm_kernels::kSmemThreads, true);
    gm::cta_main<true, true>(P, smem + w.W, smem + w.Bi, smem + w.big_doubles, &slot);
}
}  // namespace

namespace gm_kernels {
// the dynamic limit is the opt-in maximum minus the kernel's static shared memory
template <class K>
static cudaError_t raise_limit(K kernel, size_t smem_max) {
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, kernel);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_max - a.sharedSizeBytes));
}
cudaError_t reg_set_smem_limit(size_t smem_max) {
    cudaError_t e = raise_limit(simplex_wave_reg, smem_max);
    if (e != cudaSuccess) return e;
    return raise_limit(simplex_wave_reg_warm, smem_max);
}
cudaError_t reg_prepare(size_t smem, int* ctas_per_sm) {
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, simplex_wave_reg, kSmemThreads, smem);
}
void reg_launch(const gm::BatchParams& P, int grid, size_t smem, cudaStream_t st) {
    if (P.warm_parent || P.bi_out || P.trace || P.robust) simplex_wave_reg_warm<<<grid, kSmemThreads, smem, st>>>(P);
    else simplex_wave_reg<<<grid, kSmemThreads, smem, st>>>(P);
}
}  // namespace gm_kernels
