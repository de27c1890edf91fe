// kernels_generic.cu — tiers 2-5: one kernel, the tier only decides where W, the basis inverse and the vectors live.
//   2: everything in shared memory (256 threads, 2 CTAs / SM)
//   3: basis inverse + vectors in shared memory, W in HBM
//   4: W and basis inverse in HBM behind a TMA staging ring, vectors in shared memory
//   5: everything in HBM
#include "kernels.h"

namespace {
__global__ void __launch_bounds__(gm_kernels::kHbmThreads, 1) simplex_wave_generic(gm::BatchParams P) {
    extern __shared__ __align__(128) double smem[];
    __shared__ int slot;
    __shared__ unsigned long long bars[8];
    const int T = (int)blockDim.x;
    const bool hbm = P.hbm_layout != 0;
    const gm::WsLayout w = gm::ws_layout(P.m0 + P.L, P.n0 + P.L, T, false, hbm, P.tier == 2 || P.tier == 3);
    double* work = P.work ? P.work + (size_t)blockIdx.x * P.work_stride : nullptr;
    // one call site: the solver is large and fully inlined
    double *wbase, *bibase, *small, *ring = nullptr;
    if (P.tier == 2) {
        wbase = smem + w.W; bibase = smem + w.Bi; small = smem + w.big_doubles;
    } else if (P.tier == 3) {
        wbase = work; bibase = smem; small = smem + (w.big_doubles - w.Bi);
    } else if (P.tier == 4) {
        wbase = work + w.W; bibase = work + w.Bi;
        small = smem + (size_t)P.ring_stages * (P.ring_stage_bytes / 8);
        if (P.ring_stages > 0) ring = smem;
    } else {
        wbase = work + w.W; bibase = work + w.Bi; small = work + w.big_doubles;
    }
    gm::cta_main<false>(P, wbase, bibase, small, &slot, ring, bars);
}
}  // namespace

namespace gm_kernels {
cudaError_t generic_set_smem_limit(size_t smem_max) {
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, simplex_wave_generic);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(simplex_wave_generic, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(smem_max - a.sharedSizeBytes));
}
cudaError_t generic_prepare(int block, size_t smem, int* ctas_per_sm) {
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, simplex_wave_generic, block, smem);
}
void generic_launch(const gm::BatchParams& P, int grid, int block, size_t smem, cudaStream_t st) {
    simplex_wave_generic<<<grid, block, smem, st>>>(P);
}
}  // namespace gm_kernels
