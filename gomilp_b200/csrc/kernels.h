// kernels.h — host-callable launchers of the simplex wave kernels. The kernels live in their own translation
// units (kernels_reg.cu: tier 1; kernels_generic.cu: tiers 2-5) so that they compile in parallel.
#pragma once
#include <cuda_runtime.h>

#include "simplex_cta.cuh"

namespace gm_kernels {
constexpr int kSmemThreads = 256;  // tiers 1, 2: one LP per CTA, W in shared memory
constexpr int kHbmThreads = 512;   // tiers 3-5: W (and possibly Bi) in HBM

// Set the dynamic shared-memory limit, report how many CTAs fit per SM, launch.
cudaError_t reg_prepare(size_t smem, int* ctas_per_sm);
void reg_launch(const gm::BatchParams& P, int grid, size_t smem, cudaStream_t st);
cudaError_t generic_prepare(int block, size_t smem, int* ctas_per_sm);
void generic_launch(const gm::BatchParams& P, int grid, int block, size_t smem, cudaStream_t st);
// tier 6 (kernels_coop.cu): cooperative launch, `P.coop_G` CTAs per LP
cudaError_t coop_prepare(size_t smem_max);
cudaError_t coop_occupancy(int block, size_t smem, int* ctas_per_sm);
cudaError_t coop_launch(const gm::BatchParams& P, int grid, int block, size_t smem, cudaStream_t st);
// every kernel's dynamic shared-memory limit is raised once per device (gm_init), never per launch
cudaError_t reg_set_smem_limit(size_t smem_max);
cudaError_t generic_set_smem_limit(size_t smem_max);
}  // namespace gm_kernels
