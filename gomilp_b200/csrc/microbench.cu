// microbench.cu — measured denominators for the roofline of the shared-memory resident tiers ("smem peak
// to be measured by a microbenchmark"). gm_microbench_smem_gbs: every SM streams its shared memory with conflict-free
// 16-byte loads (the access width of the pricing / FTRAN loops of tier 1) from as many resident warps as fit.
#include <cuda_runtime.h>

#include "engine.h"

namespace {
__global__ void __launch_bounds__(1024, 1) smem_stream(double* sink, int iters) {
    extern __shared__ double2 sm2[];
    const int t = threadIdx.x, T = blockDim.x;
    const int words = 48 * 1024 / 16;  // 48 KB of double2
    for (int i = t; i < words; i += T) sm2[i] = make_double2(i, -i);
    __syncthreads();
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int idx = t;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double2 v = sm2[(idx + u * 1024) % words];
            if (u & 1) { a0 += v.x; a1 += v.y; } else { a2 += v.x; a3 += v.y; }
        }
        idx = (idx + 8 * 1024 + 32) % words;
    }
    if (a0 + a1 + a2 + a3 == 12345.678) sink[0] = a0;  // keeps the loads alive
}
}  // namespace

extern "C" int gm_microbench_smem_gbs(double* gbs_out) {
    using namespace gm_engine;
    if (!gbs_out) return GM_ERR_BAD_ARGUMENT;
    DeviceCtx* d = nullptr;
    int rc = current_device(&d);
    if (rc != GM_OK) return rc;
    StreamLease lease(d);
    CK(lease.err);
    cudaStream_t st = lease.se->s;
    double* sink = nullptr;
    CK(cudaMallocAsync(&sink, 8, st));
    const int iters = 4096, block = 1024, grid = d->sms;
    const size_t smem = 48 * 1024;
    smem_stream<<<grid, block, smem, st>>>(sink, 64);  // warm-up
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(lease.se->e[0], st));
        smem_stream<<<grid, block, smem, st>>>(sink, iters);
        CK(cudaEventRecord(lease.se->e[1], st));
        CK(cudaStreamSynchronize(st));
        const double bytes = (double)grid * block * (double)iters * 8 * 16;
        const double g = bytes / (ms(lease.se->e[0], lease.se->e[1]) * 1e-3) / 1e9;
        if (g > best) best = g;
    }
    CK(cudaGetLastError());
    cudaFreeAsync(sink, st);
    cudaStreamSynchronize(st);
    *gbs_out = best;
    return GM_OK;
}
