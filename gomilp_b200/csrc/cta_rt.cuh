// cta_rt.cuh — the handful of CTA-level primitives the simplex kernels are written against.
//
// On the device (the product build, nvcc sm_100a) they are the CUDA intrinsics. With -DGM_EMULATE the
// same kernel source is compiled by g++ against tests/emu/cta_emu.hpp, a fiber-based single-CTA
// emulator used ONLY by the CPU test-suite to exercise the kernel's control flow where there is no
// GPU (this container has none). The emulator is never linked into libgomilp_b200.so.
#pragma once

#ifdef GM_EMULATE
#include "cta_emu.hpp"
#else
#include <cuda_runtime.h>
#define GM_DEV __device__ __forceinline__
#define GM_DEV_NOINLINE __device__ __noinline__
GM_DEV int gm_tid() { return (int)threadIdx.x; }
GM_DEV int gm_nthreads() { return (int)blockDim.x; }
GM_DEV void gm_sync() { __syncthreads(); }
GM_DEV double gm_shfl_down(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
GM_DEV int gm_shfl_down(int v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
GM_DEV double gm_shfl_xor(double v, int d) { return __shfl_xor_sync(0xffffffffu, v, d); }
GM_DEV int gm_shfl_xor(int v, int d) { return __shfl_xor_sync(0xffffffffu, v, d); }
GM_DEV double gm_shfl_idx(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
GM_DEV unsigned gm_ballot(int pred) { return __ballot_sync(0xffffffffu, pred); }
GM_DEV int gm_warp_min_int(int v) { return __reduce_min_sync(0xffffffffu, v); }
GM_DEV unsigned gm_warp_min_u32(unsigned v) { return __reduce_min_sync(0xffffffffu, v); }
GM_DEV unsigned gm_warp_max_u32(unsigned v) { return __reduce_max_sync(0xffffffffu, v); }
GM_DEV unsigned long long gm_d2bits(double v) { return (unsigned long long)__double_as_longlong(v); }
GM_DEV double gm_bits2d(unsigned long long b) { return __longlong_as_double((long long)b); }
GM_DEV int gm_popc(unsigned v) { return __popc(v); }
GM_DEV void gm_syncwarp() { __syncwarp(); }
GM_DEV int gm_shfl_idx(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
GM_DEV int gm_any(int pred) { return __any_sync(0xffffffffu, pred); }
GM_DEV int gm_atomic_add(int* p, int v) { return atomicAdd(p, v); }
GM_DEV double gm_ldg(const double* p) { return __ldg(p); }

// ---- multi-CTA groups (the cooperative kernel: several CTAs of one cooperative launch work on one LP) ----------
GM_DEV int gm_block_id() { return (int)blockIdx.x; }
GM_DEV void gm_threadfence() { __threadfence(); }
GM_DEV void gm_spin_pause() {}
GM_DEV long long gm_clock() { return clock64(); }
// Hides where a value came from, so that the compiler keeps it instead of recomputing it from kernel parameters.
template <class P>
GM_DEV void gm_opaque(P*& p) { asm volatile("" : "+l"(p)); }
GM_DEV void gm_opaque_i(int& v) { asm volatile("" : "+r"(v)); }
// Waits until *ready > item (a counter the copy engine advances from another stream, hence system scope). Bounded:
// ~10 s of polling, then false.
GM_DEV bool gm_wait_ready(const int* ready, int item) {
    for (int spin = 0; spin < (1 << 25); ++spin) {
        int v;
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(ready) : "memory");
        if (v > item) return true;
        __nanosleep(256);
    }
    return false;
}
GM_DEV void gm_atomic_add_u64(unsigned long long* p, unsigned long long v) { atomicAdd(p, v); }
GM_DEV void gm_red_release_add_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
GM_DEV unsigned long long gm_ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// One FP64 tensor-core step (SASS DMMA): D(8x8) = A(8x4) * B(4x8) + C, warp-wide. Lane l holds A[l>>2][l&3],
// B[l&3][l>>2] and C/D[l>>2][2*(l&3) + {0,1}].
GM_DEV void gm_dmma_8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on a shared-memory mbarrier -------------------
GM_DEV unsigned gm_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
GM_DEV void gm_mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gm_smem_u32(bar)), "r"(count) : "memory");
}
// orders this thread's earlier generic-proxy writes (st.global / st.shared) before later async-proxy (TMA) accesses
GM_DEV void gm_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
GM_DEV void gm_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
GM_DEV void gm_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gm_smem_u32(bar)), "r"(bytes) : "memory");
}
// dst (shared, 16 B aligned) <- src (global, 16 B aligned), bytes a multiple of 16
GM_DEV void gm_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     gm_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(gm_smem_u32(bar))
                 : "memory");
}
// Waits for the phase with the given parity; bounded so that a programming error cannot hang the GPU.
GM_DEV bool gm_mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned a = gm_smem_u32(bar);
    for (int spin = 0; spin < (1 << 24); ++spin) {
        unsigned ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}
#endif
