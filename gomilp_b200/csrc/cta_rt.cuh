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
#endif
