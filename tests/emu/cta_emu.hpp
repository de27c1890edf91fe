// tests/emu/cta_emu.hpp — TEST INFRASTRUCTURE ONLY.
//
// A single-CTA emulator: every CUDA thread of one thread block is a ucontext fiber scheduled
// round-robin on one OS thread; __syncthreads() and the warp shuffles/votes are rendezvous points.
// It lets the CPU test-suite (which has no GPU) drive the SAME kernel source that nvcc compiles for
// sm_100a (gomilp_b200/csrc/simplex_cta.cuh, built here with g++ -DGM_EMULATE) through its control
// flow: phase transitions, Bland fallback, status codes. It is a debugging aid for the kernel, not a
// solver: nothing under gomilp_b200/ links it and the product library has no CPU path.
//
// Deliberate limitations: deterministic (optionally shuffled) scheduling means a missing barrier is
// only caught when it changes a result; compute-sanitizer racecheck on the GPU is the real check.
#pragma once
#include <ucontext.h>

#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <vector>

namespace emu {

enum Wait { RUN = 0, WAIT_BLOCK = 1, WAIT_WARP = 2, DONE = 3 };

struct Cta {
    int T = 0;        // threads per CTA
    int NB = 1;       // CTAs emulated together (a cooperative launch); fiber f = block * T + thread
    int cur = 0;
    std::vector<ucontext_t> ctx;
    std::vector<std::vector<char>> stacks;
    std::vector<int> state;
    std::vector<uint64_t> slot;  // shuffle exchange slots
    std::vector<int> vote;
    ucontext_t sched;
    std::function<void()> body;
    long barriers = 0;
    unsigned rng = 12345;
    bool shuffle_order = false;
};

inline Cta*& current() {
    static thread_local Cta* c = nullptr;
    return c;
}

inline void yield_as(int st) {
    Cta* c = current();
    c->state[c->cur] = st;
    swapcontext(&c->ctx[c->cur], &c->sched);
}

inline void trampoline() {
    Cta* c = current();
    c->body();
    c->state[c->cur] = DONE;
    swapcontext(&c->ctx[c->cur], &c->sched);
}

// Runs body() on T fibers to completion. Throws on barrier divergence (some threads exit or wait on a
// different barrier kind while others wait forever).
inline void run_grid(int nblocks, int T, std::function<void()> body, bool shuffle_order = false,
                     size_t stack_bytes = 256 * 1024);
inline void run_cta(int T, std::function<void()> body, bool shuffle_order = false, size_t stack_bytes = 256 * 1024) {
    run_grid(1, T, std::move(body), shuffle_order, stack_bytes);
}
// `nblocks` CTAs of T threads each, all resident at once (what a cooperative launch guarantees). Block barriers
// release per CTA; CTAs talk to each other only through memory (atomics + spin loops that call gm_spin_pause()).
inline void run_grid(int nblocks, int Tper, std::function<void()> body, bool shuffle_order, size_t stack_bytes) {
    Cta c;
    c.T = Tper;
    c.NB = nblocks;
    const int T = Tper * nblocks;
    c.body = std::move(body);
    c.ctx.resize(T);
    c.stacks.resize(T);
    c.state.assign(T, RUN);
    c.slot.assign(T, 0);
    c.vote.assign(T, 0);
    c.shuffle_order = shuffle_order;
    Cta* prev = current();
    current() = &c;
    for (int t = 0; t < T; ++t) {
        c.stacks[t].resize(stack_bytes);
        getcontext(&c.ctx[t]);
        c.ctx[t].uc_stack.ss_sp = c.stacks[t].data();
        c.ctx[t].uc_stack.ss_size = stack_bytes;
        c.ctx[t].uc_link = &c.sched;
        makecontext(&c.ctx[t], (void (*)())trampoline, 0);
    }
    std::vector<int> order(T);
    for (int t = 0; t < T; ++t) order[t] = t;
    for (;;) {
        bool progressed = false;
        if (c.shuffle_order) {
            for (int t = T - 1; t > 0; --t) {
                c.rng = c.rng * 1664525u + 1013904223u;
                int j = (int)((c.rng >> 8) % (unsigned)(t + 1));
                std::swap(order[t], order[j]);
            }
        }
        for (int k = 0; k < T; ++k) {
            int t = order[k];
            if (c.state[t] != RUN) continue;
            c.cur = t;
            progressed = true;
            swapcontext(&c.sched, &c.ctx[t]);
        }
        // release complete warps waiting on a warp rendezvous
        for (int w0 = 0; w0 < T; w0 += 32) {
            int w1 = std::min(T, w0 + 32);
            bool all = true;
            for (int t = w0; t < w1; ++t)
                if (c.state[t] != WAIT_WARP) { all = false; break; }
            if (all) {
                for (int t = w0; t < w1; ++t) c.state[t] = RUN;
                progressed = true;
            }
        }
        int ndone = 0, nrun = 0;
        bool released = false;
        for (int b0 = 0; b0 < T; b0 += Tper) {  // block barriers release per CTA
            int nblock = 0;
            for (int t = b0; t < b0 + Tper; ++t) nblock += c.state[t] == WAIT_BLOCK;
            if (nblock == Tper) {
                for (int t = b0; t < b0 + Tper; ++t) c.state[t] = RUN;
                c.barriers++;
                released = true;
            }
        }
        for (int t = 0; t < T; ++t) {
            ndone += c.state[t] == DONE;
            nrun += c.state[t] == RUN;
        }
        if (ndone == T) break;
        if (released) continue;
        if (nrun == 0 && !progressed) {
            current() = prev;
            throw std::runtime_error("cta_emu: barrier divergence / deadlock");
        }
        if (nrun == 0) {
            // nothing runnable and no warp/block released: divergence
            bool any_release = false;
            for (int t = 0; t < T; ++t) any_release |= c.state[t] == RUN;
            if (!any_release) {
                current() = prev;
                throw std::runtime_error("cta_emu: barrier divergence (mixed waits)");
            }
        }
    }
    current() = prev;
}

template <class V>
inline V shfl_generic(V v, int src_lane_abs_valid, int src) {
    Cta* c = current();
    uint64_t bits = 0;
    static_assert(sizeof(V) <= 8, "shuffle payload");
    std::memcpy(&bits, &v, sizeof(V));
    c->slot[c->cur] = bits;
    yield_as(WAIT_WARP);
    V out = v;
    if (src_lane_abs_valid) {
        uint64_t b = c->slot[src];
        std::memcpy(&out, &b, sizeof(V));
    }
    yield_as(WAIT_WARP);
    return out;
}

}  // namespace emu

struct alignas(16) double2 {
    double x, y;
};
#define GM_DEV inline
#define GM_DEV_NOINLINE inline
inline int gm_tid() { return emu::current()->cur % emu::current()->T; }
inline int gm_nthreads() { return emu::current()->T; }
inline void gm_sync() { emu::yield_as(emu::WAIT_BLOCK); }
template <class V>
inline V gm_shfl_down_t(V v, int d) {
    emu::Cta* c = emu::current();
    int lane = c->cur & 31;
    int src = c->cur + d;
    bool ok = (lane + d) < 32 && src < (c->T * c->NB);
    return emu::shfl_generic(v, ok, src);
}
template <class V>
inline V gm_shfl_xor_t(V v, int d) {
    emu::Cta* c = emu::current();
    int src = (c->cur & ~31) | ((c->cur & 31) ^ d);
    bool ok = src < (c->T * c->NB);
    return emu::shfl_generic(v, ok, src);
}
inline double gm_shfl_down(double v, int d) { return gm_shfl_down_t(v, d); }
inline int gm_shfl_down(int v, int d) { return gm_shfl_down_t(v, d); }
inline double gm_shfl_xor(double v, int d) { return gm_shfl_xor_t(v, d); }
inline int gm_shfl_xor(int v, int d) { return gm_shfl_xor_t(v, d); }
inline double gm_shfl_idx(double v, int src_lane) {
    emu::Cta* c = emu::current();
    int src = (c->cur & ~31) | (src_lane & 31);
    return emu::shfl_generic(v, src < (c->T * c->NB), src);
}
inline int gm_shfl_idx(int v, int src_lane) {
    emu::Cta* c = emu::current();
    int src = (c->cur & ~31) | (src_lane & 31);
    return emu::shfl_generic(v, src < (c->T * c->NB), src);
}
inline unsigned gm_ballot(int pred) {
    emu::Cta* c = emu::current();
    c->vote[c->cur] = pred ? 1 : 0;
    emu::yield_as(emu::WAIT_WARP);
    int w0 = c->cur & ~31, w1 = std::min((c->T * c->NB), w0 + 32);
    unsigned r = 0;
    for (int t = w0; t < w1; ++t) r |= (unsigned)c->vote[t] << (t - w0);
    emu::yield_as(emu::WAIT_WARP);
    return r;
}
inline int gm_warp_min_int(int v) {
    emu::Cta* c = emu::current();
    c->vote[c->cur] = v;
    emu::yield_as(emu::WAIT_WARP);
    int w0 = c->cur & ~31, w1 = std::min((c->T * c->NB), w0 + 32), r = INT_MAX;
    for (int t = w0; t < w1; ++t) r = std::min(r, c->vote[t]);
    emu::yield_as(emu::WAIT_WARP);
    return r;
}
inline unsigned gm_warp_min_u32(unsigned v) {
    emu::Cta* c = emu::current();
    c->vote[c->cur] = (int)v;
    emu::yield_as(emu::WAIT_WARP);
    int w0 = c->cur & ~31, w1 = std::min((c->T * c->NB), w0 + 32);
    unsigned r = 0xffffffffu;
    for (int t = w0; t < w1; ++t) r = std::min(r, (unsigned)c->vote[t]);
    emu::yield_as(emu::WAIT_WARP);
    return r;
}
inline unsigned gm_warp_max_u32(unsigned v) {
    emu::Cta* c = emu::current();
    c->vote[c->cur] = (int)v;
    emu::yield_as(emu::WAIT_WARP);
    int w0 = c->cur & ~31, w1 = std::min((c->T * c->NB), w0 + 32);
    unsigned r = 0;
    for (int t = w0; t < w1; ++t) r = std::max(r, (unsigned)c->vote[t]);
    emu::yield_as(emu::WAIT_WARP);
    return r;
}
inline unsigned long long gm_d2bits(double v) {
    unsigned long long b;
    std::memcpy(&b, &v, 8);
    return b;
}
inline double gm_bits2d(unsigned long long b) {
    double v;
    std::memcpy(&v, &b, 8);
    return v;
}
inline int gm_popc(unsigned v) { return __builtin_popcount(v); }
inline void gm_syncwarp() { emu::yield_as(emu::WAIT_WARP); }
inline int gm_any(int pred) {
    emu::Cta* c = emu::current();
    c->vote[c->cur] = pred ? 1 : 0;
    emu::yield_as(emu::WAIT_WARP);
    int w0 = c->cur & ~31, w1 = std::min((c->T * c->NB), w0 + 32), r = 0;
    for (int t = w0; t < w1; ++t) r |= c->vote[t];
    emu::yield_as(emu::WAIT_WARP);
    return r;
}
inline int gm_atomic_add(int* p, int v) {
    int o = *p;
    *p = o + v;
    return o;
}
inline double gm_ldg(const double* p) { return *p; }

// ---- multi-CTA groups ------------------------------------------------------------------------------------------
inline int gm_block_id() { return emu::current()->cur / emu::current()->T; }
inline void gm_threadfence() {}
inline void gm_spin_pause() { emu::yield_as(emu::RUN); }
inline long long gm_clock() { return 0; }
template <class P>
inline void gm_opaque(P*&) {}
inline void gm_opaque_i(int&) {}
inline bool gm_wait_ready(const int* ready, int item) { return *ready > item; }
inline void gm_atomic_add_u64(unsigned long long* p, unsigned long long v) { *p += v; }
inline void gm_red_release_add_u64(unsigned long long* p, unsigned long long v) { *p += v; }
inline unsigned long long gm_ld_acquire_u64(const unsigned long long* p) { return *(volatile const unsigned long long*)p; }
// D(8x8) = A(8x4) * B(4x8) + C with the m8n8k4 fragment layout of mma.sync (see cta_rt.cuh)
inline void gm_dmma_8x8x4(double& c0, double& c1, double a, double b) {
    emu::Cta* c = emu::current();
    const int lane = c->cur & 31, w0 = c->cur & ~31;
    uint64_t ab, bb;
    std::memcpy(&ab, &a, 8);
    std::memcpy(&bb, &b, 8);
    c->slot[c->cur] = ab;
    emu::yield_as(emu::WAIT_WARP);
    double arow[4];
    for (int k = 0; k < 4; ++k) std::memcpy(&arow[k], &c->slot[w0 + (lane >> 2) * 4 + k], 8);  // A[row][k]
    emu::yield_as(emu::WAIT_WARP);
    c->slot[c->cur] = bb;
    emu::yield_as(emu::WAIT_WARP);
    for (int u = 0; u < 2; ++u) {
        const int col = 2 * (lane & 3) + u;
        double acc = u == 0 ? c0 : c1;
        for (int k = 0; k < 4; ++k) {
            double bv;
            std::memcpy(&bv, &c->slot[w0 + col * 4 + k], 8);  // B[k][col] is held by lane col*4 + k
            acc += arow[k] * bv;
        }
        (u == 0 ? c0 : c1) = acc;
    }
    emu::yield_as(emu::WAIT_WARP);
}

// ---- TMA bulk copy + mbarrier stand-ins: the copy happens at issue time, the barrier counts bytes ----------
// An emulated mbarrier word: low 32 bits = bytes still expected in the current phase (two's complement while
// copies outrun expect_tx), bit 63 = current phase parity.
inline void gm_mbar_init(unsigned long long* bar, int) { *bar = 0; }
inline void gm_mbar_fence_init() {}
inline void gm_fence_proxy_async() {}
inline void emu_mbar_add(unsigned long long* bar, long long delta, bool is_expect) {
    long long pending = (long long)(int)(unsigned)(*bar & 0xffffffffull);
    unsigned long long phase = *bar >> 63;
    unsigned long long armed = (*bar >> 62) & 1ull;
    pending += delta;
    if (is_expect) armed = 1;
    if (armed && pending == 0) { phase ^= 1ull; armed = 0; }
    *bar = (phase << 63) | (armed << 62) | (unsigned long long)(unsigned)(int)pending;
}
inline void gm_mbar_expect_tx(unsigned long long* bar, unsigned bytes) { emu_mbar_add(bar, (long long)bytes, true); }
inline void gm_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    if ((reinterpret_cast<uintptr_t>(dst) & 15) || (reinterpret_cast<uintptr_t>(src) & 15) || (bytes & 15))
        throw std::runtime_error("cta_emu: misaligned bulk copy");
    std::memcpy(dst, src, bytes);
    emu_mbar_add(bar, -(long long)bytes, false);
}
inline bool gm_mbar_wait(unsigned long long* bar, unsigned parity) {
    for (int spin = 0; spin < (1 << 20); ++spin) {
        if ((unsigned)(*bar >> 63) != parity) return true;  // the phase with this parity has completed
        emu::yield_as(emu::RUN);
    }
    return false;
}
