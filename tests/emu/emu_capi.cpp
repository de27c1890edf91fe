// tests/emu/emu_capi.cpp — TEST INFRASTRUCTURE ONLY. Compiles the product kernel source
// (gomilp_b200/csrc/simplex_cta.cuh) with g++ -DGM_EMULATE against the fiber CTA emulator, so the
// CPU test-suite can drive the kernel's control flow on small LPs in a container without a GPU.
// Never linked into libgomilp_b200.so; see cta_emu.hpp.
#define GM_EMULATE 1
#include "cta_emu.hpp"
#include "../../gomilp_b200/csrc/simplex_cta.cuh"

static const int* g_emu_lp_list = nullptr;
static int g_emu_robust = 0;  // > 0: emulated launches solve with BatchParams::robust
static int g_emu_coop_pan = 1;  // 0: force the inversion panel of the cooperative tier into (emulated) HBM
static int g_emu_quad = 0;  // 1: run the generic solver as tier 2 (quad-mapped main loop)  // retry launches: work item k solves LP list[k]

extern "C" {

// Same meaning as the device batch entry: `count` LPs, shared (stride 0) or per-LP roots, optional
// branch rows. T = emulated threads per CTA (power of two, multiple of 32).
int emu_simplex_batch_ex(int count, const double* c, const double* A, const double* b, long long c_stride,
                         long long A_stride, long long b_stride, int lda, int m0, int n0, int L, const int* bvar,
                         const double* bsign, const double* brhs, const long long* initial_basic, double tol,
                         int max_pivots, int refactor_period, int* status, double* optF, double* x, long long x_stride,
                         int x_len, long long* basis, int* stats, int T, int shuffle_order, int reg, int ring_stages,
                         int ring_stage_bytes, const int* warm_parent, const long long* warm_basis,
                         const double* warm_bi, double* bi_out);

int emu_simplex_batch(int count, const double* c, const double* A, const double* b, long long c_stride,
                      long long A_stride, long long b_stride, int lda, int m0, int n0, int L, const int* bvar,
                      const double* bsign, const double* brhs, const long long* initial_basic, double tol,
                      int max_pivots, int refactor_period, int* status, double* optF, double* x, long long x_stride,
                      int x_len, long long* basis, int* stats, int T, int shuffle_order, int reg, int ring_stages,
                      int ring_stage_bytes) {
    return emu_simplex_batch_ex(count, c, A, b, c_stride, A_stride, b_stride, lda, m0, n0, L, bvar, bsign, brhs,
                                initial_basic, tol, max_pivots, refactor_period, status, optF, x, x_stride, x_len, basis,
                                stats, T, shuffle_order, reg, ring_stages, ring_stage_bytes, nullptr, nullptr, nullptr,
                                nullptr);
}

int emu_simplex_batch_ex(int count, const double* c, const double* A, const double* b, long long c_stride,
                         long long A_stride, long long b_stride, int lda, int m0, int n0, int L, const int* bvar,
                         const double* bsign, const double* brhs, const long long* initial_basic, double tol,
                         int max_pivots, int refactor_period, int* status, double* optF, double* x, long long x_stride,
                         int x_len, long long* basis, int* stats, int T, int shuffle_order, int reg, int ring_stages,
                         int ring_stage_bytes, const int* warm_parent, const long long* warm_basis,
                         const double* warm_bi, double* bi_out) {
    gm::BatchParams P;
    std::memset(&P, 0, sizeof(P));
    P.c = c; P.A = A; P.b = b;
    P.c_stride = c_stride; P.A_stride = A_stride; P.b_stride = b_stride;
    P.lda = lda; P.m0 = m0; P.n0 = n0; P.L = L;
    P.bvar = bvar; P.bsign = bsign; P.brhs = brhs;
    P.initial_basic = initial_basic;
    P.tol = tol; P.count = count; P.max_pivots = max_pivots; P.refactor_period = refactor_period;
    P.status = status; P.optF = optF; P.x = x; P.x_stride = x_stride; P.x_len = x_len;
    P.basis = basis; P.stats = stats;
    P.warm_parent = warm_parent; P.warm_basis = warm_basis; P.warm_bi = warm_bi; P.bi_out = bi_out;
    P.lp_list = g_emu_lp_list;
    P.robust = g_emu_robust > 0;
    int queue = 0;
    P.queue = &queue;
    if (reg && (T != 256 || m0 + L > 64)) return -2;
    const bool hbm = ring_stages > 0;
    P.hbm_layout = hbm ? 1 : 0;
    P.ring_stages = ring_stages;
    P.ring_stage_bytes = ring_stage_bytes;
    P.tier = reg ? 1 : (g_emu_quad ? 2 : (hbm ? 4 : 5));
    gm::WsLayout w = gm::ws_layout(m0 + L, n0 + L, T, reg != 0, hbm, P.tier == 2);
    std::vector<double> ringbuf((size_t)ring_stages * ring_stage_bytes / 8 + 32, 0.0);
    double* ring = hbm ? reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(ringbuf.data()) + 127) & ~uintptr_t(127)) : nullptr;
    unsigned long long bars[8] = {0};
    std::vector<double> bigbuf(w.big_doubles + 40, 0.0), small(w.small_bytes / 8 + 8, 0.0);
    double* big = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(bigbuf.data()) + 127) & ~uintptr_t(127));
    int slot = 0;
    try {
        if (reg) emu::run_cta(T, [&]() { gm::cta_main<true>(P, big + w.W, big + w.Bi, small.data(), &slot); }, shuffle_order != 0);
        else emu::run_cta(T, [&]() { gm::cta_main<false>(P, big + w.W, big + w.Bi, small.data(), &slot, ring, bars); }, shuffle_order != 0);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return -1;
    }
    return 0;
}


// The cooperative tier (simplex_cta.cuh, COOP): `groups` groups of G emulated CTAs of T threads each, all resident at
// once like a cooperative launch. trace (optional): int[trace_cap][4] pivot records of LP trace_lp.
int emu_coop_batch(int count, const double* c, const double* A, const double* b, long long c_stride,
                   long long A_stride, long long b_stride, int lda, int m0, int n0, int L, const int* bvar,
                   const double* bsign, const double* brhs, const long long* initial_basic, double tol, int max_pivots,
                   int refactor_period, int* status, double* optF, double* x, long long x_stride, int x_len,
                   long long* basis, int* stats, int T, int G, int groups, int shuffle_order, int* trace, int trace_cap,
                   int trace_lp) {
    gm::BatchParams P;
    std::memset(&P, 0, sizeof(P));
    P.c = c; P.A = A; P.b = b;
    P.c_stride = c_stride; P.A_stride = A_stride; P.b_stride = b_stride;
    P.lda = lda; P.m0 = m0; P.n0 = n0; P.L = L;
    P.bvar = bvar; P.bsign = bsign; P.brhs = brhs;
    P.initial_basic = initial_basic;
    P.tol = tol; P.count = count; P.max_pivots = max_pivots; P.refactor_period = refactor_period;
    P.status = status; P.optF = optF; P.x = x; P.x_stride = x_stride; P.x_len = x_len;
    P.basis = basis; P.stats = stats;
    P.trace = trace; P.trace_cap = trace_cap; P.trace_lp = trace_lp;
    P.robust = g_emu_robust > 0;
    int queue = 0;
    P.queue = &queue;
    P.hbm_layout = 1;
    P.tier = 6;
    P.coop_G = G;
    if (T % 32 != 0 || G < 1 || groups < 1) return -2;
    const gm::CoopLayout cl = gm::coop_layout(m0 + L, n0 + L, T, G, g_emu_coop_pan ? (size_t)200 * 1024 : 0);
    P.coop_pan = cl.pan_nb;
    P.coop_small = cl.small_in_smem;
    std::vector<double> work((size_t)groups * cl.group_doubles + 16, 0.0);
    std::vector<unsigned long long> bars(groups, 0ull);
    P.work = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(work.data()) + 31) & ~uintptr_t(31));
    P.work_stride = (long long)cl.group_doubles;
    P.coop_bar = bars.data();
    const int nblocks = G * groups;
    std::vector<std::vector<double>> smem(nblocks, std::vector<double>(cl.smem_bytes / 8 + 8, 0.0));
    std::vector<int> slots(nblocks, 0);
    try {
        emu::run_grid(nblocks, T, [&]() {
            const int b2 = gm_block_id();
            gm::coop_cta_main(P, smem[b2].data(), &slots[b2]);
        }, shuffle_order != 0);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return -1;
    }
    return 0;
}

}  // extern "C"

// ---- emulator-backed stand-ins for the wave entry points, so that the product's B&B host
// (gomilp_b200/csrc/bnb_host.cpp, compiled into this TEST library unchanged) can be replayed on the CPU.
#include <map>
#include "../../include/gomilp_b200.h"

namespace {
struct EmuRoot {
    std::vector<double> c, A, b;
    int m0, n0;
    std::vector<double> prev_bi;
    std::vector<long long> prev_basis;
    int64_t prev_nodes = 0;
    int prev_m = 0;
};
std::map<gm_root_t, EmuRoot> g_roots;
gm_root_t g_next = 1;
int g_T = 64;
int g_reg = 0;
}  // namespace

extern "C" {
void emu_set_threads(int T, int reg) { g_T = T; g_reg = reg; }
void emu_set_quad(int on) { g_emu_quad = on; }
void emu_set_coop_pan(int on) { g_emu_coop_pan = on; }
int gm_upload_root(const double* c0, const double* A0, int64_t lda, const double* b0, int64_t m0, int64_t n0,
                   gm_root_t* out) {
    EmuRoot r;
    r.m0 = (int)m0; r.n0 = (int)n0;
    r.c.assign(c0, c0 + n0);
    r.b.assign(b0, b0 + m0);
    r.A.resize((size_t)m0 * n0);
    for (int64_t i = 0; i < m0; ++i)
        for (int64_t j = 0; j < n0; ++j) r.A[(size_t)i * n0 + j] = A0[(size_t)i * lda + j];
    *out = g_next++;
    g_roots[*out] = std::move(r);
    return GM_OK;
}
int gm_free_root(gm_root_t h) { return g_roots.erase(h) ? GM_OK : GM_ERR_BAD_HANDLE; }
int gm_last_timing(gm_timing* t) { std::memset(t, 0, sizeof(*t)); return GM_OK; }
int gm_solve_wave(gm_root_t root, int64_t nodes, int64_t L, const int32_t* bvar, const double* bsign,
                  const double* brhs, int32_t* status, double* z, double* x, int64_t* basis, int32_t* stats) {
    auto it = g_roots.find(root);
    if (it == g_roots.end()) return GM_ERR_BAD_HANDLE;
    EmuRoot& r = it->second;
    int rc = emu_simplex_batch((int)nodes, r.c.data(), r.A.data(), r.b.data(), 0, 0, 0, r.n0, r.m0, r.n0, (int)L, bvar,
                               bsign, brhs, nullptr, 0.0, 0, 0, status, z, x, r.n0, r.n0,
                               reinterpret_cast<long long*>(basis), stats, g_T, 0,
                               (g_reg && r.m0 + (int)L <= 64) ? 1 : 0, 0, 0);
    return rc == 0 ? GM_OK : GM_ERR_CUDA;
}
int gm_solve_wave_warm(gm_root_t root, int64_t nodes, int64_t L, const int32_t* bvar, const double* bsign,
                       const double* brhs, const int32_t* parent, int32_t* status, double* z, double* x, int64_t* basis,
                       int32_t* stats) {
    auto it = g_roots.find(root);
    if (it == g_roots.end()) return GM_ERR_BAD_HANDLE;
    EmuRoot& r = it->second;
    const int64_t m = r.m0 + L;
    std::vector<double> cur_bi((size_t)nodes * m * m, 0.0);
    std::vector<long long> cur_basis((size_t)nodes * m, -1);
    const bool can_warm = parent && L >= 1 && r.prev_nodes > 0 && r.prev_m == m - 1;
    const bool reg = g_reg && m <= 64;
    int rc = emu_simplex_batch_ex((int)nodes, r.c.data(), r.A.data(), r.b.data(), 0, 0, 0, r.n0, r.m0, r.n0, (int)L, bvar,
                                  bsign, brhs, nullptr, 0.0, 0, 0, status, z, x, r.n0, r.n0, cur_basis.data(), stats,
                                  reg ? 256 : g_T, 0, reg ? 1 : 0, 0, 0, can_warm ? parent : nullptr,
                                  can_warm ? r.prev_basis.data() : nullptr, can_warm ? r.prev_bi.data() : nullptr,
                                  cur_bi.data());
    // nodes whose warm start died are re-solved cold in place (engine.cu does the same with a second launch)
    std::vector<int> retry;
    for (int64_t i = 0; i < nodes; ++i)
        if (status[i] == GM_ERR_WARM_RETRY) retry.push_back((int)i);
    if (rc == 0 && !retry.empty()) {
        g_emu_lp_list = retry.data();
        rc = emu_simplex_batch_ex((int)retry.size(), r.c.data(), r.A.data(), r.b.data(), 0, 0, 0, r.n0, r.m0, r.n0, (int)L,
                                  bvar, bsign, brhs, nullptr, 0.0, 0, 0, status, z, x, r.n0, r.n0, cur_basis.data(), stats,
                                  reg ? 256 : g_T, 0, reg ? 1 : 0, 0, 0, nullptr, nullptr, nullptr, cur_bi.data());
        g_emu_lp_list = nullptr;
    }
    if (basis) std::memcpy(basis, cur_basis.data(), sizeof(long long) * cur_basis.size());
    r.prev_bi.swap(cur_bi);
    r.prev_basis.swap(cur_basis);
    r.prev_nodes = nodes;
    r.prev_m = (int)m;
    return rc == 0 ? GM_OK : GM_ERR_CUDA;
}
int gm_thread_robust(int delta) { g_emu_robust += delta; if (g_emu_robust < 0) g_emu_robust = 0; return g_emu_robust; }
void emu_set_robust(int on) { g_emu_robust = on ? 1 : 0; }
// the device-side scan (bnb_device.cu) is CUDA only: not available in the emulator build
int gm_milp_solve_device(int64_t, const double*, int64_t, const double*, const double*, int64_t, const double*,
                         const double*, const uint8_t*, int32_t, int32_t, int64_t, double, double*, gm_milp_result*,
                         gm_decision_cb, gm_wave_cb, void*) {
    return GM_ERR_NO_DEVICE;
}
}  // extern "C"

#include "../../gomilp_b200/csrc/bnb_host.cpp"
