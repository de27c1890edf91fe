// engine.cu — C ABI of libgomilp_b200.so (include/gomilp_b200.h): device management, host<->device
// staging, tier selection and launch of the simplex wave kernel. No CPU fallback anywhere: every
// compute entry point returns GM_ERR_NO_DEVICE when no CUDA device is usable.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gomilp_b200.h"
#include "kernels.h"

namespace {

using gm_kernels::kHbmThreads;
using gm_kernels::kSmemThreads;

struct Root {
    double *c = nullptr, *A = nullptr, *b = nullptr;
    int m0 = 0, n0 = 0;
    // warm-start state: final bases / inverses of the previous wave, kept in HBM for the children
    double* prev_bi = nullptr;
    long long* prev_basis = nullptr;
    int64_t prev_nodes = 0;
    int prev_m = 0;
};

struct Engine {
    std::mutex mu;
    bool ready = false;
    int device = -1;
    int sms = 0;
    size_t smem_optin = 0;
    gm_options opt{0, 0, 0, 0};
    std::map<gm_root_t, Root> roots;
    gm_root_t next_root = 1;
};
Engine g;

thread_local std::string t_err;
thread_local gm_timing t_timing{};

int fail(cudaError_t e, const char* what) {
    t_err = std::string(what) + ": " + cudaGetErrorString(e);
    return GM_ERR_CUDA;
}
#define CK(call)                                   \
    do {                                           \
        cudaError_t e_ = (call);                   \
        if (e_ != cudaSuccess) return fail(e_, #call); \
    } while (0)

int ensure_ready() {
    if (g.ready) {
        cudaSetDevice(g.device);  // bind the calling (possibly new) host thread
        return GM_OK;
    }
    return gm_init(0);
}

}  // namespace

namespace gm_internal {

// Launches the wave kernel over `P.count` LPs whose problem/outputs are already device-resident.
// Fills work/queue fields of P. Asynchronous on `stream`; kernel time is recorded into ev0/ev1 if given.
int launch_wave(gm::BatchParams P, cudaStream_t stream, cudaEvent_t ev0, cudaEvent_t ev1, gm_timing* tm) {
    const int m = P.m0 + P.L, n = P.n0 + P.L;
    if (P.count <= 0) return GM_OK;
    if (m <= 0 || n <= 0) return GM_ERR_BAD_SHAPE;
    P.max_pivots = g.opt.max_pivots;
    P.refactor_period = g.opt.refactor_period;
    const gm::WsLayout wr = gm::ws_layout(m, n, kSmemThreads, true);
    const gm::WsLayout w1 = gm::ws_layout(m, n, kSmemThreads, false, false, true);   // tier 2
    const gm::WsLayout w3 = gm::ws_layout(m, n, kHbmThreads, false, true, true);    // tier 3
    const gm::WsLayout w2 = gm::ws_layout(m, n, kHbmThreads, false, true);
    const size_t smem_reg = wr.big_bytes + wr.small_bytes;
    const size_t smem_all = w1.big_bytes + w1.small_bytes;
    const bool fits_reg = m <= 64 && smem_reg + 64 <= g.smem_optin;
    const bool fits_smem = smem_all + 64 <= g.smem_optin;
    const bool fits_small = w2.small_bytes + 64 <= g.smem_optin;
    const size_t bi_bytes = (w3.big_doubles - w3.Bi) * sizeof(double);
    const bool fits_bismem = bi_bytes + w3.small_bytes + 64 <= g.smem_optin;
    int tier = g.opt.force_tier;
    if (tier == 0) tier = fits_reg ? 1 : (fits_smem ? 2 : (fits_bismem ? 3 : (fits_small ? 4 : 5)));
    if ((tier == 1 && !fits_reg) || (tier == 2 && !fits_smem) || (tier == 3 && !fits_bismem) ||
        (tier == 4 && !fits_small) || tier < 1 || tier > 5)
        return GM_ERR_TOO_LARGE;

    int* queue = nullptr;
    CK(cudaMallocAsync(&queue, sizeof(int), stream));
    CK(cudaMemsetAsync(queue, 0, sizeof(int), stream));
    P.queue = queue;
    double* work = nullptr;
    int grid = 0, block = 0;
    size_t smem = 0;
    P.tier = tier;
    if (tier == 1 || tier == 2) {
        block = kSmemThreads;
        smem = tier == 1 ? smem_reg : smem_all;
        int per_sm = 0;
        if (tier == 1) CK(gm_kernels::reg_prepare(smem, &per_sm));
        else CK(gm_kernels::generic_prepare(block, smem, &per_sm));
        if (per_sm < 1) per_sm = 1;
        grid = (int)std::min<long long>(P.count, (long long)g.sms * per_sm);
        if (ev0) CK(cudaEventRecord(ev0, stream));
        if (tier == 1) gm_kernels::reg_launch(P, grid, smem, stream);
        else gm_kernels::generic_launch(P, grid, block, smem, stream);
    } else {
        block = kHbmThreads;
        P.hbm_layout = 1;
        // per-CTA HBM slice: W only (tier 3), W + Bi (tier 4), everything (tier 5)
        const size_t per_cta = tier == 3 ? w3.Bi : (tier == 4 ? w2.big_doubles : (w2.big_doubles + w2.small_bytes / 8 + 8));
        grid = (int)std::min<long long>(P.count, (long long)g.sms);
        CK(cudaMallocAsync(&work, per_cta * sizeof(double) * grid, stream));
        P.work = work;
        P.work_stride = (long long)per_cta;
        if (tier == 3) {
            smem = bi_bytes + w3.small_bytes;
        } else if (tier == 4) {
            // TMA staging ring: up to 3 stages of 32 KB if they fit beside the vectors
            const size_t stage = 32768;
            long long room = (long long)g.smem_optin - (long long)w2.small_bytes - 256;
            int ns = (int)std::min<long long>(3, room / (long long)stage);
            if (ns < 2 || g.opt.reserved == 1) ns = 0;  // options.reserved = 1 disables the ring (A/B measurements)
            P.ring_stages = ns;
            P.ring_stage_bytes = ns ? (int)stage : 0;
            P.stream_min_m = 384;
            smem = w2.small_bytes + (size_t)ns * stage;
        } else {
            smem = 0;
        }
        int per_sm = 0;
        CK(gm_kernels::generic_prepare(block, smem, &per_sm));
        if (ev0) CK(cudaEventRecord(ev0, stream));
        gm_kernels::generic_launch(P, grid, block, smem, stream);
    }
    CK(cudaGetLastError());
    if (ev1) CK(cudaEventRecord(ev1, stream));
    if (work) CK(cudaFreeAsync(work, stream));
    CK(cudaFreeAsync(queue, stream));
    if (tm) {
        tm->tier = tier;
        tm->grid = grid;
        tm->block = block;
        tm->smem_bytes = (int64_t)smem;
        tm->launches += 1;
        tm->lps += P.count;
    }
    return GM_OK;
}

int root_lookup(gm_root_t h, const double** c, const double** A, const double** b, int* m0, int* n0) {
    std::lock_guard<std::mutex> lk(g.mu);
    auto it = g.roots.find(h);
    if (it == g.roots.end()) return GM_ERR_BAD_HANDLE;
    *c = it->second.c; *A = it->second.A; *b = it->second.b;
    *m0 = it->second.m0; *n0 = it->second.n0;
    return GM_OK;
}

}  // namespace gm_internal

using gm_internal::launch_wave;

extern "C" {

int gm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* gm_last_error(void) { return t_err.c_str(); }

int gm_init(int device) {
    std::lock_guard<std::mutex> lk(g.mu);
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        t_err = "no CUDA device (the engine has no CPU fallback)";
        return GM_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) return GM_ERR_BAD_ARGUMENT;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    g.device = device;
    g.sms = prop.multiProcessorCount;
    g.smem_optin = prop.sharedMemPerBlockOptin;
    // keep stream-ordered allocations cached across calls (the default pool trims at every synchronize)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    g.ready = true;
    return GM_OK;
}

int gm_shutdown(void) {
    std::lock_guard<std::mutex> lk(g.mu);
    if (!g.ready) return GM_OK;
    for (auto& kv : g.roots) {
        cudaFree(kv.second.c);
        cudaFree(kv.second.A);
        cudaFree(kv.second.b);
        cudaFree(kv.second.prev_bi);
        cudaFree(kv.second.prev_basis);
    }
    g.roots.clear();
    g.ready = false;
    return GM_OK;
}

int gm_set_options(const gm_options* opt) {
    if (!opt) return GM_ERR_BAD_ARGUMENT;
    std::lock_guard<std::mutex> lk(g.mu);
    g.opt = *opt;
    return GM_OK;
}

int gm_last_timing(gm_timing* out) {
    if (!out) return GM_ERR_BAD_ARGUMENT;
    *out = t_timing;
    return GM_OK;
}

int gm_simplex_batch_device(int64_t count, const double* d_c, const double* d_A, const double* d_b, int64_t m,
                            int64_t n, double tol, int32_t* d_status, double* d_optF, double* d_optX,
                            int64_t* d_basis, int32_t* d_stats, void* stream) {
    int rc = ensure_ready();
    if (rc != GM_OK) return rc;
    if (count < 0 || m <= 0 || n <= 0 || count > INT32_MAX || m > (1 << 20) || n > (1 << 20)) return GM_ERR_BAD_SHAPE;
    if (!d_c || !d_A || !d_b || !d_status || !d_optF || !d_optX) return GM_ERR_BAD_ARGUMENT;
    gm::BatchParams P;
    std::memset(&P, 0, sizeof(P));
    P.c = d_c; P.A = d_A; P.b = d_b;
    P.c_stride = n; P.A_stride = m * n; P.b_stride = m;
    P.lda = (int)n; P.m0 = (int)m; P.n0 = (int)n; P.L = 0;
    P.tol = tol; P.count = (int)count;
    P.status = d_status; P.optF = d_optF; P.x = d_optX; P.x_stride = n; P.x_len = (int)n;
    P.basis = reinterpret_cast<long long*>(d_basis); P.stats = d_stats;
    t_timing = gm_timing{};
    return launch_wave(P, (cudaStream_t)stream, nullptr, nullptr, &t_timing);
}

}  // extern "C"

namespace {
struct DevBuf {
    void* p = nullptr;
    cudaStream_t s;
    explicit DevBuf(cudaStream_t s_) : s(s_) {}
    ~DevBuf() { if (p) cudaFreeAsync(p, s); }
    cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes ? bytes : 8, s); }
    template <class T> T* as() { return static_cast<T*>(p); }
};
struct StreamEvents {
    cudaStream_t s = nullptr;
    cudaEvent_t e[6] = {};
    cudaError_t init() {
        cudaError_t r = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        for (int i = 0; i < 6 && r == cudaSuccess; ++i) r = cudaEventCreate(&e[i]);
        return r;
    }
    ~StreamEvents() {
        for (auto& ev : e) if (ev) cudaEventDestroy(ev);
        if (s) cudaStreamDestroy(s);
    }
};
float ms(cudaEvent_t a, cudaEvent_t b) {
    float t = 0;
    cudaEventElapsedTime(&t, a, b);
    return t;
}
}  // namespace

// shared body of the host-buffer entry points: per-LP roots (stride != 0) or wave over a device root
static int run_host_call(gm::BatchParams P, const double* h_c, const double* h_A, int64_t h_lda, const double* h_b,
                         const int64_t* h_ib, const int32_t* h_bvar, const double* h_bsign, const double* h_brhs,
                         int32_t* status, double* optF, double* optX, int64_t* basis, int32_t* stats,
                         long long* d_basis_keep = nullptr) {
    StreamEvents se;
    CK(se.init());
    cudaStream_t st = se.s;
    const int64_t count = P.count, m0 = P.m0, n0 = P.n0, L = P.L, m = m0 + L;
    DevBuf dc(st), dA(st), db(st), dib(st), dbv(st), dbs(st), dbr(st), dst(st), dF(st), dX(st), dB(st), dS(st);
    CK(cudaEventRecord(se.e[0], st));
    if (h_A) {  // per-LP roots travel with the call
        CK(dc.alloc(sizeof(double) * count * n0));
        CK(dA.alloc(sizeof(double) * count * m0 * n0));
        CK(db.alloc(sizeof(double) * count * m0));
        CK(cudaMemcpyAsync(dc.p, h_c, sizeof(double) * count * n0, cudaMemcpyHostToDevice, st));
        if (h_lda == n0) {
            CK(cudaMemcpyAsync(dA.p, h_A, sizeof(double) * count * m0 * n0, cudaMemcpyHostToDevice, st));
        } else {  // strided single matrix (gm_simplex with lda > n)
            CK(cudaMemcpy2DAsync(dA.p, sizeof(double) * n0, h_A, sizeof(double) * h_lda, sizeof(double) * n0,
                                 count * m0, cudaMemcpyHostToDevice, st));
        }
        CK(cudaMemcpyAsync(db.p, h_b, sizeof(double) * count * m0, cudaMemcpyHostToDevice, st));
        P.c = dc.as<double>(); P.A = dA.as<double>(); P.b = db.as<double>();
        P.lda = (int)n0;
    }
    if (h_ib) {
        CK(dib.alloc(sizeof(int64_t) * count * m));
        CK(cudaMemcpyAsync(dib.p, h_ib, sizeof(int64_t) * count * m, cudaMemcpyHostToDevice, st));
        P.initial_basic = dib.as<long long>();
    }
    if (L > 0) {
        CK(dbv.alloc(sizeof(int32_t) * count * L));
        CK(dbs.alloc(sizeof(double) * count * L));
        CK(dbr.alloc(sizeof(double) * count * L));
        CK(cudaMemcpyAsync(dbv.p, h_bvar, sizeof(int32_t) * count * L, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(dbs.p, h_bsign, sizeof(double) * count * L, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(dbr.p, h_brhs, sizeof(double) * count * L, cudaMemcpyHostToDevice, st));
        P.bvar = dbv.as<int>(); P.bsign = dbs.as<double>(); P.brhs = dbr.as<double>();
    }
    CK(dst.alloc(sizeof(int32_t) * count));
    CK(dF.alloc(sizeof(double) * count));
    CK(dX.alloc(sizeof(double) * count * P.x_len));
    P.status = dst.as<int>(); P.optF = dF.as<double>(); P.x = dX.as<double>(); P.x_stride = P.x_len;
    if (d_basis_keep) P.basis = d_basis_keep;  // caller-owned device buffer (warm start keeps it for the next wave)
    else if (basis) { CK(dB.alloc(sizeof(int64_t) * count * m)); P.basis = dB.as<long long>(); }
    if (stats) { CK(dS.alloc(sizeof(int32_t) * count * 8)); P.stats = dS.as<int>(); }
    t_timing = gm_timing{};
    int rc = launch_wave(P, st, se.e[1], se.e[2], &t_timing);
    if (rc != GM_OK) { cudaStreamSynchronize(st); return rc; }
    CK(cudaMemcpyAsync(status, dst.p, sizeof(int32_t) * count, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(optF, dF.p, sizeof(double) * count, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(optX, dX.p, sizeof(double) * count * P.x_len, cudaMemcpyDeviceToHost, st));
    if (basis) CK(cudaMemcpyAsync(basis, P.basis, sizeof(int64_t) * count * m, cudaMemcpyDeviceToHost, st));
    if (stats) CK(cudaMemcpyAsync(stats, dS.p, sizeof(int32_t) * count * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(se.e[3], st));
    CK(cudaStreamSynchronize(st));
    t_timing.h2d_ms = ms(se.e[0], se.e[1]);
    t_timing.kernel_ms = ms(se.e[1], se.e[2]);
    t_timing.d2h_ms = ms(se.e[2], se.e[3]);
    return GM_OK;
}

// Large host batches: the batch is cut in chunks; chunk k+1 crosses PCIe on the copy stream while chunk k is
// being solved, and consecutive chunks are launched on alternating compute streams so that the tail of
// one launch overlaps the head of the next. Same results as one launch (LPs are independent).
static int run_host_batch_pipelined(gm::BatchParams P, const double* h_c, const double* h_A, const double* h_b,
                                    int32_t* status, double* optF, double* optX, int64_t* basis, int32_t* stats,
                                    int chunks) {
    const int64_t count = P.count, m = P.m0, n = P.n0;
    cudaStream_t cs = nullptr, ks[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev(3 * chunks + 2, nullptr);
    auto cleanup = [&]() {
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        if (cs) cudaStreamDestroy(cs);
        for (auto& k : ks) if (k) cudaStreamDestroy(k);
    };
#define CKP(call)                                                 \
    do {                                                          \
        cudaError_t e_ = (call);                                  \
        if (e_ != cudaSuccess) { cudaDeviceSynchronize(); cleanup(); return fail(e_, #call); } \
    } while (0)
    CKP(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    CKP(cudaStreamCreateWithFlags(&ks[0], cudaStreamNonBlocking));
    CKP(cudaStreamCreateWithFlags(&ks[1], cudaStreamNonBlocking));
    for (auto& e : ev) CKP(cudaEventCreate(&e));
    double *dc = nullptr, *dA = nullptr, *db = nullptr, *dF = nullptr, *dX = nullptr;
    int32_t *dst = nullptr, *dS = nullptr;
    long long* dB = nullptr;
    CKP(cudaMallocAsync(&dc, sizeof(double) * count * n, cs));
    CKP(cudaMallocAsync(&dA, sizeof(double) * count * m * n, cs));
    CKP(cudaMallocAsync(&db, sizeof(double) * count * m, cs));
    CKP(cudaMallocAsync(&dF, sizeof(double) * count, cs));
    CKP(cudaMallocAsync(&dX, sizeof(double) * count * n, cs));
    CKP(cudaMallocAsync(&dst, sizeof(int32_t) * count, cs));
    if (basis) CKP(cudaMallocAsync(&dB, sizeof(long long) * count * m, cs));
    if (stats) CKP(cudaMallocAsync(&dS, sizeof(int32_t) * count * 8, cs));
    CKP(cudaEventRecord(ev[3 * chunks], cs));  // allocations done, start of the transfer clock
    t_timing = gm_timing{};
    const int64_t per = (count + chunks - 1) / chunks;
    int rc = GM_OK;
    for (int k = 0; k < chunks && rc == GM_OK; ++k) {
        const int64_t off = k * per, cnt = std::min<int64_t>(per, count - off);
        if (cnt <= 0) break;
        CKP(cudaMemcpyAsync(dc + off * n, h_c + off * n, sizeof(double) * cnt * n, cudaMemcpyHostToDevice, cs));
        CKP(cudaMemcpyAsync(dA + off * m * n, h_A + off * m * n, sizeof(double) * cnt * m * n, cudaMemcpyHostToDevice, cs));
        CKP(cudaMemcpyAsync(db + off * m, h_b + off * m, sizeof(double) * cnt * m, cudaMemcpyHostToDevice, cs));
        CKP(cudaEventRecord(ev[3 * k], cs));
        cudaStream_t st = ks[k & 1];
        CKP(cudaStreamWaitEvent(st, ev[3 * chunks], 0));
        CKP(cudaStreamWaitEvent(st, ev[3 * k], 0));
        gm::BatchParams Q = P;
        Q.c = dc + off * n; Q.A = dA + off * m * n; Q.b = db + off * m; Q.lda = (int)n;
        Q.count = (int)cnt;
        Q.status = dst + off; Q.optF = dF + off; Q.x = dX + off * n; Q.x_stride = n;
        Q.basis = basis ? dB + off * m : nullptr;
        Q.stats = stats ? dS + off * 8 : nullptr;
        rc = launch_wave(Q, st, ev[3 * k + 1], ev[3 * k + 2], &t_timing);
        if (rc != GM_OK) break;
    }
    // results come back in one piece at the end: a device->host copy into pageable memory (a Go slice, a numpy
    // array) blocks the calling thread, which inside the loop would serialise the chunks
    CKP(cudaEventRecord(ev[3 * chunks + 1], cs));
    if (rc == GM_OK) {
        for (int k = 0; k < chunks; ++k) CKP(cudaStreamWaitEvent(cs, ev[3 * k + 2], 0));
        CKP(cudaMemcpyAsync(status, dst, sizeof(int32_t) * count, cudaMemcpyDeviceToHost, cs));
        CKP(cudaMemcpyAsync(optF, dF, sizeof(double) * count, cudaMemcpyDeviceToHost, cs));
        CKP(cudaMemcpyAsync(optX, dX, sizeof(double) * count * n, cudaMemcpyDeviceToHost, cs));
        if (basis) CKP(cudaMemcpyAsync(basis, dB, sizeof(int64_t) * count * m, cudaMemcpyDeviceToHost, cs));
        if (stats) CKP(cudaMemcpyAsync(stats, dS, sizeof(int32_t) * count * 8, cudaMemcpyDeviceToHost, cs));
    }
    CKP(cudaStreamSynchronize(cs));
    CKP(cudaStreamSynchronize(ks[0]));
    CKP(cudaStreamSynchronize(ks[1]));
    if (rc == GM_OK) {
        t_timing.h2d_ms = ms(ev[3 * chunks], ev[3 * chunks + 1]);
        double kms = 0;
        for (int k = 0; k < chunks; ++k) kms += ms(ev[3 * k + 1], ev[3 * k + 2]);
        t_timing.kernel_ms = kms;  // sum over chunk launches (they overlap the copies)
        t_timing.d2h_ms = 0;
    }
    cudaFreeAsync(dc, cs); cudaFreeAsync(dA, cs); cudaFreeAsync(db, cs); cudaFreeAsync(dF, cs);
    cudaFreeAsync(dX, cs); cudaFreeAsync(dst, cs);
    if (dB) cudaFreeAsync(dB, cs);
    if (dS) cudaFreeAsync(dS, cs);
    cudaStreamSynchronize(cs);
    cleanup();
#undef CKP
    return rc;
}

extern "C" {

int gm_simplex_batch(int64_t count, const double* c, const double* A, const double* b, int64_t m, int64_t n,
                     double tol, int32_t* status, double* optF, double* optX, int64_t* basis, int32_t* stats) {
    int rc = ensure_ready();
    if (rc != GM_OK) return rc;
    if (count < 0 || m <= 0 || n <= 0 || count > INT32_MAX || m > (1 << 20) || n > (1 << 20)) return GM_ERR_BAD_SHAPE;
    if (count == 0) return GM_OK;
    if (!c || !A || !b || !status || !optF || !optX) return GM_ERR_BAD_ARGUMENT;
    gm::BatchParams P;
    std::memset(&P, 0, sizeof(P));
    P.c_stride = n; P.A_stride = m * n; P.b_stride = m;
    P.m0 = (int)m; P.n0 = (int)n; P.L = 0; P.tol = tol; P.count = (int)count; P.x_len = (int)n;
    // pipeline when the input is big enough for PCIe time to matter and every chunk still fills the GPU
    const double in_bytes = (double)count * (double)(m * n + m + n) * 8.0;
    int chunks = (int)std::min<int64_t>(8, count / (4 * (int64_t)g.sms));
    if (in_bytes < 32e6) chunks = 1;
    if (chunks >= 2) return run_host_batch_pipelined(P, c, A, b, status, optF, optX, basis, stats, chunks);
    return run_host_call(P, c, A, n, b, nullptr, nullptr, nullptr, nullptr, status, optF, optX, basis, stats);
}

int gm_simplex(const double* c, const double* A, int64_t lda, const double* b, int64_t m, int64_t n, double tol,
               const int64_t* initialBasic, double* optF, double* optX, int64_t* basisOut, int64_t* pivots) {
    int rc = ensure_ready();
    if (rc != GM_OK) return rc;
    if (m <= 0 || n <= 0 || lda < n || m > (1 << 20) || n > (1 << 20)) return GM_ERR_BAD_SHAPE;  // simplex.go:387-398 panics
    if (!c || !A || !b || !optF || !optX) return GM_ERR_BAD_ARGUMENT;
    gm::BatchParams P;
    std::memset(&P, 0, sizeof(P));
    P.c_stride = n; P.A_stride = m * n; P.b_stride = m;
    P.m0 = (int)m; P.n0 = (int)n; P.L = 0; P.tol = tol; P.count = 1; P.x_len = (int)n;
    int32_t status = GM_ERR_CUDA;
    int32_t stats[8] = {0};
    rc = run_host_call(P, c, A, lda, b, initialBasic, nullptr, nullptr, nullptr, &status, optF, optX, basisOut, stats);
    if (rc != GM_OK) return rc;
    if (pivots) *pivots = (int64_t)stats[0] + stats[1];
    return status;
}

int gm_upload_root(const double* c0, const double* A0, int64_t lda, const double* b0, int64_t m0, int64_t n0,
                   gm_root_t* out) {
    int rc = ensure_ready();
    if (rc != GM_OK) return rc;
    if (m0 <= 0 || n0 <= 0 || lda < n0) return GM_ERR_BAD_SHAPE;
    if (!c0 || !A0 || !b0 || !out) return GM_ERR_BAD_ARGUMENT;
    Root r;
    r.m0 = (int)m0; r.n0 = (int)n0;
    CK(cudaMalloc(&r.c, sizeof(double) * n0));
    CK(cudaMalloc(&r.A, sizeof(double) * m0 * n0));
    CK(cudaMalloc(&r.b, sizeof(double) * m0));
    CK(cudaMemcpy(r.c, c0, sizeof(double) * n0, cudaMemcpyHostToDevice));
    CK(cudaMemcpy2D(r.A, sizeof(double) * n0, A0, sizeof(double) * lda, sizeof(double) * n0, m0, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(r.b, b0, sizeof(double) * m0, cudaMemcpyHostToDevice));
    std::lock_guard<std::mutex> lk(g.mu);
    *out = g.next_root++;
    g.roots[*out] = r;
    return GM_OK;
}

int gm_free_root(gm_root_t root) {
    std::lock_guard<std::mutex> lk(g.mu);
    auto it = g.roots.find(root);
    if (it == g.roots.end()) return GM_ERR_BAD_HANDLE;
    cudaFree(it->second.c);
    cudaFree(it->second.A);
    cudaFree(it->second.b);
    cudaFree(it->second.prev_bi);
    cudaFree(it->second.prev_basis);
    g.roots.erase(it);
    return GM_OK;
}

int gm_solve_wave(gm_root_t root, int64_t nodes, int64_t L, const int32_t* bvar, const double* bsign,
                  const double* brhs, int32_t* status, double* z, double* x, int64_t* basis, int32_t* stats) {
    int rc = ensure_ready();
    if (rc != GM_OK) return rc;
    if (nodes < 0 || L < 0 || nodes > INT32_MAX) return GM_ERR_BAD_SHAPE;
    if (nodes == 0) return GM_OK;
    if (!status || !z || !x || (L > 0 && (!bvar || !bsign || !brhs))) return GM_ERR_BAD_ARGUMENT;
    gm::BatchParams P;
    std::memset(&P, 0, sizeof(P));
    int m0, n0;
    rc = gm_internal::root_lookup(root, &P.c, &P.A, &P.b, &m0, &n0);
    if (rc != GM_OK) return rc;
    for (int64_t i = 0; i < nodes * L; ++i)
        if (bvar[i] < 0 || bvar[i] >= n0) return GM_ERR_BAD_ARGUMENT;
    P.m0 = m0; P.n0 = n0; P.lda = n0; P.L = (int)L; P.tol = 0.0;  // subproblem.go:154,172 pass tol = 0
    P.count = (int)nodes; P.x_len = n0;
    return run_host_call(P, nullptr, nullptr, 0, nullptr, nullptr, bvar, bsign, brhs, status, z, x, basis, stats);
}

int gm_solve_wave_warm(gm_root_t root, int64_t nodes, int64_t L, const int32_t* bvar, const double* bsign,
                       const double* brhs, const int32_t* parent, int32_t* status, double* z, double* x,
                       int64_t* basis, int32_t* stats) {
    int rc = ensure_ready();
    if (rc != GM_OK) return rc;
    if (nodes < 0 || L < 0 || nodes > INT32_MAX) return GM_ERR_BAD_SHAPE;
    if (nodes == 0) return GM_OK;
    if (!status || !z || !x || (L > 0 && (!bvar || !bsign || !brhs))) return GM_ERR_BAD_ARGUMENT;
    Root* r = nullptr;
    {
        std::lock_guard<std::mutex> lk(g.mu);
        auto it = g.roots.find(root);
        if (it == g.roots.end()) return GM_ERR_BAD_HANDLE;
        r = &it->second;
    }
    const int m0 = r->m0, n0 = r->n0;
    const int64_t m = m0 + L;
    for (int64_t i = 0; i < nodes * L; ++i)
        if (bvar[i] < 0 || bvar[i] >= n0) return GM_ERR_BAD_ARGUMENT;
    const bool can_warm = parent && L >= 1 && r->prev_bi && r->prev_m == m - 1;
    if (can_warm)
        for (int64_t i = 0; i < nodes; ++i)
            if (parent[i] >= r->prev_nodes) return GM_ERR_BAD_ARGUMENT;

    StreamEvents se;
    CK(se.init());
    cudaStream_t st = se.s;
    gm::BatchParams P;
    std::memset(&P, 0, sizeof(P));
    P.c = r->c; P.A = r->A; P.b = r->b;
    P.m0 = m0; P.n0 = n0; P.lda = n0; P.L = (int)L; P.tol = 0.0;  // subproblem.go:154,172 pass tol = 0
    P.count = (int)nodes; P.x_len = n0; P.x_stride = n0;
    // this wave's bases / inverses stay on the device for the next wave (if they fit comfortably)
    double* cur_bi = nullptr;
    long long* cur_basis = nullptr;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const size_t need = (size_t)nodes * (size_t)m * (size_t)m * 8 + (size_t)nodes * (size_t)m * 8;
    if (need < free_b / 2) {  // stream-ordered pool: no device-wide synchronisation per wave
        CK(cudaMallocAsync(&cur_bi, (size_t)nodes * m * m * 8, st));
        CK(cudaMallocAsync(&cur_basis, (size_t)nodes * m * 8, st));
    }
    DevBuf dbv(st), dbs(st), dbr(st), dpar(st), dst(st), dF(st), dX(st), dB(st), dS(st), dlist(st);
    CK(cudaEventRecord(se.e[0], st));
    if (L > 0) {
        CK(dbv.alloc(sizeof(int32_t) * nodes * L));
        CK(dbs.alloc(sizeof(double) * nodes * L));
        CK(dbr.alloc(sizeof(double) * nodes * L));
        CK(cudaMemcpyAsync(dbv.p, bvar, sizeof(int32_t) * nodes * L, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(dbs.p, bsign, sizeof(double) * nodes * L, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(dbr.p, brhs, sizeof(double) * nodes * L, cudaMemcpyHostToDevice, st));
        P.bvar = dbv.as<int>(); P.bsign = dbs.as<double>(); P.brhs = dbr.as<double>();
    }
    if (can_warm) {
        CK(dpar.alloc(sizeof(int32_t) * nodes));
        CK(cudaMemcpyAsync(dpar.p, parent, sizeof(int32_t) * nodes, cudaMemcpyHostToDevice, st));
        P.warm_parent = dpar.as<int>();
        P.warm_basis = r->prev_basis;
        P.warm_bi = r->prev_bi;
    }
    CK(dst.alloc(sizeof(int32_t) * nodes));
    CK(dF.alloc(sizeof(double) * nodes));
    CK(dX.alloc(sizeof(double) * nodes * n0));
    CK(dS.alloc(sizeof(int32_t) * nodes * 8));
    P.status = dst.as<int>(); P.optF = dF.as<double>(); P.x = dX.as<double>(); P.stats = dS.as<int>();
    if (cur_basis) P.basis = cur_basis;
    else if (basis) { CK(dB.alloc(sizeof(int64_t) * nodes * m)); P.basis = dB.as<long long>(); }
    P.bi_out = cur_bi;
    t_timing = gm_timing{};
    rc = launch_wave(P, st, se.e[1], se.e[2], &t_timing);
    if (rc == GM_OK) {
        CK(cudaMemcpyAsync(status, dst.p, sizeof(int32_t) * nodes, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        t_timing.kernel_ms = ms(se.e[1], se.e[2]);
        // nodes whose warm start died are re-solved from scratch, in place, by a second launch
        std::vector<int> retry;
        for (int64_t i = 0; i < nodes; ++i)
            if (status[i] == GM_ERR_WARM_RETRY) retry.push_back((int)i);
        if (!retry.empty()) {
            CK(dlist.alloc(sizeof(int) * retry.size()));
            CK(cudaMemcpyAsync(dlist.p, retry.data(), sizeof(int) * retry.size(), cudaMemcpyHostToDevice, st));
            gm::BatchParams Q = P;
            Q.warm_parent = nullptr;
            Q.lp_list = dlist.as<int>();
            Q.count = (int)retry.size();
            rc = launch_wave(Q, st, se.e[3], se.e[4], &t_timing);
            if (rc == GM_OK) {
                CK(cudaMemcpyAsync(status, dst.p, sizeof(int32_t) * nodes, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                t_timing.kernel_ms += ms(se.e[3], se.e[4]);
            }
        }
    }
    if (rc == GM_OK) {
        CK(cudaMemcpyAsync(z, dF.p, sizeof(double) * nodes, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(x, dX.p, sizeof(double) * nodes * n0, cudaMemcpyDeviceToHost, st));
        if (basis) CK(cudaMemcpyAsync(basis, P.basis, sizeof(int64_t) * nodes * m, cudaMemcpyDeviceToHost, st));
        if (stats) CK(cudaMemcpyAsync(stats, dS.p, sizeof(int32_t) * nodes * 8, cudaMemcpyDeviceToHost, st));
    }
    if (r->prev_bi) cudaFreeAsync(r->prev_bi, st);
    if (r->prev_basis) cudaFreeAsync(r->prev_basis, st);
    cudaStreamSynchronize(st);
    r->prev_bi = cur_bi;
    r->prev_basis = cur_basis;
    r->prev_nodes = cur_bi ? nodes : 0;
    r->prev_m = (int)m;
    return rc;
}

}  // extern "C"
