// engine.cu — C ABI of libgomilp_b200.so (include/gomilp_b200.h): device management, host<->device
// staging, tier selection and launch of the simplex wave kernels. No CPU fallback anywhere: every
// compute entry point returns GM_ERR_NO_DEVICE when no CUDA device is usable.
//
// Threading / devices (SURVEY.md 8b "Threading"): the engine keeps one context PER DEVICE. A host thread is bound
// to the device of its last gm_init(device) call (thread-local); threads that never called gm_init use the first
// device any thread initialised. Roots remember their device, so a wave always runs where its root lives. Every
// call takes its stream and events from a per-device pool; nothing global is written on the launch path except
// under the engine mutex (options snapshot, root table).
#include "engine.h"

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace gm_engine {

Engine g;
thread_local int t_device = -1;
thread_local std::string t_err;
thread_local gm_timing t_timing{};
thread_local TraceReq t_trace;
thread_local int t_robust = 0;

int fail(cudaError_t e, const char* what) {
    t_err = std::string(what) + ": " + cudaGetErrorString(e);
    return GM_ERR_CUDA;
}

cudaError_t StreamEvents::init() {
    cudaError_t r = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    for (int i = 0; i < 6 && r == cudaSuccess; ++i) r = cudaEventCreate(&e[i]);
    if (r == cudaSuccess) r = cudaMallocHost(&h_counts, 64 * sizeof(int));
    return r;
}
StreamEvents::~StreamEvents() {
    if (arena && arena_free) arena_free(arena);  // frees on `s`: before the stream goes
    if (h_counts) cudaFreeHost(h_counts);
    for (auto& ev : e)
        if (ev) cudaEventDestroy(ev);
    if (s) cudaStreamDestroy(s);
}

// A stream + 6 events for the duration of one call, recycled through the device's pool.
StreamLease::StreamLease(DeviceCtx* d) : dev(d) {
    {
        std::lock_guard<std::mutex> lk(dev->mu);
        if (!dev->pool.empty()) {
            se = dev->pool.back();
            dev->pool.pop_back();
        }
    }
    if (!se) {
        se = new StreamEvents();
        err = se->init();
    }
}
StreamLease::~StreamLease() {
    if (!se) return;
    if (err != cudaSuccess) { delete se; return; }
    std::lock_guard<std::mutex> lk(dev->mu);
    dev->pool.push_back(se);
}

static int init_device(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        t_err = "no CUDA device (the engine has no CPU fallback)";
        return GM_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n || device >= kMaxDevices) return GM_ERR_BAD_ARGUMENT;
    std::lock_guard<std::mutex> lk(g.mu);
    DeviceCtx& d = g.dev[device];
    CK(cudaSetDevice(device));
    if (!d.ready) {
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        d.id = device;
        d.sms = prop.multiProcessorCount;
        d.smem_optin = prop.sharedMemPerBlockOptin;
        d.coop_ok = prop.cooperativeLaunch != 0;
        // dynamic shared-memory limits are a per-device function attribute: raised once here, never per launch
        CK(gm_kernels::reg_set_smem_limit(d.smem_optin));
        CK(gm_kernels::generic_set_smem_limit(d.smem_optin));
        CK(gm_kernels::coop_prepare(d.smem_optin));
        CK(gm_bnb_kernels_prepare());
        // keep stream-ordered allocations cached across calls (the default pool trims at every synchronize)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        d.ready = true;
    }
    if (g.default_device < 0) g.default_device = device;
    return GM_OK;
}

// Binds the calling thread to its device (see the header comment) and returns that device's context.
int current_device(DeviceCtx** out) {
    int dev = t_device;
    if (dev < 0) {
        std::lock_guard<std::mutex> lk(g.mu);
        dev = g.default_device;
    }
    if (dev < 0) dev = 0;
    if (!g.dev[dev].ready) {
        const int rc = init_device(dev);
        if (rc != GM_OK) return rc;
    }
    CK(cudaSetDevice(dev));
    *out = &g.dev[dev];
    return GM_OK;
}

int device_of_root(gm_root_t h, Root* out, DeviceCtx** dev) {
    {
        std::lock_guard<std::mutex> lk(g.mu);
        auto it = g.roots.find(h);
        if (it == g.roots.end()) return GM_ERR_BAD_HANDLE;
        *out = it->second;
    }
    CK(cudaSetDevice(out->device));
    *dev = &g.dev[out->device];
    return GM_OK;
}

// Launches the wave kernel over `P.count` LPs whose problem/outputs are already device-resident.
// Fills work/queue fields of P. Asynchronous on `stream`; kernel time is recorded into ev0/ev1 if given.
int launch_wave(DeviceCtx& d, gm::BatchParams P, cudaStream_t stream, cudaEvent_t ev0, cudaEvent_t ev1, gm_timing* tm) {
    using gm_kernels::kHbmThreads;
    using gm_kernels::kSmemThreads;
    const int m = P.m0 + P.L, n = P.n0 + P.L;
    if (P.count <= 0) return GM_OK;
    if (m <= 0 || n <= 0) return GM_ERR_BAD_SHAPE;
    gm_options opt;
    {
        std::lock_guard<std::mutex> lk(g.mu);
        opt = g.opt;
    }
    P.max_pivots = opt.max_pivots;
    P.refactor_period = opt.refactor_period;
    P.robust = (opt.robust != 0 || t_robust > 0) ? 1 : 0;
    const gm::WsLayout wr = gm::ws_layout(m, n, kSmemThreads, true);
    const gm::WsLayout w1 = gm::ws_layout(m, n, kSmemThreads, false, false, true);   // tier 2
    const gm::WsLayout w3 = gm::ws_layout(m, n, kHbmThreads, false, true, true);    // tier 3
    const gm::WsLayout w2 = gm::ws_layout(m, n, kHbmThreads, false, true);
    const size_t smem_reg = wr.big_bytes + wr.small_bytes;
    const size_t smem_all = w1.big_bytes + w1.small_bytes;
    const bool fits_reg = m <= 64 && smem_reg + 256 <= d.smem_optin;
    const bool fits_smem = smem_all + 256 <= d.smem_optin;
    const bool fits_small = w2.small_bytes + 256 <= d.smem_optin;
    const size_t bi_bytes = (w3.big_doubles - w3.Bi) * sizeof(double);
    const bool fits_bismem = bi_bytes + w3.small_bytes + 256 <= d.smem_optin;
    // tier 6 (cooperative): fewer LPs than SMs and an LP big enough that one CTA per LP would crawl
    int G = 0;
    if (d.coop_ok && m < n) {
        G = opt.coop_group > 0 ? opt.coop_group : std::max(1, d.sms / P.count);
        G = std::min(G, std::max(1, m / 2));
        G = std::max(1, std::min(G, d.sms));
    }
    const gm::CoopLayout cl = gm::coop_layout(m, n, kHbmThreads, G > 0 ? G : 1, d.smem_optin - 512);
    const bool fits_coop = G >= 1 && cl.smem_bytes + 256 <= d.smem_optin;
    int tier = opt.force_tier;
    if (tier == 0) {
        tier = fits_reg ? 1 : (fits_smem ? 2 : (fits_bismem ? 3 : (fits_small ? 4 : 5)));
        // more SMs than LPs: several CTAs per LP. HBM-resident shapes: always (with one CTA per LP its fused
        // update + FTRAN pass still moves 2 m^2 instead of 3 m^2 words per pivot and beats the TMA-ring tier 4).
        if (fits_coop && ((tier >= 2 && G >= 2 && m >= 96) || tier >= 4)) tier = 6;
    }
    if ((tier == 1 && !fits_reg) || (tier == 2 && !fits_smem) || (tier == 3 && !fits_bismem) ||
        (tier == 4 && !fits_small) || (tier == 6 && !fits_coop) || tier < 1 || tier > 6)
        return GM_ERR_TOO_LARGE;

    // pivot trace requested for this call (gm_trace_arm)
    int* d_trace = nullptr;
    if (t_trace.armed && !t_trace.in_flight) {
        CK(cudaMallocAsync(&d_trace, sizeof(int) * 4 * (size_t)t_trace.cap, stream));
        CK(cudaMemsetAsync(d_trace, 0xff, sizeof(int) * 4 * (size_t)t_trace.cap, stream));
        P.trace = d_trace;
        P.trace_cap = (int)t_trace.cap;
        P.trace_lp = (int)t_trace.lp;
        t_trace.in_flight = true;
        t_trace.d_buf = d_trace;
        t_trace.stream = stream;
    }

    int* queue = nullptr;
    const size_t qbytes = 16 + 8 * (size_t)std::max(1, P.count);  // queue counter, then one barrier counter per group
    CK(cudaMallocAsync(&queue, qbytes, stream));
    CK(cudaMemsetAsync(queue, 0, qbytes, stream));
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
        grid = (int)std::min<long long>(P.count, (long long)d.sms * per_sm);
        if (ev0) CK(cudaEventRecord(ev0, stream));
        if (tier == 1) gm_kernels::reg_launch(P, grid, smem, stream);
        else gm_kernels::generic_launch(P, grid, block, smem, stream);
    } else if (tier == 6) {
        block = kHbmThreads;
        P.hbm_layout = 1;
        P.coop_G = G;
        P.coop_pan = cl.pan_nb;
        P.coop_small = cl.small_in_smem;
        if (t_trace.prof_armed && !t_trace.prof_in_flight) {  // leader clock cycles per activity (gm_profile_arm)
            CK(cudaMallocAsync(&t_trace.d_prof, sizeof(long long) * 16 * (size_t)P.count, stream));
            CK(cudaMemsetAsync(t_trace.d_prof, 0, sizeof(long long) * 16 * (size_t)P.count, stream));
            P.prof = t_trace.d_prof;
            t_trace.prof_count = P.count;
            t_trace.prof_in_flight = true;
            t_trace.prof_stream = stream;
        }
        const int groups = std::min(P.count, d.sms / G);
        grid = groups * G;
        smem = cl.smem_bytes;
        int per_sm = 0;
        CK(gm_kernels::coop_occupancy(block, smem, &per_sm));
        if (per_sm < 1) return GM_ERR_TOO_LARGE;
        CK(cudaMallocAsync(&work, cl.group_doubles * sizeof(double) * groups, stream));
        P.work = work;
        P.work_stride = (long long)cl.group_doubles;
        // the group barrier counters live behind the queue counter (zeroed above), 8-byte aligned
        P.coop_bar = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(queue) + 8);
        if (ev0) CK(cudaEventRecord(ev0, stream));
        CK(gm_kernels::coop_launch(P, grid, block, smem, stream));
    } else {
        block = kHbmThreads;
        P.hbm_layout = 1;
        // per-CTA HBM slice: W only (tier 3), W + Bi (tier 4), everything (tier 5)
        const size_t per_cta = tier == 3 ? w3.Bi : (tier == 4 ? w2.big_doubles : (w2.big_doubles + w2.small_bytes / 8 + 8));
        grid = (int)std::min<long long>(P.count, (long long)d.sms);
        CK(cudaMallocAsync(&work, per_cta * sizeof(double) * grid, stream));
        P.work = work;
        P.work_stride = (long long)per_cta;
        if (tier == 3) {
            smem = bi_bytes + w3.small_bytes;
        } else if (tier == 4) {
            // TMA staging ring: up to 3 stages of 32 KB if they fit beside the vectors
            const size_t stage = 32768;
            long long room = (long long)d.smem_optin - (long long)w2.small_bytes - 256;
            int ns = (int)std::min<long long>(3, room / (long long)stage);
            if (ns < 2 || opt.reserved == 1) ns = 0;  // options.reserved = 1 disables the ring (A/B measurements)
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

// Called by the host-buffer entry points once their stream has been synchronised: brings an armed trace back.
void finish_trace() {
    if (t_trace.prof_in_flight) {
        t_trace.prof.assign((size_t)t_trace.prof_count * 16, 0);
        cudaMemcpyAsync(t_trace.prof.data(), t_trace.d_prof, sizeof(long long) * 16 * (size_t)t_trace.prof_count,
                        cudaMemcpyDeviceToHost, t_trace.prof_stream);
        cudaStreamSynchronize(t_trace.prof_stream);
        cudaFreeAsync(t_trace.d_prof, t_trace.prof_stream);
        t_trace.prof_in_flight = false;
        t_trace.prof_armed = false;
        t_trace.d_prof = nullptr;
    }
    if (!t_trace.in_flight) return;
    t_trace.rows.assign((size_t)t_trace.cap * 4, -1);
    cudaMemcpyAsync(t_trace.rows.data(), t_trace.d_buf, sizeof(int) * 4 * (size_t)t_trace.cap, cudaMemcpyDeviceToHost,
                    t_trace.stream);
    cudaStreamSynchronize(t_trace.stream);
    cudaFreeAsync(t_trace.d_buf, t_trace.stream);
    t_trace.in_flight = false;
    t_trace.armed = false;
    t_trace.d_buf = nullptr;
}

float ms(cudaEvent_t a, cudaEvent_t b) {
    float t = 0;
    cudaEventElapsedTime(&t, a, b);
    return t;
}

}  // namespace gm_engine

using namespace gm_engine;

extern "C" {

int gm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* gm_last_error(void) { return t_err.c_str(); }

int gm_init(int device) {
    const int rc = init_device(device);
    if (rc == GM_OK) t_device = device;
    return rc;
}

int gm_shutdown(void) {
    std::lock_guard<std::mutex> lk(g.mu);
    for (auto& kv : g.roots) {
        cudaSetDevice(kv.second.device);
        cudaFree(kv.second.c);
        cudaFree(kv.second.A);
        cudaFree(kv.second.b);
        cudaFree(kv.second.prev_bi);
        cudaFree(kv.second.prev_basis);
    }
    g.roots.clear();
    for (int i = 0; i < kMaxDevices; ++i) {
        DeviceCtx& d = g.dev[i];
        if (!d.ready) continue;
        cudaSetDevice(i);
        std::lock_guard<std::mutex> lk2(d.mu);
        for (auto* se : d.pool) delete se;
        d.pool.clear();
        d.ready = false;
    }
    g.default_device = -1;
    t_device = -1;
    return GM_OK;
}

int gm_set_options(const gm_options* opt) {
    if (!opt) return GM_ERR_BAD_ARGUMENT;
    std::lock_guard<std::mutex> lk(g.mu);
    g.opt = *opt;
    return GM_OK;
}

int gm_thread_robust(int delta) {
    t_robust += delta;
    if (t_robust < 0) t_robust = 0;
    return t_robust;
}

int gm_last_timing(gm_timing* out) {
    if (!out) return GM_ERR_BAD_ARGUMENT;
    *out = t_timing;
    return GM_OK;
}

int gm_trace_arm(int64_t lp_index, int64_t cap) {
    if (lp_index < 0 || cap <= 0 || cap > (1 << 22)) return GM_ERR_BAD_ARGUMENT;
    t_trace.armed = true;
    t_trace.in_flight = false;
    t_trace.lp = lp_index;
    t_trace.cap = cap;
    t_trace.rows.clear();
    return GM_OK;
}

int gm_profile_arm(void) {
    t_trace.prof_armed = true;
    t_trace.prof_in_flight = false;
    t_trace.prof.clear();
    return GM_OK;
}

int64_t gm_profile_fetch(int64_t* out, int64_t lps) {
    if (!out) return 0;
    const int64_t k = std::min<int64_t>(lps, (int64_t)t_trace.prof.size() / 16);
    std::memcpy(out, t_trace.prof.data(), sizeof(int64_t) * 16 * (size_t)k);
    return k;
}

int64_t gm_trace_fetch(int32_t* out, int64_t cap) {
    if (!out) return 0;
    const int64_t rows = std::min<int64_t>(cap, (int64_t)t_trace.rows.size() / 4);
    std::memcpy(out, t_trace.rows.data(), sizeof(int32_t) * 4 * (size_t)rows);
    return rows;
}

int gm_simplex_batch_device(int64_t count, const double* d_c, const double* d_A, const double* d_b, int64_t m,
                            int64_t n, double tol, int32_t* d_status, double* d_optF, double* d_optX,
                            int64_t* d_basis, int32_t* d_stats, void* stream) {
    DeviceCtx* d = nullptr;
    int rc = current_device(&d);
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
    const bool was_armed = t_trace.armed;
    t_trace.armed = false;  // asynchronous entry point: a trace cannot be brought back
    rc = launch_wave(*d, P, (cudaStream_t)stream, nullptr, nullptr, &t_timing);
    t_trace.armed = was_armed;
    return rc;
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
    void* release() { void* q = p; p = nullptr; return q; }
};
}  // namespace

// shared body of the host-buffer entry points: per-LP roots (stride != 0) or wave over a device root
static int run_host_call(DeviceCtx& d, gm::BatchParams P, const double* h_c, const double* h_A, int64_t h_lda,
                         const double* h_b, const int64_t* h_ib, const int32_t* h_bvar, const double* h_bsign,
                         const double* h_brhs, int32_t* status, double* optF, double* optX, int64_t* basis,
                         int32_t* stats) {
    StreamLease lease(&d);
    CK(lease.err);
    StreamEvents& se = *lease.se;
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
    if (basis) { CK(dB.alloc(sizeof(int64_t) * count * m)); P.basis = dB.as<long long>(); }
    if (stats) { CK(dS.alloc(sizeof(int32_t) * count * 8)); P.stats = dS.as<int>(); }
    t_timing = gm_timing{};
    int rc = launch_wave(d, P, st, se.e[1], se.e[2], &t_timing);
    if (rc != GM_OK) { cudaStreamSynchronize(st); finish_trace(); return rc; }
    CK(cudaMemcpyAsync(status, dst.p, sizeof(int32_t) * count, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(optF, dF.p, sizeof(double) * count, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(optX, dX.p, sizeof(double) * count * P.x_len, cudaMemcpyDeviceToHost, st));
    if (basis) CK(cudaMemcpyAsync(basis, P.basis, sizeof(int64_t) * count * m, cudaMemcpyDeviceToHost, st));
    if (stats) CK(cudaMemcpyAsync(stats, dS.p, sizeof(int32_t) * count * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(se.e[3], st));
    CK(cudaStreamSynchronize(st));
    finish_trace();
    t_timing.h2d_ms = ms(se.e[0], se.e[1]);
    t_timing.kernel_ms = ms(se.e[1], se.e[2]);
    t_timing.d2h_ms = ms(se.e[2], se.e[3]);
    return GM_OK;
}

// Large host batches: the batch is cut in chunks; chunk k+1 crosses PCIe on the copy stream while chunk k is
// being solved, and consecutive chunks are launched on alternating compute streams so that the tail of
// one launch overlaps the head of the next. Same results as one launch (LPs are independent).
static int run_host_batch_pipelined(DeviceCtx& d, gm::BatchParams P, const double* h_c, const double* h_A,
                                    const double* h_b, int32_t* status, double* optF, double* optX, int64_t* basis,
                                    int32_t* stats, int chunks) {
    const int64_t count = P.count, m = P.m0, n = P.n0;
    StreamLease l0(&d), l1(&d), l2(&d);
    CK(l0.err); CK(l1.err); CK(l2.err);
    cudaStream_t cs = l0.se->s, ks[2] = {l1.se->s, l2.se->s};
    std::vector<cudaEvent_t> ev(3 * chunks + 2, nullptr);
    auto cleanup = [&]() {
        for (auto& e : ev) if (e) cudaEventDestroy(e);
    };
#define CKP(call)                                                 \
    do {                                                          \
        cudaError_t e_ = (call);                                  \
        if (e_ != cudaSuccess) { cudaDeviceSynchronize(); cleanup(); return fail(e_, #call); } \
    } while (0)
    for (auto& e : ev) CKP(cudaEventCreateWithFlags(&e, cudaEventDefault));
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
        rc = launch_wave(d, Q, st, ev[3 * k + 1], ev[3 * k + 2], &t_timing);
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

// Large PINNED host batches: ONE launch. The kernel starts at once and its CTAs pull LPs from the queue, but an LP is
// only started when the copy stream's arrival counter says its inputs are in HBM (BatchParams::ready); the batch crosses
// PCIe in slices on a second stream while the kernel is already solving the first ones. Compared with one launch per
// slice (run_host_batch_pipelined) there is no under-filled tail per slice: the time is max(copy, solve) plus the
// results' trip back.
static int run_host_batch_streamed(DeviceCtx& d, gm::BatchParams P, const double* h_c, const double* h_A,
                                   const double* h_b, int32_t* status, double* optF, double* optX, int64_t* basis,
                                   int32_t* stats) {
    const int64_t count = P.count, m = P.m0, n = P.n0;
    StreamLease lk(&d), lc(&d);
    CK(lk.err); CK(lc.err);
    cudaStream_t ks = lk.se->s, cs = lc.se->s;
    DevBuf dc(ks), dA(ks), db(ks), dF(ks), dX(ks), dst(ks), dB(ks), dS(ks), dready(ks);
    CK(dc.alloc(sizeof(double) * count * n));
    CK(dA.alloc(sizeof(double) * count * m * n));
    CK(db.alloc(sizeof(double) * count * m));
    CK(dF.alloc(sizeof(double) * count));
    CK(dX.alloc(sizeof(double) * count * n));
    CK(dst.alloc(sizeof(int32_t) * count));
    if (basis) CK(dB.alloc(sizeof(long long) * count * m));
    if (stats) CK(dS.alloc(sizeof(int32_t) * count * 8));
    CK(dready.alloc(sizeof(int)));
    CK(cudaMemsetAsync(dready.p, 0, sizeof(int), ks));
    CK(cudaEventRecord(lk.se->e[0], ks));         // buffers exist, counter is zero
    CK(cudaStreamWaitEvent(cs, lk.se->e[0], 0));
    P.c = dc.as<double>(); P.A = dA.as<double>(); P.b = db.as<double>(); P.lda = (int)n;
    P.status = dst.as<int>(); P.optF = dF.as<double>(); P.x = dX.as<double>(); P.x_stride = n;
    P.basis = basis ? dB.as<long long>() : nullptr;
    P.stats = stats ? dS.as<int>() : nullptr;
    P.ready = dready.as<int>();
    t_timing = gm_timing{};
    int rc = launch_wave(d, P, ks, lk.se->e[1], lk.se->e[2], &t_timing);   // spins on `ready` until data arrives
    // slices of the batch, then the new arrival count, on the copy stream (pinned memory: truly asynchronous)
    // c and b are small: they cross first, in one piece each; A follows in slices, each followed by the new count
    const int slices = (int)std::min<int64_t>(24, std::max<int64_t>(1, count / 160));
    const int64_t per = (count + slices - 1) / slices;
    int* hc = lc.se->h_counts;
    CK(cudaEventRecord(lc.se->e[0], cs));
    CK(cudaMemcpyAsync(dc.p, h_c, sizeof(double) * count * n, cudaMemcpyHostToDevice, cs));
    CK(cudaMemcpyAsync(db.p, h_b, sizeof(double) * count * m, cudaMemcpyHostToDevice, cs));
    for (int k = 0; k < slices && rc == GM_OK; ++k) {
        const int64_t off = k * per, cnt = std::min<int64_t>(per, count - off);
        if (cnt <= 0) break;
        CK(cudaMemcpyAsync(dA.as<double>() + off * m * n, h_A + off * m * n, sizeof(double) * cnt * m * n, cudaMemcpyHostToDevice, cs));
        hc[k] = (int)(off + cnt);
        CK(cudaMemcpyAsync(dready.p, &hc[k], sizeof(int), cudaMemcpyHostToDevice, cs));
    }
    if (rc != GM_OK) {  // the launch failed: make sure nothing is left waiting, then report
        hc[63] = (int)count;
        cudaMemcpyAsync(dready.p, &hc[63], sizeof(int), cudaMemcpyHostToDevice, cs);
        cudaStreamSynchronize(cs);
        cudaStreamSynchronize(ks);
        return rc;
    }
    CK(cudaEventRecord(lc.se->e[1], cs));
    CK(cudaMemcpyAsync(status, dst.p, sizeof(int32_t) * count, cudaMemcpyDeviceToHost, ks));
    CK(cudaMemcpyAsync(optF, dF.p, sizeof(double) * count, cudaMemcpyDeviceToHost, ks));
    CK(cudaMemcpyAsync(optX, dX.p, sizeof(double) * count * n, cudaMemcpyDeviceToHost, ks));
    if (basis) CK(cudaMemcpyAsync(basis, dB.p, sizeof(int64_t) * count * m, cudaMemcpyDeviceToHost, ks));
    if (stats) CK(cudaMemcpyAsync(stats, dS.p, sizeof(int32_t) * count * 8, cudaMemcpyDeviceToHost, ks));
    CK(cudaEventRecord(lk.se->e[3], ks));
    CK(cudaStreamSynchronize(cs));
    CK(cudaStreamSynchronize(ks));
    t_timing.h2d_ms = ms(lc.se->e[0], lc.se->e[1]);
    t_timing.kernel_ms = ms(lk.se->e[1], lk.se->e[2]);  // includes waiting for arrivals
    t_timing.d2h_ms = ms(lk.se->e[2], lk.se->e[3]);
    return GM_OK;
}

extern "C" {

int gm_simplex_batch(int64_t count, const double* c, const double* A, const double* b, int64_t m, int64_t n,
                     double tol, int32_t* status, double* optF, double* optX, int64_t* basis, int32_t* stats) {
    DeviceCtx* d = nullptr;
    int rc = current_device(&d);
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
    int chunks = (int)std::min<int64_t>(8, count / (4 * (int64_t)d->sms));
    if (in_bytes < 32e6 || t_trace.armed) chunks = 1;
    if (chunks >= 2) {
        // pinned inputs (cudaHostAlloc / cudaHostRegister / torch pin_memory): one launch gated on arrival counters
        cudaPointerAttributes pa, pb, pc;
        const bool pinned = g.opt.reserved2 != 1 &&
                            cudaPointerGetAttributes(&pa, A) == cudaSuccess && pa.type == cudaMemoryTypeHost &&
                            cudaPointerGetAttributes(&pb, b) == cudaSuccess && pb.type == cudaMemoryTypeHost &&
                            cudaPointerGetAttributes(&pc, c) == cudaSuccess && pc.type == cudaMemoryTypeHost;
        cudaGetLastError();
        if (pinned) return run_host_batch_streamed(*d, P, c, A, b, status, optF, optX, basis, stats);
        return run_host_batch_pipelined(*d, P, c, A, b, status, optF, optX, basis, stats, chunks);
    }
    return run_host_call(*d, P, c, A, n, b, nullptr, nullptr, nullptr, nullptr, status, optF, optX, basis, stats);
}

int gm_simplex(const double* c, const double* A, int64_t lda, const double* b, int64_t m, int64_t n, double tol,
               const int64_t* initialBasic, double* optF, double* optX, int64_t* basisOut, int64_t* pivots) {
    DeviceCtx* d = nullptr;
    int rc = current_device(&d);
    if (rc != GM_OK) return rc;
    if (m <= 0 || n <= 0 || lda < n || m > (1 << 20) || n > (1 << 20)) return GM_ERR_BAD_SHAPE;  // simplex.go:387-398 panics
    if (!c || !A || !b || !optF || !optX) return GM_ERR_BAD_ARGUMENT;
    gm::BatchParams P;
    std::memset(&P, 0, sizeof(P));
    P.c_stride = n; P.A_stride = m * n; P.b_stride = m;
    P.m0 = (int)m; P.n0 = (int)n; P.L = 0; P.tol = tol; P.count = 1; P.x_len = (int)n;
    int32_t status = GM_ERR_CUDA;
    int32_t stats[8] = {0};
    rc = run_host_call(*d, P, c, A, lda, b, initialBasic, nullptr, nullptr, nullptr, &status, optF, optX, basisOut, stats);
    if (rc != GM_OK) return rc;
    if (pivots) *pivots = (int64_t)stats[0] + stats[1];
    return status;
}

int gm_upload_root(const double* c0, const double* A0, int64_t lda, const double* b0, int64_t m0, int64_t n0,
                   gm_root_t* out) {
    DeviceCtx* d = nullptr;
    int rc = current_device(&d);
    if (rc != GM_OK) return rc;
    if (m0 <= 0 || n0 <= 0 || lda < n0) return GM_ERR_BAD_SHAPE;
    if (!c0 || !A0 || !b0 || !out) return GM_ERR_BAD_ARGUMENT;
    StreamLease lease(d);
    CK(lease.err);
    cudaStream_t st = lease.se->s;
    DevBuf dc(st), dA(st), db(st);
    CK(dc.alloc(sizeof(double) * n0));
    CK(dA.alloc(sizeof(double) * m0 * n0));
    CK(db.alloc(sizeof(double) * m0));
    CK(cudaMemcpyAsync(dc.p, c0, sizeof(double) * n0, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpy2DAsync(dA.p, sizeof(double) * n0, A0, sizeof(double) * lda, sizeof(double) * n0, m0,
                         cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(db.p, b0, sizeof(double) * m0, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));  // the caller's buffers are not retained past the return (cgo pointer rule)
    Root r;
    r.device = d->id;
    r.m0 = (int)m0; r.n0 = (int)n0;
    r.c = static_cast<double*>(dc.release());
    r.A = static_cast<double*>(dA.release());
    r.b = static_cast<double*>(db.release());
    std::lock_guard<std::mutex> lk(g.mu);
    *out = g.next_root++;
    g.roots[*out] = r;
    return GM_OK;
}

int gm_free_root(gm_root_t root) {
    Root r;
    {
        std::lock_guard<std::mutex> lk(g.mu);
        auto it = g.roots.find(root);
        if (it == g.roots.end()) return GM_ERR_BAD_HANDLE;
        r = it->second;
        g.roots.erase(it);
    }
    cudaSetDevice(r.device);
    cudaFree(r.c);
    cudaFree(r.A);
    cudaFree(r.b);
    cudaFree(r.prev_bi);
    cudaFree(r.prev_basis);
    if (t_device >= 0) cudaSetDevice(t_device);
    return GM_OK;
}

int gm_solve_wave(gm_root_t root, int64_t nodes, int64_t L, const int32_t* bvar, const double* bsign,
                  const double* brhs, int32_t* status, double* z, double* x, int64_t* basis, int32_t* stats) {
    if (nodes < 0 || L < 0 || nodes > INT32_MAX) return GM_ERR_BAD_SHAPE;
    if (nodes == 0) return GM_OK;
    if (!status || !z || !x || (L > 0 && (!bvar || !bsign || !brhs))) return GM_ERR_BAD_ARGUMENT;
    Root r;
    DeviceCtx* d = nullptr;
    int rc = device_of_root(root, &r, &d);
    if (rc != GM_OK) return rc;
    gm::BatchParams P;
    std::memset(&P, 0, sizeof(P));
    P.c = r.c; P.A = r.A; P.b = r.b;
    const int m0 = r.m0, n0 = r.n0;
    for (int64_t i = 0; i < nodes * L; ++i)
        if (bvar[i] < 0 || bvar[i] >= n0) return GM_ERR_BAD_ARGUMENT;
    P.m0 = m0; P.n0 = n0; P.lda = n0; P.L = (int)L; P.tol = 0.0;  // subproblem.go:154,172 pass tol = 0
    P.count = (int)nodes; P.x_len = n0;
    return run_host_call(*d, P, nullptr, nullptr, 0, nullptr, nullptr, bvar, bsign, brhs, status, z, x, basis, stats);
}

int gm_solve_wave_warm(gm_root_t root, int64_t nodes, int64_t L, const int32_t* bvar, const double* bsign,
                       const double* brhs, const int32_t* parent, int32_t* status, double* z, double* x,
                       int64_t* basis, int32_t* stats) {
    if (nodes < 0 || L < 0 || nodes > INT32_MAX) return GM_ERR_BAD_SHAPE;
    if (nodes == 0) return GM_OK;
    if (!status || !z || !x || (L > 0 && (!bvar || !bsign || !brhs))) return GM_ERR_BAD_ARGUMENT;
    Root r;
    DeviceCtx* d = nullptr;
    int rc = device_of_root(root, &r, &d);
    if (rc != GM_OK) return rc;
    const int m0 = r.m0, n0 = r.n0;
    const int64_t m = m0 + L;
    for (int64_t i = 0; i < nodes * L; ++i)
        if (bvar[i] < 0 || bvar[i] >= n0) return GM_ERR_BAD_ARGUMENT;
    const bool can_warm = parent && L >= 1 && r.prev_bi && r.prev_m == m - 1;
    if (can_warm)
        for (int64_t i = 0; i < nodes; ++i)
            if (parent[i] >= r.prev_nodes) return GM_ERR_BAD_ARGUMENT;

    StreamLease lease(d);
    CK(lease.err);
    StreamEvents& se = *lease.se;
    cudaStream_t st = se.s;
    gm::BatchParams P;
    std::memset(&P, 0, sizeof(P));
    P.c = r.c; P.A = r.A; P.b = r.b;
    P.m0 = m0; P.n0 = n0; P.lda = n0; P.L = (int)L; P.tol = 0.0;  // subproblem.go:154,172 pass tol = 0
    P.count = (int)nodes; P.x_len = n0; P.x_stride = n0;
    // this wave's bases / inverses stay on the device for the next wave (if they fit comfortably); they are owned
    // by DevBufs until the wave has succeeded, so that an early return can neither leak nor publish them
    DevBuf cur_bi(st), cur_basis(st);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const size_t need = (size_t)nodes * (size_t)m * (size_t)m * 8 + (size_t)nodes * (size_t)m * 8;
    const bool keep = need < free_b / 2;
    if (keep) {  // stream-ordered pool: no device-wide synchronisation per wave
        CK(cur_bi.alloc((size_t)nodes * m * m * 8));
        CK(cur_basis.alloc((size_t)nodes * m * 8));
        // a node that never writes its basis must read as "no basis" (-1) to the next wave
        CK(cudaMemsetAsync(cur_basis.p, 0xff, (size_t)nodes * m * 8, st));
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
        P.warm_basis = r.prev_basis;
        P.warm_bi = r.prev_bi;
    }
    CK(dst.alloc(sizeof(int32_t) * nodes));
    CK(dF.alloc(sizeof(double) * nodes));
    CK(dX.alloc(sizeof(double) * nodes * n0));
    CK(dS.alloc(sizeof(int32_t) * nodes * 8));
    P.status = dst.as<int>(); P.optF = dF.as<double>(); P.x = dX.as<double>(); P.stats = dS.as<int>();
    if (keep) P.basis = cur_basis.as<long long>();
    else if (basis) { CK(dB.alloc(sizeof(int64_t) * nodes * m)); P.basis = dB.as<long long>(); }
    P.bi_out = keep ? cur_bi.as<double>() : nullptr;
    t_timing = gm_timing{};
    rc = launch_wave(*d, P, st, se.e[1], se.e[2], &t_timing);
    if (rc == GM_OK) {
        CK(cudaMemcpyAsync(status, dst.p, sizeof(int32_t) * nodes, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        finish_trace();
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
            rc = launch_wave(*d, Q, st, se.e[3], se.e[4], &t_timing);
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
    cudaStreamSynchronize(st);
    // publish this wave's state for the children only when it is complete; otherwise the root goes cold
    std::lock_guard<std::mutex> lk(g.mu);
    auto it = g.roots.find(root);
    if (it != g.roots.end()) {
        Root& rr = it->second;
        if (rr.prev_bi) cudaFreeAsync(rr.prev_bi, st);
        if (rr.prev_basis) cudaFreeAsync(rr.prev_basis, st);
        const bool pub = rc == GM_OK && keep;
        rr.prev_bi = pub ? static_cast<double*>(cur_bi.release()) : nullptr;
        rr.prev_basis = pub ? static_cast<long long*>(cur_basis.release()) : nullptr;
        rr.prev_nodes = pub ? nodes : 0;
        rr.prev_m = (int)m;
    }
    return rc;
}

}  // extern "C"
