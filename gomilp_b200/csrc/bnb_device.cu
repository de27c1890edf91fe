// bnb_device.cu — the branch-and-bound wavefront with the decisions made ON THE DEVICE, sharded over the GPUs of a
// box when a communicator is set (gm_comm_init). C ABI: gm_milp_solve_device, gm_comm_*.
//
// Replaces, relative to /root/reference:
//   enumerationTree.startSearch / bufferManager / solveWorker   tree.go:66-205   -> one wave per BFS level per GPU
//   checkSolution                                               tree.go:207-263  -> wave_scan (exclusive prefix-min)
//   feasibleForIP / isAllInteger                                tree.go:276-297  -> node_check (exact x == trunc(x))
//   solution.branch / getChild + the heuristics                 subproblem.go:193-259, branching.go:17-94
//   the candidate / incumbent channels                          tree.go:55-57    -> one ncclAllGather per wave
//
// Why a scan replays the reference: with one worker the reference checks candidates in FIFO order and a node only
// ever compares its z with the incumbent at that moment (tree.go:225-252). The incumbent is replaced iff the node is
// integer feasible and strictly better, so "the incumbent when node k is checked" is the EXCLUSIVE PREFIX MINIMUM of
// z over the integer-feasible nodes before k (seeded with the incumbent of the previous waves), child ids are an
// exclusive prefix sum of the branching flags, and the first solver failure (a panic in the reference, tree.go:272)
// cuts the wave. All three are scans over 32-byte node records, so x never leaves the device: a wave costs one solve
// launch, node_check, [one all-gather of the records], wave_scan and a 64-byte summary read.
//
// Multi-GPU: rank r solves the contiguous FIFO block [r*chunk, (r+1)*chunk) of every wave on its GPU;
// the records are all-gathered over NVLink (NCCL), every rank runs the same scan over the full wave and so holds the
// full next wave's descriptors (L x 20 B per node) without a second exchange. Pruning is therefore global and
// identical to the 1-GPU / 1-worker order. The incumbent's x stays on the rank that solved it and is broadcast once.
#include <dlfcn.h>
#include <fcntl.h>
#include <nccl.h>
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>

#include "engine.h"

using namespace gm_engine;

namespace {

struct NodeRec {  // 32 bytes
    double z;
    double floor_x;  // floor(x[bvar])
    int status;      // gm_status of the relaxation
    int intfeas;     // feasibleForIP
    int bvar;        // branching variable the heuristic picks
    int pivots;
};

struct WaveSummary {  // written by wave_scan, read by the host once per wave
    int next_count;   // children created
    int panic;        // gm_milp_status of a panic, 0 = none
    int panic_lp;
    int processed;    // nodes checked (stops at a panic)
    int inc_idx;      // node of this wave that became the incumbent last, -1 = none
    int have_inc;
    int root_feasible;
    int pad;
    long long pivots;
    double inc_z;
};

__device__ __forceinline__ bool is_integral(double v, double itol) {
    if (itol <= 0.0) return v == trunc(v);  // tree.go:290-297
    return fabs(v - nearbyint(v)) <= itol;
}

// One CTA per node of this rank's block: feasibleForIP + the branching point (branching.go) + floor(x_j).
// mode COMPAT: the reference never propagates the heuristic, so it is always maxFun, which always returns the last
// integer-flagged index (`compat_var`). FIXED: the heuristic is evaluated on the fractional integers.
__global__ void __launch_bounds__(128) node_check(int nodes, int n0, const int* __restrict__ status,
                                                  const double* __restrict__ z, const double* __restrict__ x,
                                                  const int* __restrict__ stats, const unsigned char* __restrict__ integ,
                                                  const double* __restrict__ c0, int mode, int heuristic, int compat_var,
                                                  double itol, const int* __restrict__ bvar_desc, int L,
                                                  NodeRec* __restrict__ rec) {
    __shared__ double s_score[128];
    __shared__ int s_idx[128];
    __shared__ int s_frac;
    const int k = blockIdx.x, t = threadIdx.x;
    if (k >= nodes) return;
    const int st = status[k];
    const double* xk = x + (size_t)k * n0;
    if (t == 0) s_frac = 0;
    __syncthreads();
    double best = -1.0;
    int besti = INT_MAX;
    if (st == GM_OK) {
        const int last = L > 0 ? bvar_desc[(size_t)k * L + L - 1] : -1;
        const int start = last < 0 ? 0 : (last + 1) % n0;
        for (int i = t; i < n0; i += 128) {
            if (!integ[i]) continue;
            const double v = xk[i];
            if (is_integral(v, itol)) continue;
            s_frac = 1;
            double score;
            if (heuristic == GM_BRANCH_NAIVE) {
                const int dist = (i - start + n0) % n0;  // first fractional integer in cyclic order from `start`
                score = (double)(n0 - dist);
            } else if (heuristic == GM_BRANCH_MOST_INFEASIBLE) {
                const double f = v - floor(v);
                score = 0.5 - fabs(0.5 - f);
            } else {
                score = fabs(c0[i]);
            }
            if (score > best) { best = score; besti = i; }  // strict: the first maximum wins inside a thread
        }
    }
    s_score[t] = best;
    s_idx[t] = besti;
    __syncthreads();
    for (int d = 64; d >= 1; d >>= 1) {
        if (t < d) {
            const double os = s_score[t + d];
            const int oi = s_idx[t + d];
            if (oi != INT_MAX && (s_idx[t] == INT_MAX || os > s_score[t] || (os == s_score[t] && oi < s_idx[t]))) {
                s_score[t] = os;
                s_idx[t] = oi;
            }
        }
        __syncthreads();
    }
    if (t == 0) {
        NodeRec r;
        r.z = z[k];
        r.status = st;
        r.intfeas = (st == GM_OK && !s_frac) ? 1 : 0;
        int var = compat_var;
        if (mode != GM_BNB_COMPAT && s_idx[0] != INT_MAX) var = s_idx[0];
        r.bvar = var;
        r.floor_x = st == GM_OK ? floor(xk[var]) : 0.0;
        r.pivots = stats[(size_t)k * 8] + stats[(size_t)k * 8 + 1];
        rec[k] = r;
    }
}

// Warm-started waves: the nodes whose warm start died (GM_ERR_WARM_RETRY) are re-solved cold by a second launch over
// this list. out[0] = how many.
__global__ void __launch_bounds__(256) warm_retry_list(int nodes, const int* __restrict__ status, int* __restrict__ list,
                                                       int* __restrict__ out) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nodes; k += gridDim.x * blockDim.x)
        if (status[k] == GM_ERR_WARM_RETRY) list[atomicAdd(out, 1)] = k;
}

// checkSolution over the whole wave in FIFO order, one CTA. See the file header for why it is a scan.
constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads) wave_scan(int count, int wave_no, int L, const NodeRec* __restrict__ rec,
                                                          int have_inc0, double inc0_z, long long id_base,
                                                          const int* __restrict__ cur_bvar,
                                                          const double* __restrict__ cur_bsign,
                                                          const double* __restrict__ cur_brhs, int* __restrict__ decision,
                                                          int* __restrict__ next_bvar, double* __restrict__ next_bsign,
                                                          double* __restrict__ next_brhs, int* __restrict__ next_parent,
                                                          const long long* __restrict__ cur_id,
                                                          long long* __restrict__ next_id, WaveSummary* out) {
    __shared__ double s_min[kScanThreads];
    __shared__ int s_sum[kScanThreads];
    __shared__ int s_i[kScanThreads];
    __shared__ long long s_piv[kScanThreads];
    const int t = threadIdx.x;
    const double inf = INFINITY;
    if (wave_no == 0) {  // the root: tree.go:72-93
        if (t == 0) {
            WaveSummary s{};
            const NodeRec r = rec[0];
            s.processed = 1;
            s.pivots = r.pivots;
            s.inc_idx = -1;
            s.have_inc = 0;
            s.inc_z = inf;
            int dec = GM_DEC_NONE;
            if (r.status != GM_OK) {  // subproblem.go:173-176: any root failure panics
                s.panic = GM_MILP_PANIC_ROOT;
                s.panic_lp = r.status;
            } else if (r.intfeas) {
                dec = GM_DEC_INITIAL_RX_FEASIBLE_FOR_IP;
                s.root_feasible = 1;
                s.have_inc = 1;
                s.inc_idx = 0;
                s.inc_z = r.z;
            } else if (inf > r.z) {
                dec = GM_DEC_BETTER_THAN_INCUMBENT_BRANCHING;
                for (int ch = 0; ch < 2; ++ch) {
                    next_bvar[ch] = r.bvar;
                    next_bsign[ch] = ch == 0 ? 1.0 : -1.0;
                    next_brhs[ch] = ch == 0 ? r.floor_x : -(r.floor_x + 1.0);
                    next_parent[ch] = 0;
                    next_id[ch] = id_base + 1 + ch;
                }
                s.next_count = 2;
            } else if (inf <= r.z) {
                dec = GM_DEC_WORSE_THAN_INCUMBENT;
            } else {
                s.panic = GM_MILP_PANIC_UNEXPECTED_CASE;
                s.panic_lp = r.status;
            }
            decision[0] = dec;
            *out = s;
        }
        return;
    }
    // ---- first panic (independent of the incumbent): tree.go:266-273 / :253-256
    int kp = INT_MAX;
    for (int k = t; k < count; k += kScanThreads) {
        const int st = rec[k].status;
        const bool bad = (st != GM_OK && st != GM_ERR_INFEASIBLE && st != GM_ERR_SINGULAR) || (st == GM_OK && rec[k].z != rec[k].z);
        if (bad && k < kp) kp = k;
    }
    s_i[t] = kp;
    __syncthreads();
    for (int d = kScanThreads / 2; d >= 1; d >>= 1) {
        if (t < d) s_i[t] = min(s_i[t], s_i[t + d]);
        __syncthreads();
    }
    kp = s_i[0];
    __syncthreads();
    const int lim = kp == INT_MAX ? count : kp;  // nodes [0, lim) get a decision
    double carry_min = have_inc0 ? inc0_z : inf;
    int carry_sum = 0;
    int last_feas = -1;
    long long piv = 0;
    for (int base = 0; base < lim; base += kScanThreads) {
        const int k = base + t;
        NodeRec r{};
        bool live = k < lim;
        if (live) r = rec[k];
        const bool ok = live && r.status == GM_OK;
        // ---- exclusive prefix minimum of the integer-feasible objectives
        const double cand = (ok && r.intfeas) ? r.z : inf;
        s_min[t] = cand;
        __syncthreads();
        for (int d = 1; d < kScanThreads; d <<= 1) {
            const double o = t >= d ? s_min[t - d] : inf;
            __syncthreads();
            s_min[t] = fmin(s_min[t], o);
            __syncthreads();
        }
        const double inc_before = fmin(carry_min, t > 0 ? s_min[t - 1] : inf);
        const double chunk_min = s_min[kScanThreads - 1];
        // ---- decision (tree.go:225-252)
        int dec = GM_DEC_NONE, branch = 0, feas = 0;
        if (live) {
            if (!ok) dec = r.status == GM_ERR_INFEASIBLE ? GM_DEC_SUBPROBLEM_IS_DEGENERATE : GM_DEC_SUBPROBLEM_NOT_FEASIBLE;
            else if (inc_before <= r.z) dec = GM_DEC_WORSE_THAN_INCUMBENT;
            else if (r.intfeas) { dec = GM_DEC_BETTER_THAN_INCUMBENT_FEASIBLE; feas = 1; }
            else { dec = GM_DEC_BETTER_THAN_INCUMBENT_BRANCHING; branch = 1; }
            decision[k] = dec;
        }
        // ---- exclusive prefix sum of the branching flags -> child slots; last new incumbent; pivots
        s_sum[t] = branch;
        s_i[t] = feas ? k : -1;
        s_piv[t] = live ? r.pivots : 0;
        __syncthreads();
        for (int d = 1; d < kScanThreads; d <<= 1) {
            const int o = t >= d ? s_sum[t - d] : 0;
            const int oi = t >= d ? s_i[t - d] : -1;
            const long long op = t >= d ? s_piv[t - d] : 0;
            __syncthreads();
            s_sum[t] += o;
            s_i[t] = max(s_i[t], oi);
            s_piv[t] += op;
            __syncthreads();
        }
        if (branch) {
            const int slot = carry_sum + s_sum[t] - 1;  // exclusive rank among the branching nodes
            for (int ch = 0; ch < 2; ++ch) {
                const size_t dst = (size_t)(2 * slot + ch) * (L + 1);
                for (int l = 0; l < L; ++l) {
                    next_bvar[dst + l] = cur_bvar[(size_t)k * L + l];
                    next_bsign[dst + l] = cur_bsign[(size_t)k * L + l];
                    next_brhs[dst + l] = cur_brhs[(size_t)k * L + l];
                }
                next_bvar[dst + L] = r.bvar;                                   // getChild, subproblem.go:230-259
                next_bsign[dst + L] = ch == 0 ? 1.0 : -1.0;                    // x_j <= floor | -x_j <= -(floor + 1)
                next_brhs[dst + L] = ch == 0 ? r.floor_x : -(r.floor_x + 1.0);
                next_parent[2 * slot + ch] = k;
                next_id[2 * slot + ch] = id_base + 2 * (long long)slot + 1 + ch;
            }
        }
        carry_sum += s_sum[kScanThreads - 1];
        carry_min = fmin(carry_min, chunk_min);
        last_feas = max(last_feas, s_i[kScanThreads - 1]);
        piv += s_piv[kScanThreads - 1];
        __syncthreads();
    }
    if (t == 0) {
        WaveSummary s{};
        s.next_count = 2 * carry_sum;
        s.processed = lim;
        s.pivots = piv;
        s.inc_idx = last_feas;
        s.have_inc = (have_inc0 || last_feas >= 0) ? 1 : 0;
        s.inc_z = last_feas >= 0 ? rec[last_feas].z : (have_inc0 ? inc0_z : inf);
        if (kp != INT_MAX) {
            const NodeRec r = rec[kp];
            s.processed = kp + 1;
            s.pivots += r.pivots;
            s.panic = r.status == GM_OK ? GM_MILP_PANIC_UNEXPECTED_CASE : GM_MILP_PANIC_SOLVER_FAILURE;
            s.panic_lp = r.status;
            decision[kp] = GM_DEC_NONE;
        }
        *out = s;
    }
}

// ---- NCCL, loaded at run time: libgomilp_b200.so has no link-time dependency on it, and in a process that already
// loaded a libnccl (PyTorch bundles one) the same library instance is reused. --------------------------------------
struct Nccl {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load() {
        if (h) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) return false;
        GetUniqueId = (decltype(GetUniqueId))dlsym(h, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(h, "ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
        AllGather = (decltype(AllGather))dlsym(h, "ncclAllGather");
        Broadcast = (decltype(Broadcast))dlsym(h, "ncclBroadcast");
        GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
        return GetUniqueId && CommInitRank && CommDestroy && AllGather && Broadcast;
    }
};
Nccl nccl;
std::mutex nccl_mu;

// NCCL announces its version on stdout when a process makes its first communicator; callers of this library own
// stdout (bench.py prints exactly one JSON line), so the announcement goes to /dev/null.
struct StdoutSilencer {
    int saved = -1;
    StdoutSilencer() {
        fflush(stdout);
        saved = dup(1);
        const int devnull = open("/dev/null", O_WRONLY);
        if (saved >= 0 && devnull >= 0) dup2(devnull, 1);
        if (devnull >= 0) close(devnull);
    }
    ~StdoutSilencer() {
        fflush(stdout);
        if (saved >= 0) { dup2(saved, 1); close(saved); }
    }
};

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = -1;
};
thread_local Comm t_comm;

int nccl_fail(ncclResult_t r, const char* what) {
    t_err = std::string(what) + ": " + (nccl.GetErrorString ? nccl.GetErrorString(r) : "nccl error");
    return GM_ERR_CUDA;
}
#define NK(call)                                            \
    do {                                                    \
        ncclResult_t r_ = (call);                           \
        if (r_ != ncclSuccess) return nccl_fail(r_, #call); \
    } while (0)

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    cudaStream_t s;
    explicit Buf(cudaStream_t s_) : s(s_) {}
    Buf(Buf&& o) noexcept : p(o.p), cap(o.cap), s(o.s) { o.p = nullptr; o.cap = 0; }  // owning: moves only
    Buf(const Buf&) = delete;
    Buf& operator=(const Buf&) = delete;
    ~Buf() { if (p) cudaFreeAsync(p, s); }
    cudaError_t reserve(size_t bytes) {  // contents are NOT preserved
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeAsync(p, s);
        p = nullptr;
        cap = bytes + bytes / 2 + 256;
        return cudaMallocAsync(&p, cap, s);
    }
    template <class T> T* as() { return static_cast<T*>(p); }
};

// Every device buffer of one search. Lives with the leased stream (StreamEvents::arena) and is reused by the next
// search on that stream: the waves double in width, so a fresh set of buffers means ~15 growing allocations per
// buffer and, with warm start, GBs of stream-ordered allocation per call.
struct BnbArena {
    Buf d_integ, d_rec, d_dec, d_sum, d_incx;
    Buf desc_v[2], desc_s[2], desc_r[2], par[2], ids[2];
    Buf d_status, d_z, d_x, d_stats;
    Buf wbi[2], wbasis[2], d_retry, d_nretry;
    WaveSummary* h_sum = nullptr;  // pinned
    int* h_nretry = nullptr;       // pinned
    explicit BnbArena(cudaStream_t st)
        : d_integ(st), d_rec(st), d_dec(st), d_sum(st), d_incx(st), desc_v{Buf(st), Buf(st)}, desc_s{Buf(st), Buf(st)},
          desc_r{Buf(st), Buf(st)}, par{Buf(st), Buf(st)}, ids{Buf(st), Buf(st)}, d_status(st), d_z(st), d_x(st),
          d_stats(st), wbi{Buf(st), Buf(st)}, wbasis{Buf(st), Buf(st)}, d_retry(st), d_nretry(st) {}
    ~BnbArena() {
        if (h_sum) cudaFreeHost(h_sum);
        if (h_nretry) cudaFreeHost(h_nretry);
    }
    BnbArena(const BnbArena&) = delete;
    BnbArena& operator=(const BnbArena&) = delete;
};

}  // namespace

namespace gm_engine {
cudaError_t gm_bnb_kernels_prepare() { return cudaSuccess; }
}  // namespace gm_engine

extern "C" {

int gm_comm_unique_id(void* id128) {
    if (!id128) return GM_ERR_BAD_ARGUMENT;
    std::lock_guard<std::mutex> lk(nccl_mu);
    if (!nccl.load()) { t_err = "libnccl.so.2 not found"; return GM_ERR_CUDA; }
    ncclUniqueId id;
    {
        StdoutSilencer quiet;
        NK(nccl.GetUniqueId(&id));
    }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(id128, &id, 128);
    return GM_OK;
}

int gm_comm_init(int32_t rank, int32_t world, const void* id128) {
    if (!id128 || world < 1 || rank < 0 || rank >= world) return GM_ERR_BAD_ARGUMENT;
    DeviceCtx* d = nullptr;
    int rc = current_device(&d);
    if (rc != GM_OK) return rc;
    {
        std::lock_guard<std::mutex> lk(nccl_mu);
        if (!nccl.load()) { t_err = "libnccl.so.2 not found"; return GM_ERR_CUDA; }
    }
    if (t_comm.comm) { nccl.CommDestroy(t_comm.comm); t_comm = Comm{}; }
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    ncclComm_t c = nullptr;
    {
        StdoutSilencer quiet;
        NK(nccl.CommInitRank(&c, world, id, rank));
    }
    t_comm.comm = c;
    t_comm.rank = rank;
    t_comm.world = world;
    t_comm.device = d->id;
    return GM_OK;
}

int gm_comm_destroy(void) {
    if (t_comm.comm) nccl.CommDestroy(t_comm.comm);
    t_comm = Comm{};
    return GM_OK;
}

int gm_milp_solve_device(int64_t nvar, const double* c, int64_t meq, const double* A, const double* b, int64_t nineq,
                         const double* G, const double* h, const uint8_t* integrality, int32_t heuristic, int32_t mode,
                         int64_t node_limit, double time_limit_s, double* x_out, gm_milp_result* result,
                         gm_decision_cb on_decision, gm_wave_cb on_wave, void* user) {
    if (!result || !x_out || !c || !integrality || nvar <= 0 || meq < 0 || nineq < 0) return GM_ERR_BAD_ARGUMENT;
    if ((meq > 0 && (!A || !b)) || (nineq > 0 && (!G || !h)) || meq + nineq == 0) return GM_ERR_BAD_ARGUMENT;
    const auto t0 = std::chrono::steady_clock::now();
    std::memset(result, 0, sizeof(*result));
    DeviceCtx* d = nullptr;
    int rc = current_device(&d);
    if (rc != GM_OK) { result->status = GM_MILP_ENGINE_ERROR; result->lp_status = rc; return rc; }
    const int world = t_comm.comm ? t_comm.world : 1, rank = t_comm.comm ? t_comm.rank : 0;
    if (t_comm.comm && t_comm.device != d->id) return GM_ERR_BAD_ARGUMENT;
    const bool warm = (mode & GM_BNB_WARM_START) != 0;
    const double itol = warm ? 1e-9 : 0.0;
    const int bmode = mode & 3;
    RobustScope robust_scope((mode & GM_BNB_ROBUST) != 0);

    // toInitialSubproblem (ilp.go:43-71) + convertToEqualities (subproblem.go:81-139): [A 0; G I]
    const int64_t m0 = meq + nineq, n0 = nvar + nineq;
    std::vector<double> c0(n0, 0.0), A0((size_t)m0 * n0, 0.0), b0(m0, 0.0);
    std::vector<uint8_t> integ(n0, 0);
    for (int64_t j = 0; j < nvar; ++j) { c0[j] = c[j]; integ[j] = integrality[j] ? 1 : 0; }
    for (int64_t i = 0; i < meq; ++i) {
        for (int64_t j = 0; j < nvar; ++j) A0[(size_t)i * n0 + j] = A[(size_t)i * nvar + j];
        b0[i] = b[i];
    }
    for (int64_t i = 0; i < nineq; ++i) {
        for (int64_t j = 0; j < nvar; ++j) A0[(size_t)(meq + i) * n0 + j] = G[(size_t)i * nvar + j];
        A0[(size_t)(meq + i) * n0 + nvar + i] = 1.0;
        b0[meq + i] = h[i];
    }
    // branching.go:54-72 (maxFun): the running maximum is never stored back, so the last integer-flagged index wins
    int compat_var = 0;
    for (int64_t i = 0; i < n0; ++i)
        if (integ[i] && std::fabs(c0[i]) >= 0.0) compat_var = (int)i;

    gm_root_t root = 0;
    rc = gm_upload_root(c0.data(), A0.data(), n0, b0.data(), m0, n0, &root);
    if (rc != GM_OK) { result->status = GM_MILP_ENGINE_ERROR; result->lp_status = rc; return rc; }
    Root r;
    DeviceCtx* dr = nullptr;
    rc = device_of_root(root, &r, &dr);
    if (rc != GM_OK) return rc;

    StreamLease lease(d);
    CK(lease.err);
    StreamEvents& se = *lease.se;
    cudaStream_t st = se.s;
    int engine_rc = GM_OK;
    {
        if (se.arena == nullptr) {
            se.arena = new BnbArena(st);
            se.arena_free = [](void* a) { delete static_cast<BnbArena*>(a); };
        }
        BnbArena& ar = *static_cast<BnbArena*>(se.arena);
        Buf &d_integ = ar.d_integ, &d_rec = ar.d_rec, &d_dec = ar.d_dec, &d_sum = ar.d_sum, &d_incx = ar.d_incx;
        Buf (&desc_v)[2] = ar.desc_v, (&desc_s)[2] = ar.desc_s, (&desc_r)[2] = ar.desc_r;
        Buf (&par)[2] = ar.par, (&ids)[2] = ar.ids;
        Buf &d_status = ar.d_status, &d_z = ar.d_z, &d_x = ar.d_x, &d_stats = ar.d_stats;
        // warm start (GM_BNB_WARM_START, one GPU): every node's final basis and inverse stay in HBM for its children
        Buf (&wbi)[2] = ar.wbi, (&wbasis)[2] = ar.wbasis;
        Buf &d_retry = ar.d_retry, &d_nretry = ar.d_nretry;
        int*& h_nretry = ar.h_nretry;
        bool prev_kept = false;
        WaveSummary*& h_sum = ar.h_sum;
        std::vector<NodeRec> h_rec;
        std::vector<int> h_dec, h_par;
        std::vector<long long> h_ids, h_pids;
#define CKE(call)                                                              \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) { engine_rc = fail(e_, #call); goto done; }     \
    } while (0)
#define NKE(call)                                                              \
    do {                                                                       \
        ncclResult_t r_ = (call);                                              \
        if (r_ != ncclSuccess) { engine_rc = nccl_fail(r_, #call); goto done; } \
    } while (0)
        bool have_inc = false, timed_out = false;
        double inc_z = std::numeric_limits<double>::infinity();
        int inc_owner = -1;  // rank whose d_incx holds the incumbent's x
        long long next_id = 0;
        int panic = 0, panic_lp = 0;
        int64_t count = 1, wave_no = 0;
        int L = 0, cur = 0;
        std::vector<long long> prev_ids;  // ids of the previous wave (parents), kept only for the decision callback

        if (h_sum == nullptr) CKE(cudaMallocHost(&h_sum, sizeof(WaveSummary)));
        if (h_nretry == nullptr) CKE(cudaMallocHost(&h_nretry, sizeof(int)));
        CKE(d_nretry.reserve(sizeof(int)));
        CKE(d_integ.reserve(n0));
        CKE(cudaMemcpyAsync(d_integ.p, integ.data(), n0, cudaMemcpyHostToDevice, st));
        CKE(d_sum.reserve(sizeof(WaveSummary)));
        CKE(d_incx.reserve(sizeof(double) * n0));
        CKE(ids[0].reserve(sizeof(long long)));
        CKE(cudaMemsetAsync(ids[0].p, 0, sizeof(long long), st));
        CKE(desc_v[0].reserve(8)); CKE(desc_s[0].reserve(8)); CKE(desc_r[0].reserve(8)); CKE(par[0].reserve(8));
        prev_ids.assign(1, 0);

        while (count > 0 && !panic) {
            if (node_limit > 0) {  // the context deadline of ilp.go:92-99, expressed as a node budget
                const int64_t left = node_limit - result->nodes;
                if (left <= 0) { timed_out = true; break; }
                if (count > left) { count = left; timed_out = true; }
            }
            if (time_limit_s > 0 && wave_no > 0 && world == 1) {
                const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                if (el > time_limit_s) { timed_out = true; break; }
            }
            // ---- this rank's contiguous FIFO block
            const int64_t chunk = (count + world - 1) / world;
            const int64_t lo = std::min<int64_t>(count, rank * chunk), hi = std::min<int64_t>(count, lo + chunk);
            const int64_t mine = hi - lo;
            CKE(d_rec.reserve(sizeof(NodeRec) * chunk * world));
            CKE(d_dec.reserve(sizeof(int) * count));
            const int nxt = cur ^ 1;
            CKE(desc_v[nxt].reserve(sizeof(int) * 2 * count * (L + 1)));
            CKE(desc_s[nxt].reserve(sizeof(double) * 2 * count * (L + 1)));
            CKE(desc_r[nxt].reserve(sizeof(double) * 2 * count * (L + 1)));
            CKE(par[nxt].reserve(sizeof(int) * 2 * count));
            CKE(ids[nxt].reserve(sizeof(long long) * 2 * count));
            CKE(cudaEventRecord(se.e[0], st));
            if (mine > 0) {
                CKE(d_status.reserve(sizeof(int) * mine));
                CKE(d_z.reserve(sizeof(double) * mine));
                CKE(d_x.reserve(sizeof(double) * mine * n0));
                CKE(d_stats.reserve(sizeof(int) * 8 * mine));
                gm::BatchParams P;
                std::memset(&P, 0, sizeof(P));
                P.c = r.c; P.A = r.A; P.b = r.b;
                P.m0 = (int)m0; P.n0 = (int)n0; P.lda = (int)n0; P.L = L; P.tol = 0.0;  // subproblem.go:154,172: tol = 0
                P.count = (int)mine; P.x_len = (int)n0; P.x_stride = n0;
                if (L > 0) {
                    P.bvar = desc_v[cur].as<int>() + lo * L;
                    P.bsign = desc_s[cur].as<double>() + lo * L;
                    P.brhs = desc_r[cur].as<double>() + lo * L;
                }
                P.status = d_status.as<int>(); P.optF = d_z.as<double>(); P.x = d_x.as<double>();
                P.stats = d_stats.as<int>();
                const int m = (int)m0 + L;
                bool keep = false, warmed = false;
                if (warm && world == 1) {  // children continue from [B 0; g 1]^-1 read off the parent's inverse
                    size_t free_b = 0, total_b = 0;
                    cudaMemGetInfo(&free_b, &total_b);
                    keep = ((size_t)mine * m * m * 8 + (size_t)mine * m * 8) * 2 < free_b / 2;
                    if (keep) {
                        CKE(wbi[cur].reserve((size_t)mine * m * m * 8));
                        CKE(wbasis[cur].reserve((size_t)mine * m * 8));
                        CKE(cudaMemsetAsync(wbasis[cur].p, 0xff, (size_t)mine * m * 8, st));  // "no basis" unless written
                        P.bi_out = wbi[cur].as<double>();
                        P.basis = wbasis[cur].as<long long>();
                    }
                    if (wave_no > 0 && prev_kept) {
                        P.warm_parent = par[cur].as<int>();   // written by the previous wave's scan
                        P.warm_basis = wbasis[cur ^ 1].as<long long>();
                        P.warm_bi = wbi[cur ^ 1].as<double>();
                        warmed = true;
                    }
                }
                gm_timing tm{};
                engine_rc = launch_wave(*d, P, st, nullptr, nullptr, &tm);
                if (engine_rc != GM_OK) goto done;
                if (warmed) {  // a warm path that died is re-solved from scratch, in place
                    CKE(d_retry.reserve(sizeof(int) * mine));
                    CKE(cudaMemsetAsync(d_nretry.p, 0, sizeof(int), st));
                    warm_retry_list<<<(unsigned)std::min<int64_t>(64, (mine + 255) / 256), 256, 0, st>>>(
                        (int)mine, d_status.as<int>(), d_retry.as<int>(), d_nretry.as<int>());
                    CKE(cudaGetLastError());
                    CKE(cudaMemcpyAsync(h_nretry, d_nretry.p, sizeof(int), cudaMemcpyDeviceToHost, st));
                    CKE(cudaStreamSynchronize(st));
                    if (*h_nretry > 0) {
                        gm::BatchParams Q = P;
                        Q.warm_parent = nullptr;
                        Q.lp_list = d_retry.as<int>();
                        Q.count = *h_nretry;
                        engine_rc = launch_wave(*d, Q, st, nullptr, nullptr, &tm);
                        if (engine_rc != GM_OK) goto done;
                    }
                }
                prev_kept = keep;
                node_check<<<(unsigned)mine, 128, 0, st>>>((int)mine, (int)n0, d_status.as<int>(), d_z.as<double>(),
                                                            d_x.as<double>(), d_stats.as<int>(),
                                                            d_integ.as<unsigned char>(), r.c, bmode, heuristic,
                                                            compat_var, itol, L > 0 ? desc_v[cur].as<int>() + lo * L : nullptr,
                                                            L, d_rec.as<NodeRec>() + lo);
                CKE(cudaGetLastError());
            }
            if (world > 1) {  // the wave's exchange step: 32 bytes per node over NVLink
                NKE(nccl.AllGather(d_rec.as<NodeRec>() + rank * chunk, d_rec.p, sizeof(NodeRec) * chunk, ncclChar,
                                   t_comm.comm, st));
            }
            wave_scan<<<1, kScanThreads, 0, st>>>((int)count, (int)wave_no, L, d_rec.as<NodeRec>(), have_inc ? 1 : 0,
                                                  inc_z, next_id, desc_v[cur].as<int>(), desc_s[cur].as<double>(),
                                                  desc_r[cur].as<double>(), d_dec.as<int>(), desc_v[nxt].as<int>(),
                                                  desc_s[nxt].as<double>(), desc_r[nxt].as<double>(), par[nxt].as<int>(),
                                                  ids[cur].as<long long>(), ids[nxt].as<long long>(),
                                                  d_sum.as<WaveSummary>());
            CKE(cudaGetLastError());
            CKE(cudaEventRecord(se.e[1], st));
            CKE(cudaMemcpyAsync(h_sum, d_sum.p, sizeof(WaveSummary), cudaMemcpyDeviceToHost, st));
            CKE(cudaStreamSynchronize(st));
            const WaveSummary s = *h_sum;
            const double wave_ms = ms(se.e[0], se.e[1]);
            result->waves += 1;
            result->device_ms += wave_ms;
            result->nodes += s.processed;
            result->pivots += s.pivots;
            if (on_wave) on_wave(user, wave_no, s.processed, s.pivots, wave_ms);
            if (on_decision) {  // BnbMiddleware.ProcessDecision, in FIFO order (instrumentation.go:8-15)
                h_rec.resize(count);
                h_dec.resize(count);
                h_ids.resize(count);
                h_par.resize(count);
                CKE(cudaMemcpyAsync(h_rec.data(), d_rec.p, sizeof(NodeRec) * count, cudaMemcpyDeviceToHost, st));
                CKE(cudaMemcpyAsync(h_dec.data(), d_dec.p, sizeof(int) * count, cudaMemcpyDeviceToHost, st));
                CKE(cudaMemcpyAsync(h_ids.data(), ids[cur].p, sizeof(long long) * count, cudaMemcpyDeviceToHost, st));
                if (wave_no > 0)
                    CKE(cudaMemcpyAsync(h_par.data(), par[cur].p, sizeof(int) * count, cudaMemcpyDeviceToHost, st));
                CKE(cudaStreamSynchronize(st));
                for (int k = 0; k < s.processed; ++k) {
                    if (s.panic && k == s.processed - 1) break;  // the reference panics before reporting it
                    const NodeRec& nr = h_rec[k];
                    const bool br = h_dec[k] == GM_DEC_BETTER_THAN_INCUMBENT_BRANCHING;
                    const long long pid = wave_no > 0 ? prev_ids[h_par[k]] : 0;
                    on_decision(user, h_ids[k], pid, L, nr.status, nr.z, h_dec[k], br ? nr.bvar : -1,
                                br ? nr.floor_x : 0.0);
                }
                prev_ids = h_ids;
            }
            if (s.inc_idx >= 0) {  // the incumbent's x stays on the rank that solved it
                have_inc = true;
                inc_z = s.inc_z;
                inc_owner = (int)(s.inc_idx / chunk);
                if (inc_owner == rank)
                    CKE(cudaMemcpyAsync(d_incx.p, d_x.as<double>() + (size_t)(s.inc_idx - lo) * n0, sizeof(double) * n0,
                                        cudaMemcpyDeviceToDevice, st));
            }
            if (s.panic) { panic = s.panic; panic_lp = s.panic_lp; break; }
            if (s.root_feasible) break;  // startSearch returns the root directly (tree.go:88-92)
            if (timed_out) break;
            next_id += s.next_count;
            count = s.next_count;
            cur = nxt;
            L += 1;
            ++wave_no;
        }
        // ---- the result (ilp.go:92-116)
        if (panic) {
            result->status = panic;
            result->lp_status = panic_lp;
        } else if (!have_inc) {
            result->status = timed_out ? GM_MILP_DEADLINE_EXCEEDED : GM_MILP_NO_INTEGER_FEASIBLE_SOLUTION;
        } else {
            if (world > 1) NKE(nccl.Broadcast(d_incx.p, d_incx.p, sizeof(double) * n0, ncclChar, inc_owner, t_comm.comm, st));
            std::vector<double> hx(n0);
            CKE(cudaMemcpyAsync(hx.data(), d_incx.p, sizeof(double) * n0, cudaMemcpyDeviceToHost, st));
            CKE(cudaStreamSynchronize(st));
            // a deadline returns *incumbent as is, slack entries included (ilp.go:92-99); else they are dropped (:111)
            const int64_t xl = timed_out ? n0 : nvar;
            std::memcpy(x_out, hx.data(), sizeof(double) * xl);
            result->x_len = xl;
            result->z = inc_z;
            result->status = timed_out ? GM_MILP_DEADLINE_EXCEEDED : GM_MILP_OK;
        }
    done:
        cudaStreamSynchronize(st);  // the arena (device buffers, pinned summaries) stays with the stream
#undef CKE
#undef NKE
    }
    gm_free_root(root);
    if (engine_rc != GM_OK) {
        result->status = GM_MILP_ENGINE_ERROR;
        result->lp_status = engine_rc;
        return engine_rc;
    }
    return GM_OK;
}

}  // extern "C"
