// oracle/oracle_capi.cpp — TEST INFRASTRUCTURE ONLY. C entry points of the CPU oracle so that
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs can call it
// through ctypes. Nothing under gomilp_b200/ may load this library (see oracle/README.md).
#include <atomic>
#include <thread>

#include "gomilp_bnb.hpp"

using namespace orc;

extern "C" {

// stats[8] = {pivots_phase1, pivots_phase2, bland_calls, lu_factorizations, repair_trials,
//             used_phase1, trace_len, x_is_non_nil}
// trace (optional) = int32[trace_cap][4] rows of (phase, enter, leave, bland)
int orc_simplex(const double* c, const double* A, int64_t lda, const double* b, int64_t m, int64_t n, double tol,
                const int64_t* initialBasic, double* optF, double* optX, int64_t* basisOut, int64_t* stats,
                int32_t* trace, int64_t trace_cap, int64_t max_pivots) {
    if (m <= 0 || n <= 0 || lda < n) return GM_ERR_BAD_SHAPE;
    SimplexStats st;
    st.max_pivots = max_pivots;
    std::vector<PivotRecord> tr;
    if (trace && trace_cap > 0) st.trace = &tr;
    ivec ib;
    if (initialBasic) {
        ib.resize(m);
        for (int64_t i = 0; i < m; ++i) ib[i] = (int)initialBasic[i];
    }
    SimplexResult r = simplex(c, A, (int)lda, b, (int)m, (int)n, tol, initialBasic ? ib.data() : nullptr, st);
    if (optF) *optF = r.optF;
    if (optX) {
        for (int64_t i = 0; i < n; ++i) optX[i] = r.x.empty() ? 0.0 : r.x[i];
    }
    if (basisOut) {
        for (int64_t i = 0; i < m; ++i) basisOut[i] = r.basis.empty() ? -1 : r.basis[i];
    }
    if (stats) {
        stats[0] = st.pivots_phase1;
        stats[1] = st.pivots_phase2;
        stats[2] = st.bland_calls;
        stats[3] = st.lu_factorizations;
        stats[4] = st.repair_trials;
        stats[5] = st.used_phase1;
        stats[6] = (int64_t)tr.size();
        stats[7] = r.x.empty() ? 0 : 1;
    }
    if (trace) {
        int64_t k = std::min<int64_t>(trace_cap, (int64_t)tr.size());
        for (int64_t i = 0; i < k; ++i) {
            trace[4 * i + 0] = tr[i].phase;
            trace[4 * i + 1] = tr[i].enter;
            trace[4 * i + 2] = tr[i].leave;
            trace[4 * i + 3] = tr[i].bland;
        }
    }
    return r.status;
}

// Batch of independent same-shape LPs, `threads` worker threads (one LP per thread at a time), the
// CPU-baseline form of SURVEY §8(d): A is [batch][m][n] row-major, c [batch][n], b [batch][m].
int orc_simplex_batch(int64_t batch, const double* c, const double* A, const double* b, int64_t m, int64_t n,
                      double tol, int threads, int32_t* status, double* optF, double* optX, int64_t* basis,
                      int64_t* pivots, int64_t max_pivots) {
    if (batch < 0 || m <= 0 || n <= 0) return GM_ERR_BAD_SHAPE;
    if (threads < 1) threads = 1;
    std::atomic<int64_t> next{0};
    auto work = [&]() {
        for (;;) {
            int64_t i = next.fetch_add(1);
            if (i >= batch) break;
            SimplexStats st;
            st.max_pivots = max_pivots;
            SimplexResult r = simplex(c + i * n, A + i * m * n, (int)n, b + i * m, (int)m, (int)n, tol, nullptr, st);
            if (status) status[i] = r.status;
            if (optF) optF[i] = r.optF;
            if (optX)
                for (int64_t k = 0; k < n; ++k) optX[i * n + k] = r.x.empty() ? 0.0 : r.x[k];
            if (basis)
                for (int64_t k = 0; k < m; ++k) basis[i * m + k] = r.basis.empty() ? -1 : r.basis[k];
            if (pivots) pivots[i] = st.pivots_phase1 + st.pivots_phase2;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    return GM_OK;
}

// out_scalars[8] = {milp_status, lp_status, x_len, nodes, pivots, log_len, 0, 0}; z_out separate.
int orc_bnb_solve(int64_t nvar, const double* c, int64_t meq, const double* A, const double* b, int64_t nineq,
                  const double* G, const double* h, const uint8_t* integrality, int heuristic, int mode,
                  int64_t node_limit, double time_limit_s, int64_t max_pivots_per_lp, double* x_out, double* z_out,
                  int64_t* out_scalars, int64_t log_cap, int64_t* log_id, int64_t* log_parent, int32_t* log_depth,
                  int32_t* log_lp_status, double* log_z, int32_t* log_decision, int64_t* log_pivots,
                  int32_t* log_bvar, double* log_bfloor) {
    MilpProblem p;
    p.nvar = (int)nvar;
    p.c.assign(c, c + nvar);
    p.meq = (int)meq;
    if (meq > 0) { p.A.assign(A, A + meq * nvar); p.b.assign(b, b + meq); }
    p.nineq = (int)nineq;
    if (nineq > 0) { p.G.assign(G, G + nineq * nvar); p.h.assign(h, h + nineq); }
    p.integrality.assign(integrality, integrality + nvar);
    p.heuristic = heuristic;
    BnbOptions o;
    o.mode = mode;
    o.node_limit = node_limit;
    o.time_limit_s = time_limit_s;
    o.max_pivots_per_lp = max_pivots_per_lp;
    o.keep_log = log_cap > 0;
    BnbResult r = bnb_solve(p, o);
    for (size_t i = 0; i < r.x.size(); ++i) x_out[i] = r.x[i];
    *z_out = r.z;
    out_scalars[0] = r.status;
    out_scalars[1] = r.lp_status;
    out_scalars[2] = (int64_t)r.x.size();
    out_scalars[3] = r.nodes;
    out_scalars[4] = r.pivots;
    int64_t k = std::min<int64_t>(log_cap, (int64_t)r.log.size());
    out_scalars[5] = (int64_t)r.log.size();
    for (int64_t i = 0; i < k; ++i) {
        const NodeRecord& nr = r.log[i];
        log_id[i] = nr.id;
        log_parent[i] = nr.parent;
        log_depth[i] = nr.depth;
        log_lp_status[i] = nr.lp_status;
        log_z[i] = nr.z;
        log_decision[i] = nr.decision;
        log_pivots[i] = nr.pivots;
        log_bvar[i] = nr.branch_var;
        log_bfloor[i] = nr.branch_floor;
    }
    return r.status;
}

// convertToEqualities layout pin (subproblem_test.go:296-357): aNew is (meq+nineq)×(nvar+nineq)
void orc_convert_to_equalities(int64_t nvar, const double* c, int64_t meq, const double* A, const double* b,
                               int64_t nineq, const double* G, const double* h, double* cNew, double* aNew,
                               double* bNew) {
    vec vc(c, c + nvar), vA, vb, vG(G, G + nineq * nvar), vh(h, h + nineq), oc, oa, ob;
    if (meq > 0) { vA.assign(A, A + meq * nvar); vb.assign(b, b + meq); }
    convert_to_equalities(vc, vA, vb, (int)meq, vG, vh, (int)nineq, oc, oa, ob);
    std::copy(oc.begin(), oc.end(), cNew);
    std::copy(oa.begin(), oa.end(), aNew);
    std::copy(ob.begin(), ob.end(), bNew);
}

int orc_maxfun_branch_point(int64_t n, const double* c, const uint8_t* integ) {
    return maxfun_branch_point(vec(c, c + n), std::vector<char>(integ, integ + n));
}
int orc_most_infeasible_branch_point(int64_t n, const double* c, const uint8_t* integ) {
    return most_infeasible_branch_point(vec(c, c + n), std::vector<char>(integ, integ + n));
}
int orc_feasible_for_ip(int64_t n, const uint8_t* integ, const double* x) {
    return feasible_for_ip(std::vector<char>(integ, integ + n), vec(x, x + n)) ? 1 : 0;
}
// mat.Cond(a, 1) for an m×k (m >= k) row-major matrix — exposed for estimator tests
double orc_cond1(const double* a, int64_t lda, int64_t m, int64_t k) { return cond1(a, (int)lda, (int)m, (int)k); }
// VecDense.SolveVec on a square system; returns 0 ok, 1 det==0, 2 cond>1e16
int orc_solve_vec(const double* a, int64_t n, int transpose, const double* b, double* x, double* cond) {
    return (int)solve_vec(a, (int)n, (int)n, transpose != 0, b, x, cond);
}
int orc_num_hw_threads() { return (int)std::thread::hardware_concurrency(); }
// findLinearlyIndependent (simplex.go:611-637) alone: the column indices it accepts, in acceptance order.
// Returns how many were accepted (< m: lp.ErrSingular at simplex.go:495-498).
int orc_find_linearly_independent(const double* A, int64_t lda, int64_t m, int64_t n, int64_t* idxs) {
    SimplexStats st;
    ivec r = detail::find_linearly_independent(A, (int)lda, (int)m, (int)n, st);
    for (size_t i = 0; i < r.size(); ++i) idxs[i] = r[i];
    return (int)r.size();
}

}  // extern "C"
