// oracle/lp_simplex.hpp — TEST INFRASTRUCTURE ONLY (CPU oracle; see oracle/README.md).
//
// Restatement of Gonum's dense two-phase simplex exactly as GoMILP calls it:
//   /root/reference/vendor/gonum.org/v1/gonum/optimize/convex/lp/simplex.go
//     Simplex :88-91, simplex :93-302, computeMove :306-342, replaceBland :347-383,
//     verifyInputs :385-439, initializeFromBasic :447-471, extractColumns :474-488,
//     findInitialBasic :492-607, findLinearlyIndependent :611-637, tolerances :42-58.
// Control flow, index bookkeeping (positional basic / non-basic lists mutated by swaps), first-index
// tie-breaks (floats.MinIdx), the three fresh LU solves per pivot and every tolerance follow the
// reference line by line in MEANING; the code structure is this repo's own.
#pragma once
#include <functional>

#include "../include/gomilp_status.h"
#include "gonum_linalg.hpp"

namespace orc {

struct PivotRecord {
    int phase;     // 1 = Phase I recursion, 2 = main problem
    int enter;     // variable (column) index entering the basis
    int leave;     // variable index leaving the basis
    int bland;     // 1 if chosen by replaceBland
};

struct SimplexStats {
    long pivots_phase1 = 0;
    long pivots_phase2 = 0;
    long bland_calls = 0;
    long lu_factorizations = 0;  // SolveVec + mat.Cond factorisations
    long repair_trials = 0;      // artificial-still-basic repair loop iterations, simplex.go:589-605
    int used_phase1 = 0;
    std::vector<PivotRecord>* trace = nullptr;  // optional
    long max_pivots = 0;                        // 0 = unlimited (the reference has no cap)
};

struct SimplexResult {
    int status = GM_OK;
    double optF = std::numeric_limits<double>::quiet_NaN();
    vec x;          // empty when the reference returns nil
    ivec basis;     // empty when the reference returns nil
};

namespace detail {

// m×k column-gather of row-major A (extractColumns, simplex.go:474-488)
inline void extract_columns(double* dst, int ldd, const double* A, int lda, int m, const int* cols, int k) {
    for (int j = 0; j < k; ++j)
        for (int i = 0; i < m; ++i) dst[(size_t)i * ldd + j] = A[(size_t)i * lda + cols[j]];
}

// initializeFromBasic, simplex.go:447-471. 0 = ok, 1 = singular (solve error), 2 = infeasible.
inline int initialize_from_basic(double* xb, const double* ab, int m, const double* b, SimplexStats& st) {
    st.lu_factorizations++;
    SolveErr e = solve_vec(ab, m, m, false, b, xb);
    if (e != SolveErr::None) return 1;
    for (int i = 0; i < m; ++i)
        if (xb[i] < -GM_INIT_POS_TOL) return 2;
    return 0;
}

// verifyInputs, simplex.go:385-439 (shape panics are the caller's GM_ERR_BAD_SHAPE)
inline int verify_inputs(const double* c, const double* A, int lda, const double* b, int m, int n) {
    for (int i = 0; i < m; ++i) {
        bool zero = true;
        for (int j = 0; j < n; ++j)
            if (A[(size_t)i * lda + j] != 0) { zero = false; break; }
        if (zero && b[i] != 0) return GM_ERR_INFEASIBLE;
        if (zero) return GM_ERR_ZERO_ROW;
    }
    for (int j = 0; j < n; ++j) {
        bool zero = true;
        for (int i = 0; i < m; ++i)
            if (A[(size_t)i * lda + j] != 0) { zero = false; break; }
        if (zero && c[j] < 0) return GM_ERR_UNBOUNDED;
        if (zero) return GM_ERR_ZERO_COLUMN;
    }
    return GM_OK;
}

// findLinearlyIndependent, simplex.go:611-637
inline ivec find_linearly_independent(const double* A, int lda, int m, int n, SimplexStats& st) {
    ivec idxs;
    idxs.reserve(m);
    vec columns((size_t)m * m, 0.0);
    for (int i = n - 1; i >= 0; --i) {
        if ((int)idxs.size() == m) break;
        const int k = (int)idxs.size();
        for (int r = 0; r < m; ++r) columns[(size_t)r * m + k] = A[(size_t)r * lda + i];
        if (k == 0) { idxs.push_back(i); continue; }
        st.lu_factorizations++;
        if (cond1(columns.data(), m, m, k + 1) > GM_LINDEP_COND_TOL) continue;
        idxs.push_back(i);
    }
    return idxs;
}

// computeMove, simplex.go:306-342. Returns GM_OK / GM_ERR_LINSOLVE / GM_ERR_UNBOUNDED.
inline int compute_move(double* move, int pos, const double* A, int lda, int m, const double* ab,
                        const double* xb, const int* nonbasic, SimplexStats& st) {
    vec col(m), d(m);
    const int var = nonbasic[pos];
    for (int i = 0; i < m; ++i) col[i] = A[(size_t)i * lda + var];
    st.lu_factorizations++;
    if (solve_vec(ab, m, m, false, col.data(), d.data()) != SolveErr::None) return GM_ERR_LINSOLVE;
    for (int i = 0; i < m; ++i) d[i] *= -1.0;
    for (int i = 0; i < m; ++i)
        if (std::fabs(d[i]) < GM_D_ROUND_TOL) d[i] = 0;
    if (d[min_idx(d.data(), m)] >= 0) return GM_ERR_UNBOUNDED;
    for (int i = 0; i < m; ++i) {
        if (d[i] >= 0) move[i] = std::numeric_limits<double>::infinity();
        else move[i] = xb[i] / std::fabs(d[i]);
    }
    return GM_OK;
}

// replaceBland, simplex.go:347-383
inline int replace_bland(int& replace, int& enter_pos, const double* A, int lda, int m, int nn,
                         const double* ab, const double* xb, const int* basic, const int* nonbasic,
                         const double* r, double* move, SimplexStats& st) {
    vec abtmp((size_t)m * m);
    ivec bi(m);
    for (int i = 0; i < nn; ++i) {
        if (r[i] > -GM_BLAND_NEG_TOL) continue;
        int rc = compute_move(move, i, A, lda, m, ab, xb, nonbasic, st);
        if (rc != GM_OK) return rc;
        int l = min_idx(move, m);
        if (std::fabs(move[l]) > GM_BLAND_ZERO_TOL) { replace = l; enter_pos = i; return GM_OK; }
        for (int p = 0; p < m; ++p) {
            if (move[p] > GM_BLAND_ZERO_TOL) continue;
            for (int q = 0; q < m; ++q) bi[q] = basic[q];
            bi[p] = nonbasic[i];
            extract_columns(abtmp.data(), m, A, lda, m, bi.data(), m);
            st.lu_factorizations++;
            if (cond1(abtmp.data(), m, m, m) < GM_CONDITION_TOL) { replace = p; enter_pos = i; return GM_OK; }
        }
    }
    return GM_ERR_BLAND;
}

}  // namespace detail

SimplexResult simplex_core(const int* initialBasic, const double* c, const double* A, int lda, const double* b,
                           int m, int n, double tol, SimplexStats& st, int phase);

// findInitialBasic, simplex.go:492-607. On success fills basic/ab/xb and returns GM_OK.
inline int find_initial_basic(const double* A, int lda, const double* b, int m, int n, ivec& basic, vec& ab,
                              vec& xb, SimplexStats& st) {
    using namespace detail;
    basic = find_linearly_independent(A, lda, m, n, st);
    if ((int)basic.size() != m) return GM_ERR_SINGULAR;
    ab.assign((size_t)m * m, 0.0);
    extract_columns(ab.data(), m, A, lda, m, basic.data(), m);
    xb.assign(m, 0.0);
    if (initialize_from_basic(xb.data(), ab.data(), m, b, st) == 0) return GM_OK;

    // Phase I: one artificial column that makes the all-ones vector basic-feasible (:529-556)
    st.used_phase1 = 1;
    const int j = min_idx(xb.data(), m);
    vec art(b, b + m);
    for (int i = 0; i < m; ++i) {
        if (i == j) continue;
        const int v = basic[i];
        for (int r = 0; r < m; ++r) art[r] -= A[(size_t)r * lda + v];
    }
    const int n1 = n + 1;
    vec anew((size_t)m * n1);
    for (int r = 0; r < m; ++r) {
        std::memcpy(&anew[(size_t)r * n1], A + (size_t)r * lda, sizeof(double) * n);
        anew[(size_t)r * n1 + n] = art[r];
    }
    basic[j] = n;
    vec c1(n1, 0.0);
    c1[n] = 1;
    SimplexResult p1 = simplex_core(basic.data(), c1.data(), anew.data(), n1, b, m, n1, GM_PHASE1_TOL, st, 1);
    if (p1.status != GM_OK) {
        if (p1.status == GM_PANIC_INITIAL_BASIC || p1.status == GM_ERR_ITERATION_LIMIT) return p1.status;
        return GM_ERR_PHASE1_WRAPPED + p1.status;
    }
    if (std::fabs(p1.x[n]) > GM_PHASE1_ZERO_TOL) return GM_ERR_INFEASIBLE;

    int added = -1;
    ivec& nb = p1.basis;
    for (int i = 0; i < m; ++i) {
        if (nb[i] == n) added = i;
        xb[i] = p1.x[nb[i]];
    }
    if (added == -1) {
        extract_columns(ab.data(), m, A, lda, m, nb.data(), m);
        basic = nb;
        return GM_OK;
    }
    // artificial still basic at level zero: try every non-basic column in its place (:584-606)
    std::vector<char> inb(n1, 0);
    for (int i = 0; i < m; ++i) inb[nb[i]] = 1;
    bool set = false;
    for (int i = 0; i < n1; ++i) {
        if (inb[i]) continue;
        st.repair_trials++;
        nb[added] = i;
        if (set) {
            for (int r = 0; r < m; ++r) ab[(size_t)r * m + added] = A[(size_t)r * lda + i];
        } else {
            extract_columns(ab.data(), m, A, lda, m, nb.data(), m);
            set = true;
        }
        if (initialize_from_basic(xb.data(), ab.data(), m, b, st) == 0) {
            basic = nb;
            return GM_OK;
        }
    }
    return GM_ERR_INFEASIBLE;
}

// simplex, simplex.go:93-302
inline SimplexResult simplex_core(const int* initialBasic, const double* c, const double* A, int lda,
                                  const double* b, int m, int n, double tol, SimplexStats& st, int phase) {
    using namespace detail;
    SimplexResult res;
    const double inf = std::numeric_limits<double>::infinity();
    int v = verify_inputs(c, A, lda, b, m, n);
    if (v != GM_OK) {
        res.status = v;
        res.optF = (v == GM_ERR_UNBOUNDED) ? -inf : std::numeric_limits<double>::quiet_NaN();
        return res;
    }
    if (m == n) {  // :103-119
        vec x(n, 0.0);
        st.lu_factorizations++;
        if (solve_vec(A, lda, n, false, b, x.data()) != SolveErr::None) { res.status = GM_ERR_SINGULAR; return res; }
        for (int i = 0; i < n; ++i)
            if (x[i] < 0) { res.status = GM_ERR_INFEASIBLE; return res; }
        res.optF = dot_unitary(x.data(), c, n);
        res.x = x;
        return res;
    }

    ivec basic;
    vec ab, xb;
    if (initialBasic) {  // :147-160
        basic.assign(initialBasic, initialBasic + m);
        ab.assign((size_t)m * m, 0.0);
        extract_columns(ab.data(), m, A, lda, m, basic.data(), m);
        xb.assign(m, 0.0);
        if (initialize_from_basic(xb.data(), ab.data(), m, b, st) != 0) { res.status = GM_PANIC_INITIAL_BASIC; return res; }
    } else {
        int rc = find_initial_basic(A, lda, b, m, n, basic, ab, xb, st);
        if (rc != GM_OK) { res.status = rc; return res; }
    }

    const int nn = n - m;
    ivec nonbasic;
    nonbasic.reserve(nn);
    {
        std::vector<char> inb(n, 0);
        for (int i = 0; i < m; ++i) inb[basic[i]] = 1;
        for (int i = 0; i < n; ++i)
            if (!inb[i]) nonbasic.push_back(i);
    }
    vec cb(m), cn(nn);
    for (int i = 0; i < m; ++i) cb[i] = c[basic[i]];
    for (int i = 0; i < nn; ++i) cn[i] = c[nonbasic[i]];
    vec an((size_t)m * nn);
    extract_columns(an.data(), nn, A, lda, m, nonbasic.data(), nn);

    vec r(nn), move(m), y(m), data(nn);
    int status = GM_OK;
    long iter = 0;
    for (;;) {
        if (st.max_pivots > 0 && iter >= st.max_pivots) { status = GM_ERR_ITERATION_LIMIT; break; }
        // y = ab^-T cb : LU of the materialised transpose (:236)
        st.lu_factorizations++;
        if (solve_vec(ab.data(), m, m, true, cb.data(), y.data()) != SolveErr::None) { status = GM_ERR_CONDITION; break; }
        // r = cn - an^T y : Dgemv(Trans) as m axpys skipping y_i == 0, then SubTo (:240-243)
        std::fill(data.begin(), data.end(), 0.0);
        for (int i = 0; i < m; ++i) {
            double t = 1.0 * y[i];
            if (t != 0) axpy(nn, t, &an[(size_t)i * nn], 1, data.data(), 1);
        }
        for (int k = 0; k < nn; ++k) r[k] = cn[k] - data[k];

        int e = min_idx(r.data(), nn);
        if (r[e] >= -tol) break;
        for (int k = 0; k < nn; ++k)
            if (std::fabs(r[k]) < GM_R_ROUND_TOL) r[k] = 0;

        int rc = compute_move(move.data(), e, A, lda, m, ab.data(), xb.data(), nonbasic.data(), st);
        if (rc == GM_ERR_UNBOUNDED) { res.status = rc; res.optF = -inf; return res; }
        if (rc != GM_OK) { status = rc; break; }

        int l = min_idx(move.data(), m);
        int bland = 0;
        if (move[l] <= 0) {
            st.bland_calls++;
            bland = 1;
            rc = replace_bland(l, e, A, lda, m, nn, ab.data(), xb.data(), basic.data(), nonbasic.data(), r.data(),
                               move.data(), st);
            if (rc == GM_ERR_UNBOUNDED) { res.status = rc; res.optF = -inf; return res; }
            if (rc != GM_OK) { status = rc; break; }
        }
        if (st.trace) st.trace->push_back({phase, nonbasic[e], basic[l], bland});
        if (phase == 1) st.pivots_phase1++; else st.pivots_phase2++;
        ++iter;

        std::swap(basic[l], nonbasic[e]);
        std::swap(cb[l], cn[e]);
        for (int i = 0; i < m; ++i) std::swap(ab[(size_t)i * m + l], an[(size_t)i * nn + e]);

        st.lu_factorizations++;
        SolveErr se = solve_vec(ab.data(), m, m, false, b, xb.data());
        if (se != SolveErr::None) { status = GM_ERR_CONDITION; break; }
    }
    res.status = status;
    res.optF = dot_unitary(cb.data(), xb.data(), m);
    res.x.assign(n, 0.0);
    for (int i = 0; i < m; ++i) res.x[basic[i]] = xb[i];
    res.basis = basic;
    return res;
}

// lp.Simplex(c, A, b, tol, initialBasic), simplex.go:88-91
inline SimplexResult simplex(const double* c, const double* A, int lda, const double* b, int m, int n, double tol,
                             const int* initialBasic, SimplexStats& st) {
    return simplex_core(initialBasic, c, A, lda, b, m, n, tol, st, 2);
}

}  // namespace orc
