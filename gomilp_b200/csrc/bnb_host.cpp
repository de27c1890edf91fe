// bnb_host.cpp — gm_milp_solve: GoMILP's branch-and-bound as a batched wavefront over the C ABI.
//
// Host mirror (C++; the image has no Go toolchain) of, relative to /root/reference:
//   milpProblem.toInitialSubproblem / solve      ilp.go:43-116
//   enumerationTree.startSearch / checkSolution  tree.go:66-123, 207-263 (FIFO queue -> one wave per BFS level)
//   translateSolverFailure                       tree.go:266-273, expectedFailures ilp.go:34-41
//   feasibleForIP / isAllInteger                 tree.go:276-297 (exact x == trunc(x))
//   solution.branch / getChild                   subproblem.go:193-259
//   the three branching heuristics               branching.go:17-94 (their observable behaviour, bugs included)
//   BnbMiddleware                                instrumentation.go:8-15 (gm_decision_cb; gm_wave_cb adds device timings)
// Every LP relaxation is solved by gm_solve_wave on the GPU; this file only schedules and decides.
//
// Why a wave replays the reference: with one worker the reference solves nodes in FIFO order and
// checks each candidate before the next (tree.go:103-115,196-205); the LP solve never reads the
// incumbent (subproblem.go:141-187), so solving a whole BFS level at once and then running
// checkSolution over the results in FIFO order yields the same decisions, ids and node count.
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "../../include/gomilp_b200.h"

namespace {

struct Wave {
    int64_t L = 0;
    std::vector<int64_t> id, parent;
    std::vector<int32_t> parent_idx;  // position of the parent in the previous wave (warm start)
    std::vector<int32_t> bvar;  // [nodes][L]
    std::vector<double> bsign, brhs;
    size_t nodes() const { return id.size(); }
};

// tree.go:276-297: exact x == trunc(x). itol > 0 (warm-start mode only, which is not a replay) accepts values
// within itol of an integer: a warm-started x carries 1e-16-level noise that the exact test would branch on forever.
bool is_integral(double v, double itol) {
    if (itol <= 0.0) return v == std::trunc(v);
    return std::fabs(v - std::nearbyint(v)) <= itol;
}
bool feasible_for_ip(const uint8_t* integ, const double* x, int64_t n, double itol) {
    for (int64_t i = 0; i < n; ++i)
        if (integ[i] && !is_integral(x[i], itol)) return false;
    return true;
}

// branching.go:54-72: the running maximum is never stored back, so every integer-flagged index whose
// |c_i| >= 0 (any non-NaN) overwrites the choice and the last one wins.
int64_t maxfun_point(const double* c, const uint8_t* integ, int64_t n) {
    int64_t pick = 0;
    for (int64_t i = 0; i < n; ++i)
        if (integ[i] && std::fabs(c[i]) >= 0.0) pick = i;
    return pick;
}

// FIXED mode (gm_bnb_mode): heuristics evaluated on the fractional integer variables of x.
int64_t fixed_point(int heuristic, const double* c, const double* x, const uint8_t* integ, int64_t n,
                    int64_t last_var, double itol) {
    if (heuristic == GM_BRANCH_NAIVE) {
        const int64_t start = last_var < 0 ? 0 : (last_var + 1) % n;
        for (int64_t k = 0; k < n; ++k) {
            const int64_t i = (start + k) % n;
            if (integ[i] && !is_integral(x[i], itol)) return i;
        }
        return -1;
    }
    int64_t best = -1;
    double bestv = -1;
    for (int64_t i = 0; i < n; ++i) {
        if (!integ[i] || is_integral(x[i], itol)) continue;
        double score;
        if (heuristic == GM_BRANCH_MOST_INFEASIBLE) {
            const double f = x[i] - std::floor(x[i]);
            score = 0.5 - std::fabs(0.5 - f);
        } else {
            score = std::fabs(c[i]);
        }
        if (score > bestv) { bestv = score; best = i; }
    }
    return best;
}

}  // namespace

extern "C" int gm_milp_solve(int64_t nvar, const double* c, int64_t meq, const double* A, const double* b,
                             int64_t nineq, const double* G, const double* h, const uint8_t* integrality,
                             int32_t heuristic, int32_t mode, int64_t node_limit, double time_limit_s, double* x_out,
                             gm_milp_result* result, gm_decision_cb on_decision, gm_wave_cb on_wave, void* user) {
    if (mode & GM_BNB_DEVICE_SCAN)  // decisions on the device, sharded when a communicator is set (bnb_device.cu)
        return gm_milp_solve_device(nvar, c, meq, A, b, nineq, G, h, integrality, heuristic, mode & ~GM_BNB_DEVICE_SCAN,
                                    node_limit, time_limit_s, x_out, result, on_decision, on_wave, user);
    if (!result || !x_out || !c || !integrality || nvar <= 0 || meq < 0 || nineq < 0) return GM_ERR_BAD_ARGUMENT;
    if ((meq > 0 && (!A || !b)) || (nineq > 0 && (!G || !h)) || meq + nineq == 0) return GM_ERR_BAD_ARGUMENT;
    const auto t0 = std::chrono::steady_clock::now();
    std::memset(result, 0, sizeof(*result));

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

    gm_root_t root = 0;
    int rc = gm_upload_root(c0.data(), A0.data(), n0, b0.data(), m0, n0, &root);
    if (rc != GM_OK) { result->status = GM_MILP_ENGINE_ERROR; result->lp_status = rc; return rc; }

    const double inf = std::numeric_limits<double>::infinity();
    bool have_inc = false;
    double inc_z = inf;
    std::vector<double> inc_x;
    int64_t next_id = 0;
    int panic = 0, panic_lp = 0;
    bool timed_out = false;

    const bool warm = (mode & GM_BNB_WARM_START) != 0;
    const double itol = warm ? 1e-9 : 0.0;
    struct RobustScope {  // GM_BNB_ROBUST: this thread's waves solve with gm_options.robust until we return
        bool on;
        explicit RobustScope(bool o) : on(o) { if (on) gm_thread_robust(+1); }
        ~RobustScope() { if (on) gm_thread_robust(-1); }
    } robust_scope((mode & GM_BNB_ROBUST) != 0);
    mode &= 3;
    Wave cur;
    cur.L = 0;
    cur.id.push_back(0);
    cur.parent.push_back(0);
    cur.parent_idx.push_back(-1);
    int64_t wave_no = 0;

    std::vector<int32_t> status, stats;
    std::vector<double> z, x;
    while (cur.nodes() > 0 && !panic) {
        size_t count = cur.nodes();
        if (node_limit > 0) {  // the context deadline of ilp.go:92-99, expressed as a node budget
            const int64_t left = node_limit - result->nodes;
            if (left <= 0) { timed_out = true; break; }
            if ((int64_t)count > left) { count = (size_t)left; timed_out = true; }
        }
        if (time_limit_s > 0 && wave_no > 0) {
            const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (el > time_limit_s) { timed_out = true; break; }
        }
        status.assign(count, 0);
        stats.assign(count * 8, 0);
        z.assign(count, 0.0);
        x.assign(count * (size_t)n0, 0.0);
        rc = warm ? gm_solve_wave_warm(root, (int64_t)count, cur.L, cur.bvar.data(), cur.bsign.data(),
                                       cur.brhs.data(), cur.parent_idx.data(), status.data(), z.data(), x.data(),
                                       nullptr, stats.data())
                  : gm_solve_wave(root, (int64_t)count, cur.L, cur.bvar.data(), cur.bsign.data(), cur.brhs.data(),
                                  status.data(), z.data(), x.data(), nullptr, stats.data());
        if (rc != GM_OK) {
            gm_free_root(root);
            result->status = GM_MILP_ENGINE_ERROR;
            result->lp_status = rc;
            return rc;
        }
        gm_timing tm;
        gm_last_timing(&tm);
        int64_t wave_pivots = 0;
        for (size_t k = 0; k < count; ++k) wave_pivots += (int64_t)stats[k * 8] + stats[k * 8 + 1];
        result->waves += 1;
        result->device_ms += tm.kernel_ms;
        if (on_wave) on_wave(user, wave_no, (int64_t)count, wave_pivots, tm.kernel_ms);

        Wave next;
        next.L = cur.L + 1;
        // checkSolution over the wave in FIFO order (tree.go:207-263)
        for (size_t k = 0; k < count && !panic; ++k) {
            result->nodes += 1;
            result->pivots += (int64_t)stats[k * 8] + stats[k * 8 + 1];
            const int st = status[k];
            const double zk = z[k];
            const double* xk = &x[k * (size_t)n0];
            int decision = GM_DEC_NONE;
            int32_t bv = -1;
            double bfloor = 0;
            if (wave_no == 0) {
                if (st != GM_OK) {  // subproblem.go:173-176: any root failure panics
                    panic = GM_MILP_PANIC_ROOT;
                    panic_lp = st;
                    break;
                }
                if (feasible_for_ip(integ.data(), xk, n0, itol)) {  // tree.go:88-92
                    if (on_decision)
                        on_decision(user, 0, 0, 0, st, zk, GM_DEC_INITIAL_RX_FEASIBLE_FOR_IP, -1, 0.0);
                    have_inc = true;
                    inc_z = zk;
                    inc_x.assign(xk, xk + n0);
                    break;
                }
            }
            const double incumbent_z = have_inc ? inc_z : inf;
            if (st != GM_OK) {  // translateSolverFailure, tree.go:266-273 with ilp.go:37-40
                if (st == GM_ERR_INFEASIBLE) decision = GM_DEC_SUBPROBLEM_IS_DEGENERATE;
                else if (st == GM_ERR_SINGULAR) decision = GM_DEC_SUBPROBLEM_NOT_FEASIBLE;
                else { panic = GM_MILP_PANIC_SOLVER_FAILURE; panic_lp = st; break; }
            } else if (incumbent_z <= zk) {
                decision = GM_DEC_WORSE_THAN_INCUMBENT;
            } else if (incumbent_z > zk) {
                if (feasible_for_ip(integ.data(), xk, n0, itol)) {
                    have_inc = true;
                    inc_z = zk;
                    inc_x.assign(xk, xk + n0);
                    decision = GM_DEC_BETTER_THAN_INCUMBENT_FEASIBLE;
                } else {
                    // solution.branch, subproblem.go:193-221. COMPAT: branchHeuristic is never propagated
                    // (ilp.go:59-70, subproblem.go:232-240), so it is always BRANCH_MAXFUN.
                    int64_t on;
                    if (mode == GM_BNB_COMPAT) {
                        on = maxfun_point(c0.data(), integ.data(), n0);
                    } else {
                        const int64_t last = cur.L > 0 ? cur.bvar[k * cur.L + cur.L - 1] : -1;
                        on = fixed_point(heuristic, c0.data(), xk, integ.data(), n0, last, itol);
                        if (on < 0) on = maxfun_point(c0.data(), integ.data(), n0);
                    }
                    const double fl = std::floor(xk[on]);
                    for (int child = 0; child < 2; ++child) {  // getChild, subproblem.go:230-259
                        next.id.push_back(++next_id);
                        next.parent.push_back(cur.id[k]);
                        next.parent_idx.push_back((int32_t)k);
                        for (int64_t l = 0; l < cur.L; ++l) {
                            next.bvar.push_back(cur.bvar[k * cur.L + l]);
                            next.bsign.push_back(cur.bsign[k * cur.L + l]);
                            next.brhs.push_back(cur.brhs[k * cur.L + l]);
                        }
                        next.bvar.push_back((int32_t)on);
                        next.bsign.push_back(child == 0 ? 1.0 : -1.0);     // x_on <= floor | -x_on <= -(floor+1)
                        next.brhs.push_back(child == 0 ? fl : -(fl + 1.0));
                    }
                    decision = GM_DEC_BETTER_THAN_INCUMBENT_BRANCHING;
                    bv = (int32_t)on;
                    bfloor = fl;
                }
            } else {  // NaN objective falls through the switch: tree.go:253-256 panics
                panic = GM_MILP_PANIC_UNEXPECTED_CASE;
                panic_lp = st;
                break;
            }
            if (on_decision) on_decision(user, cur.id[k], cur.parent[k], (int32_t)cur.L, st, zk, decision, bv, bfloor);
        }
        if (wave_no == 0 && have_inc && next.nodes() == 0 && !panic) {
            // root was integer feasible: startSearch returns it directly
            break;
        }
        if (timed_out) break;
        cur = std::move(next);
        ++wave_no;
    }
    gm_free_root(root);

    if (panic) {
        result->status = panic;
        result->lp_status = panic_lp;
        return GM_OK;
    }
    if (timed_out) {  // ilp.go:92-99 returns *incumbent as is: slack entries are NOT stripped
        result->status = GM_MILP_DEADLINE_EXCEEDED;
        if (have_inc) {
            std::memcpy(x_out, inc_x.data(), sizeof(double) * n0);
            result->x_len = n0;
            result->z = inc_z;
        }
        return GM_OK;
    }
    if (!have_inc) {
        result->status = GM_MILP_NO_INTEGER_FEASIBLE_SOLUTION;
        return GM_OK;
    }
    std::memcpy(x_out, inc_x.data(), sizeof(double) * nvar);  // ilp.go:111-112 drops the slack columns
    result->x_len = nvar;
    result->z = inc_z;
    result->status = GM_MILP_OK;
    return GM_OK;
}
