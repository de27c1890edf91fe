// oracle/gomilp_bnb.hpp — TEST INFRASTRUCTURE ONLY (CPU oracle; see oracle/README.md).
//
// Serial restatement of GoMILP's branch-and-bound around lp.Simplex, as the 1-worker FIFO order of
// the reference executes it (all paths relative to /root/reference/):
//   ilp.go:43-116 (toInitialSubproblem, solve), subproblem.go:55-259 (combineInequalities,
//   convertToEqualities, solve, branch, getChild), tree.go:66-297 (startSearch, checkSolution,
//   translateSolverFailure, feasibleForIP), branching.go:17-94 (the three heuristics, bugs kept).
// With one worker the reference solves nodes in FIFO order and checks each candidate before the next
// one is produced (tree.go:103-115,196-205), which is what the plain queue below does.
#pragma once
#include <chrono>
#include <deque>

#include "lp_simplex.hpp"

namespace orc {

struct MilpProblem {  // ilp.go:11-27
    int nvar = 0;
    vec c;
    int meq = 0;
    vec A;  // meq × nvar, row-major (may be empty)
    vec b;
    int nineq = 0;
    vec G;  // nineq × nvar
    vec h;
    std::vector<char> integrality;
    int heuristic = GM_BRANCH_MAXFUN;
};

struct BnbConstraint {  // subproblem.go:36-44 with gsharp kept implicit (one ±1 at `var`)
    int var;
    double sign;    // gsharp[var]
    double hsharp;
};

struct NodeRecord {
    long id = 0, parent = 0;
    int depth = 0;
    int lp_status = GM_OK;
    double z = 0;
    int decision = GM_DEC_NONE;
    long pivots = 0;
    int branch_var = -1;     // variable branched on when decision == BRANCHING
    double branch_floor = 0;
};

struct BnbResult {
    int status = GM_MILP_OK;
    int lp_status = GM_OK;  // offending lp status for the PANIC_* outcomes
    vec x;                  // length nvar (slacks stripped, ilp.go:111-112); empty if none
    double z = 0;
    long nodes = 0;  // LP relaxations solved (root included)
    long pivots = 0;
    std::vector<NodeRecord> log;  // in ProcessDecision order (tree.go:261)
};

// convertToEqualities, subproblem.go:81-139 : [A 0; G I], c' = [c;0], b' = [b;h]
inline void convert_to_equalities(const vec& c, const vec& A, const vec& b, int meq, const vec& G, const vec& h,
                                  int nineq, vec& cNew, vec& aNew, vec& bNew) {
    const int nvar = (int)c.size();
    const int nn = nvar + nineq, mm = meq + nineq;
    cNew.assign(nn, 0.0);
    std::copy(c.begin(), c.end(), cNew.begin());
    bNew.assign(mm, 0.0);
    std::copy(b.begin(), b.begin() + meq, bNew.begin());
    std::copy(h.begin(), h.begin() + nineq, bNew.begin() + meq);
    aNew.assign((size_t)mm * nn, 0.0);
    for (int i = 0; i < meq; ++i)
        for (int j = 0; j < nvar; ++j) aNew[(size_t)i * nn + j] = A[(size_t)i * nvar + j];
    for (int i = 0; i < nineq; ++i) {
        for (int j = 0; j < nvar; ++j) aNew[(size_t)(meq + i) * nn + j] = G[(size_t)i * nvar + j];
        aNew[(size_t)(meq + i) * nn + nvar + i] = 1;
    }
}

// tree.go:276-297 : exact x == trunc(x) on the integer-flagged entries
inline bool feasible_for_ip(const std::vector<char>& integrality, const vec& x) {
    for (size_t i = 0; i < x.size(); ++i)
        if (integrality[i] && !(x[i] == std::trunc(x[i]))) return false;
    return true;
}

// branching.go:54-72 — candidateValue is never updated, so the LAST integer-flagged index whose
// |c_i| >= 0 (i.e. non-NaN) wins.
inline int maxfun_branch_point(const vec& c, const std::vector<char>& integrality) {
    double candidate = 0;
    int cur = 0;
    for (size_t i = 0; i < c.size(); ++i)
        if (integrality[i] && std::fabs(c[i]) >= candidate) cur = (int)i;
    return cur;
}

// branching.go:75-94 — called with c (subproblem.go:202), candidateRemainder never updated.
inline int most_infeasible_branch_point(const vec& c, const std::vector<char>& integrality) {
    const double remainder = 1.0;
    int cur = 0;
    for (size_t i = 0; i < c.size(); ++i) {
        if (!integrality[i]) continue;
        double ip;
        double f = std::modf(c[i], &ip);
        if ((0.5 - f) <= remainder) cur = (int)i;
    }
    return cur;
}

// branching.go:17-51
inline int naive_branch_point(const std::vector<char>& integrality, const std::vector<BnbConstraint>& bnb, int ncols) {
    int on = 0;
    if (bnb.empty()) {
        for (size_t i = 0; i < integrality.size(); ++i)
            if (integrality[i]) on = (int)i;
        return on;
    }
    int cursor = bnb.back().var;
    for (int guard = 0; guard <= 2 * ncols; ++guard) {  // the reference spins forever without an integer var
        if (cursor == ncols - 1) cursor = -1;
        ++cursor;
        if (integrality[cursor]) return cursor;
    }
    return on;
}

// FIXED-mode heuristics (no reference counterpart; see gm_bnb_mode): evaluated on the integer
// variables that are fractional in x. Ties resolve to the lowest index.
inline int fixed_branch_point(int heuristic, const vec& c, const vec& x, const std::vector<char>& integrality,
                              const std::vector<BnbConstraint>& bnb) {
    int best = -1;
    double bestv = -1;
    const int n = (int)x.size();
    if (heuristic == GM_BRANCH_NAIVE) {
        int start = bnb.empty() ? 0 : (bnb.back().var + 1) % n;
        for (int k = 0; k < n; ++k) {
            int i = (start + k) % n;
            if (integrality[i] && x[i] != std::trunc(x[i])) return i;
        }
        return -1;
    }
    for (int i = 0; i < n; ++i) {
        if (!integrality[i] || x[i] == std::trunc(x[i])) continue;
        double score;
        if (heuristic == GM_BRANCH_MOST_INFEASIBLE) {
            double f = x[i] - std::floor(x[i]);
            score = 0.5 - std::fabs(0.5 - f);
        } else {
            score = std::fabs(c[i]);
        }
        if (score > bestv) { bestv = score; best = i; }
    }
    return best;
}

struct BnbOptions {
    int mode = GM_BNB_COMPAT;
    long node_limit = 0;      // 0 = unlimited; stands in for the context deadline
    double time_limit_s = 0;  // 0 = unlimited
    long max_pivots_per_lp = 0;
    bool keep_log = true;
};

inline BnbResult bnb_solve(const MilpProblem& p, const BnbOptions& opt) {
    BnbResult out;
    const auto t0 = std::chrono::steady_clock::now();
    // toInitialSubproblem, ilp.go:43-71
    vec c0 = p.c, A0 = p.A, b0 = p.b;
    std::vector<char> integ = p.integrality;
    int m0 = p.meq, n0 = p.nvar;
    if (p.nineq > 0) {
        convert_to_equalities(p.c, p.A, p.b, p.meq, p.G, p.h, p.nineq, c0, A0, b0);
        m0 = p.meq + p.nineq;
        n0 = p.nvar + p.nineq;
        integ.assign(n0, 0);
        std::copy(p.integrality.begin(), p.integrality.end(), integ.begin());
    }

    struct Node { long id, parent; std::vector<BnbConstraint> bnb; };
    struct Sol { Node node; int status; double z; vec x; long pivots; };

    auto solve_node = [&](const Node& nd) {
        Sol s{nd, GM_OK, 0, {}, 0};
        SimplexStats st;
        st.max_pivots = opt.max_pivots_per_lp;
        const int L = (int)nd.bnb.size();
        SimplexResult r;
        if (L > 0) {  // subproblem.go:150-159
            const int m = m0 + L, n = n0 + L;
            vec c(n, 0.0), b(m, 0.0), A((size_t)m * n, 0.0);
            std::copy(c0.begin(), c0.end(), c.begin());
            std::copy(b0.begin(), b0.end(), b.begin());
            for (int i = 0; i < m0; ++i)
                for (int j = 0; j < n0; ++j) A[(size_t)i * n + j] = A0[(size_t)i * n0 + j];
            for (int k = 0; k < L; ++k) {
                A[(size_t)(m0 + k) * n + nd.bnb[k].var] = nd.bnb[k].sign;
                A[(size_t)(m0 + k) * n + n0 + k] = 1;
                b[m0 + k] = nd.bnb[k].hsharp;
            }
            r = simplex(c.data(), A.data(), n, b.data(), m, n, 0.0, nullptr, st);
            if (r.status == GM_OK) r.x.resize(n0);
        } else {
            r = simplex(c0.data(), A0.data(), n0, b0.data(), m0, n0, 0.0, nullptr, st);
        }
        s.status = r.status;
        s.z = r.optF;
        s.x = r.x;
        s.pivots = st.pivots_phase1 + st.pivots_phase2;
        out.nodes++;
        out.pivots += s.pivots;
        return s;
    };

    auto record = [&](const Sol& s, int decision, int bvar = -1, double bfloor = 0) {
        if (!opt.keep_log) return;
        NodeRecord nr;
        nr.id = s.node.id;
        nr.parent = s.node.parent;
        nr.depth = (int)s.node.bnb.size();
        nr.lp_status = s.status;
        nr.z = s.z;
        nr.decision = decision;
        nr.pivots = s.pivots;
        nr.branch_var = bvar;
        nr.branch_floor = bfloor;
        out.log.push_back(nr);
    };

    Node root{0, 0, {}};
    Sol rs = solve_node(root);
    if (rs.status != GM_OK) {  // subproblem.go:173-176 panics on ANY root error
        out.status = GM_MILP_PANIC_ROOT;
        out.lp_status = rs.status;
        return out;
    }
    auto strip = [&](const vec& x) { return vec(x.begin(), x.begin() + p.nvar); };
    if (feasible_for_ip(integ, rs.x)) {  // tree.go:88-92
        record(rs, GM_DEC_INITIAL_RX_FEASIBLE_FOR_IP);
        out.x = strip(rs.x);
        out.z = rs.z;
        return out;
    }

    std::deque<Node> fifo;
    long next_id = 0;
    bool have_inc = false;
    Sol incumbent;
    int panic = 0, panic_lp = 0;

    auto check = [&](const Sol& cand) {  // tree.go:207-263
        const double incZ = have_inc ? incumbent.z : std::numeric_limits<double>::infinity();
        if (cand.status != GM_OK) {
            if (cand.status == GM_ERR_INFEASIBLE) { record(cand, GM_DEC_SUBPROBLEM_IS_DEGENERATE); return; }
            if (cand.status == GM_ERR_SINGULAR) { record(cand, GM_DEC_SUBPROBLEM_NOT_FEASIBLE); return; }
            panic = GM_MILP_PANIC_SOLVER_FAILURE;
            panic_lp = cand.status;
            return;
        }
        if (incZ <= cand.z) { record(cand, GM_DEC_WORSE_THAN_INCUMBENT); return; }
        if (incZ > cand.z) {
            if (feasible_for_ip(integ, cand.x)) {
                incumbent = cand;
                have_inc = true;
                record(cand, GM_DEC_BETTER_THAN_INCUMBENT_FEASIBLE);
                return;
            }
            int on;
            if (opt.mode == GM_BNB_COMPAT) {
                on = maxfun_branch_point(c0, integ);  // heuristic never propagated: always MAXFUN (App. B-3)
            } else {
                on = fixed_branch_point(p.heuristic, c0, cand.x, integ, cand.node.bnb);
                if (on < 0) on = maxfun_branch_point(c0, integ);
            }
            const double fl = std::floor(cand.x[on]);
            Node c1{0, cand.node.id, cand.node.bnb}, c2{0, cand.node.id, cand.node.bnb};
            c1.bnb.push_back({on, 1.0, fl});
            c2.bnb.push_back({on, -1.0, -(fl + 1)});
            c1.id = ++next_id;
            c2.id = ++next_id;
            fifo.push_back(std::move(c1));
            fifo.push_back(std::move(c2));
            record(cand, GM_DEC_BETTER_THAN_INCUMBENT_BRANCHING, on, fl);
            return;
        }
        panic = GM_MILP_PANIC_UNEXPECTED_CASE;  // NaN objective
        panic_lp = cand.status;
    };

    check(rs);
    bool timed_out = false;
    while (!panic && !fifo.empty()) {
        if (opt.node_limit > 0 && out.nodes >= opt.node_limit) { timed_out = true; break; }
        if (opt.time_limit_s > 0) {
            double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (el > opt.time_limit_s) { timed_out = true; break; }
        }
        Node nd = std::move(fifo.front());
        fifo.pop_front();
        Sol s = solve_node(nd);
        check(s);
    }
    if (panic) {
        out.status = panic;
        out.lp_status = panic_lp;
        return out;
    }
    if (timed_out) {  // ilp.go:92-99
        out.status = GM_MILP_DEADLINE_EXCEEDED;
        if (have_inc) { out.x = incumbent.x; out.z = incumbent.z; }  // `val = *incumbent`: slacks NOT stripped
        return out;
    }
    if (!have_inc) { out.status = GM_MILP_NO_INTEGER_FEASIBLE_SOLUTION; return out; }
    out.x = strip(incumbent.x);
    out.z = incumbent.z;
    return out;
}

}  // namespace orc
