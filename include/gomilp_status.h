/* gomilp_status.h — status taxonomy shared by the CUDA engine (libgomilp_b200.so) and the
 * CPU oracle (oracle/).  One code per error VALUE or PANIC CLASS that the reference's
 * lp.Simplex can produce, so the host side can rebuild an `error` that compares equal (==)
 * to the lp.Err* value GoMILP switches on (ilp.go:37-40, tree.go:77,266-273).
 *
 * Reference: vendor/gonum.org/v1/gonum/optimize/convex/lp/simplex.go:26-34 (error values),
 * :147-160 (panic on bad initialBasic), :557-559 (wrapped Phase-I error),
 * vendor/gonum.org/v1/gonum/mat/errors.go:25-43 (mat.Condition).
 */
#ifndef GOMILP_STATUS_H
#define GOMILP_STATUS_H

#ifdef __cplusplus
extern "C" {
#endif

typedef enum gm_status {
    GM_OK = 0,
    GM_ERR_INFEASIBLE = 1,      /* lp.ErrInfeasible   simplex.go:28  */
    GM_ERR_UNBOUNDED = 2,       /* lp.ErrUnbounded    simplex.go:30  (optF = -Inf) */
    GM_ERR_SINGULAR = 3,        /* lp.ErrSingular     simplex.go:31  */
    GM_ERR_ZERO_ROW = 4,        /* lp.ErrZeroRow      simplex.go:33  */
    GM_ERR_ZERO_COLUMN = 5,     /* lp.ErrZeroColumn   simplex.go:32  */
    GM_ERR_BLAND = 6,           /* lp.ErrBland        simplex.go:27  (last iterate returned) */
    GM_ERR_LINSOLVE = 7,        /* lp.ErrLinSolve     simplex.go:29  (last iterate returned) */
    GM_ERR_CONDITION = 8,       /* mat.Condition from a main-loop SolveVec, simplex.go:236-238,289-292 */
    GM_PANIC_INITIAL_BASIC = 9, /* panic(err) simplex.go:155-158: supplied basis singular/infeasible */
    /* fmt.Errorf("lp: error finding feasible basis: %s", inner) simplex.go:557-559.
     * Encoded as GM_ERR_PHASE1_WRAPPED + inner code (inner in 1..15). */
    GM_ERR_PHASE1_WRAPPED = 32,

    /* engine-level codes: no counterpart in the reference */
    GM_ERR_BAD_SHAPE = 64,       /* the reference panics: simplex.go:387-398 */
    GM_ERR_NO_DEVICE = 65,       /* no CUDA device: there is NO CPU fallback */
    GM_ERR_CUDA = 66,            /* CUDA runtime error, see gm_last_error() */
    GM_ERR_TOO_LARGE = 67,       /* shape exceeds what the selected kernel tier supports */
    GM_ERR_ITERATION_LIMIT = 68, /* safety cap hit (the reference has none and would spin) */
    GM_ERR_BAD_HANDLE = 69,
    GM_ERR_BAD_ARGUMENT = 70,
    GM_ERR_WARM_RETRY = 71       /* internal: a warm-started node must be re-solved cold (never returned to callers) */
} gm_status;

/* Tolerances of the reference, simplex.go:42-58 and the call sites that hard-wire them. */
#define GM_INIT_POS_TOL 1e-13      /* initPosTol    simplex.go:45 */
#define GM_BLAND_NEG_TOL 1e-14     /* blandNegTol   simplex.go:47 */
#define GM_R_ROUND_TOL 1e-13       /* rRoundTol     simplex.go:50 */
#define GM_D_ROUND_TOL 1e-13       /* dRoundTol     simplex.go:53 */
#define GM_PHASE1_ZERO_TOL 1e-12   /* phaseIZeroTol simplex.go:55 */
#define GM_BLAND_ZERO_TOL 1e-12    /* blandZeroTol  simplex.go:57 */
#define GM_PHASE1_TOL 1e-10        /* tol passed to the Phase-I recursion, simplex.go:556 */
#define GM_LINDEP_COND_TOL 1e12    /* findLinearlyIndependent, simplex.go:630 */
#define GM_CONDITION_TOL 1e16      /* mat.ConditionTolerance errors.go:33; replaceBland simplex.go:377 */


/* Branch-and-bound decisions, tree.go:14-23 (strings kept verbatim by the host mirror). */
typedef enum gm_decision {
    GM_DEC_NONE = 0,
    GM_DEC_SUBPROBLEM_IS_DEGENERATE = 1,        /* tree.go:17 — assigned to lp.ErrInfeasible (ilp.go:38, labels swapped) */
    GM_DEC_SUBPROBLEM_NOT_FEASIBLE = 2,         /* tree.go:18 — assigned to lp.ErrSingular   (ilp.go:39) */
    GM_DEC_WORSE_THAN_INCUMBENT = 3,            /* tree.go:19 */
    GM_DEC_BETTER_THAN_INCUMBENT_BRANCHING = 4, /* tree.go:20 */
    GM_DEC_BETTER_THAN_INCUMBENT_FEASIBLE = 5,  /* tree.go:21 */
    GM_DEC_INITIAL_RX_FEASIBLE_FOR_IP = 6       /* tree.go:22 */
} gm_decision;

/* Outcome of a whole MILP solve, ilp.go:75-116 / tree.go:66-123. */
typedef enum gm_milp_status {
    GM_MILP_OK = 0,
    GM_MILP_NO_INTEGER_FEASIBLE_SOLUTION = 1, /* ilp.go:31,102-104 */
    GM_MILP_DEADLINE_EXCEEDED = 2,            /* ctx.Err(), ilp.go:92-99: node/time budget hit; incumbent (if any) returned */
    GM_MILP_PANIC_ROOT = 3,                   /* subproblem.go:173-176: root relaxation failed -> panic(err) */
    GM_MILP_PANIC_SOLVER_FAILURE = 4,         /* tree.go:266-273: unexpected lp error on a child -> panic(err) */
    GM_MILP_PANIC_UNEXPECTED_CASE = 5,        /* tree.go:253-256: NaN objective falls through the switch */
    GM_MILP_ENGINE_ERROR = 6                  /* CUDA / engine failure (no reference counterpart) */
} gm_milp_status;

/* Branching heuristics, branching.go:6-12. */
typedef enum gm_branch_heuristic {
    GM_BRANCH_MAXFUN = 0,
    GM_BRANCH_MOST_INFEASIBLE = 1,
    GM_BRANCH_NAIVE = 2
} gm_branch_heuristic;

/* Replay mode of the B&B host. COMPAT reproduces the reference including SURVEY App. B quirks
 * 1-4 (heuristic never propagated => always MAXFUN, which always returns the last integer-flagged
 * index). FIXED propagates the heuristic and evaluates it on the fractional integer variables of x
 * so that trees terminate; it exists for throughput runs and has no reference counterpart. */
typedef enum gm_bnb_mode { GM_BNB_COMPAT = 0, GM_BNB_FIXED = 1 } gm_bnb_mode;
/* OR-ed into `mode`: children start from their parent's optimal basis (gm_solve_wave_warm) instead of from
 * scratch. Same optima where they are unique, far fewer pivots; not a pivot-for-pivot replay of the reference. */
#define GM_BNB_WARM_START 4
/* OR-ed into `mode`: checkSolution / feasibleForIP / branch run on the device (node_check + wave_scan kernels) so that
 * x never leaves the GPU, and the wave is sharded over the ranks of gm_comm_init when a communicator is set. Same
 * decisions, ids and node counts as the host replay (gm_milp_solve without the flag). GM_BNB_WARM_START is honoured
 * on one GPU (the parents' inverses stay in that GPU's HBM); with a communicator children are solved cold. */
#define GM_BNB_DEVICE_SCAN 8
/* OR-ed into `mode`: node LPs are solved with gm_options.robust = 1 (see gomilp_b200.h). The reference panics on the
 * first relaxation its simplex cannot finish (tree.go:272); with this flag degenerate searches (0-1 knapsacks with
 * bounds as rows) run through. No reference counterpart, like GM_BNB_FIXED. */
#define GM_BNB_ROBUST 16

#ifdef __cplusplus
}
#endif
#endif /* GOMILP_STATUS_H */
