"""Status taxonomy of include/gomilp_status.h (one code per lp.Err* value / panic class of the reference:
vendor/gonum.org/v1/gonum/optimize/convex/lp/simplex.go:26-34; decisions: tree.go:14-23)."""
GM_OK = 0
GM_ERR_INFEASIBLE = 1
GM_ERR_UNBOUNDED = 2
GM_ERR_SINGULAR = 3
GM_ERR_ZERO_ROW = 4
GM_ERR_ZERO_COLUMN = 5
GM_ERR_BLAND = 6
GM_ERR_LINSOLVE = 7
GM_ERR_CONDITION = 8
GM_PANIC_INITIAL_BASIC = 9
GM_ERR_PHASE1_WRAPPED = 32
GM_ERR_BAD_SHAPE = 64
GM_ERR_NO_DEVICE = 65
GM_ERR_CUDA = 66
GM_ERR_TOO_LARGE = 67
GM_ERR_ITERATION_LIMIT = 68
GM_ERR_BAD_HANDLE = 69
GM_ERR_BAD_ARGUMENT = 70

GM_DEC_NONE = 0
GM_DEC_SUBPROBLEM_IS_DEGENERATE = 1
GM_DEC_SUBPROBLEM_NOT_FEASIBLE = 2
GM_DEC_WORSE_THAN_INCUMBENT = 3
GM_DEC_BETTER_THAN_INCUMBENT_BRANCHING = 4
GM_DEC_BETTER_THAN_INCUMBENT_FEASIBLE = 5
GM_DEC_INITIAL_RX_FEASIBLE_FOR_IP = 6
DECISION_NAMES = {  # tree.go:17-22, verbatim
    1: "subproblem contains a degenerate (singular) matrix",
    2: "subproblem has no feasible solution",
    3: "worse than incumbent",
    4: "better than incumbent but not integer feasible, so branching",
    5: "better than incumbent and integer feasible, so replacing incumbent",
    6: "initial relaxation is feasible for IP",
}

GM_MILP_OK = 0
GM_MILP_NO_INTEGER_FEASIBLE_SOLUTION = 1
GM_MILP_DEADLINE_EXCEEDED = 2
GM_MILP_PANIC_ROOT = 3
GM_MILP_PANIC_SOLVER_FAILURE = 4
GM_MILP_PANIC_UNEXPECTED_CASE = 5
GM_MILP_ENGINE_ERROR = 6

GM_BRANCH_MAXFUN = 0
GM_BRANCH_MOST_INFEASIBLE = 1
GM_BRANCH_NAIVE = 2
GM_BNB_COMPAT = 0
GM_BNB_FIXED = 1
GM_BNB_WARM_START = 4  # OR-ed into mode: children start from the parent's optimal basis (not a pivot-for-pivot replay)
GM_BNB_ROBUST = 16  # OR-ed into mode: node LPs the reference's rules cannot finish are re-solved on a perturbed rhs
GM_BNB_DEVICE_SCAN = 8  # OR-ed into mode: checkSolution / branch on the device, sharded over gm_comm_init's ranks

STATUS_NAMES = {
    0: "ok", 1: "lp: problem is infeasible", 2: "lp: problem is unbounded", 3: "lp: A is singular",
    4: "lp: A has a row with all zeros", 5: "lp: A has a column with all zeros",
    6: "lp: bland: all replacements are negative or cause ill-conditioned ab",
    7: "lp: linear solve failure", 8: "matrix singular or near-singular with condition number",
    9: "panic: initial basic", 64: "bad shape", 65: "no CUDA device", 66: "CUDA error", 67: "too large",
    68: "iteration limit", 69: "bad handle", 70: "bad argument",
}
