"""CPU: the oracle (oracle/, C++ restatement of Gonum lp.Simplex + GoMILP B&B) against every known-answer
vector the reference's own tests hold for the path (tests/golden/reference_pins.json), and
against HiGHS (scipy) standing in for the reference's dead GLPK comparison (api_glpk_compare_test.go:216)."""
import numpy as np
import pytest

import oracle
from gomilp_b200 import status as S
from problems import feasible_bounded_lp, raw_lp, reference_pins

PINS = reference_pins()
MILP_STATUS = {"OK": S.GM_MILP_OK, "DEADLINE": S.GM_MILP_DEADLINE_EXCEEDED,
               "NO_INTEGER_FEASIBLE_SOLUTION": S.GM_MILP_NO_INTEGER_FEASIBLE_SOLUTION}


@pytest.mark.parametrize("case", PINS["milp"], ids=[c["src"].split(" ")[0] for c in PINS["milp"]])
def test_oracle_reproduces_reference_milp_pins_exactly(case):
    r = oracle.bnb_solve(case["c"], case["A"], case["b"], case["G"], case["h"], case["integrality"],
                         node_limit=2000)
    assert r.status == MILP_STATUS[case["want_status"]]
    if case["want_status"] == "OK":
        # the reference test uses exact float64 equality (ilp_test.go:296-303)
        assert r.x.tolist() == [float(v) for v in case["want_x"]]
        assert r.z == float(case["want_z"])


def test_oracle_convert_to_equalities_layout():
    p = PINS["convert_to_equalities"]
    c, A, b = oracle.convert_to_equalities(p["c"], p["A"], p["b"], p["G"], p["h"])
    assert c.tolist() == p["want_c"] and A.tolist() == p["want_A"] and b.tolist() == p["want_b"]


def test_oracle_singular_16x14():
    p = PINS["singular_16x14"]
    r = oracle.simplex(p["c"], p["A"], p["b"])
    assert r.status == S.GM_ERR_SINGULAR and r.x is None and np.isnan(r.optF)


def test_oracle_api_end_to_end_values():
    p = PINS["api_end_to_end"]
    r = oracle.bnb_solve(p["c"], p["A"], p["b"], p["G"], p["h"], p["integrality"])
    assert r.status == S.GM_MILP_OK and r.x.tolist() == p["want_x"]


def test_oracle_branching_and_integrality_pins():
    for k in PINS["branching_maxfun"]["cases"]:
        assert oracle.maxfun_branch_point(k["c"], k["integrality"]) == k["want"]
    for k in PINS["branching_most_infeasible"]["cases"]:
        assert oracle.most_infeasible_branch_point(k["c"], k["integrality"]) == k["want"]
    for k in PINS["feasible_for_ip"]["cases"]:
        assert oracle.feasible_for_ip(k["integrality"], k["x"]) == k["want"]


def test_oracle_status_taxonomy_on_degenerate_inputs():
    A = np.array([[1.0, 2.0, 3.0], [0.0, 0.0, 0.0]])
    assert oracle.simplex([1, 1, 1], A, [1, 1]).status == S.GM_ERR_INFEASIBLE   # zero row, b != 0
    assert oracle.simplex([1, 1, 1], A, [1, 0]).status == S.GM_ERR_ZERO_ROW
    A = np.array([[1.0, 0.0, 3.0], [2.0, 0.0, 1.0]])
    assert oracle.simplex([1, -1, 1], A, [1, 1]).status == S.GM_ERR_UNBOUNDED   # zero column, c < 0
    assert oracle.simplex([1, 1, 1], A, [1, 1]).status == S.GM_ERR_ZERO_COLUMN
    r = oracle.simplex([-1.0, 0.0], np.array([[1.0, -1.0]]), [1.0])
    assert r.status == S.GM_ERR_UNBOUNDED and r.optF == -np.inf and r.x is None


def test_oracle_agrees_with_highs_on_random_lps():
    from scipy.optimize import linprog
    rng = np.random.default_rng(155)
    checked = 0
    for _ in range(60):
        m = int(rng.integers(2, 12))
        n = int(rng.integers(m + 1, 2 * m + 6))
        c, A, b = feasible_bounded_lp(rng, m, n)
        r = oracle.simplex(c, A, b)
        h = linprog(c, A_eq=A, b_eq=b, bounds=(0, None), method="highs")
        assert r.status == S.GM_OK and h.status == 0
        assert abs(r.optF - h.fun) <= 0.005  # the GLPK test's tolerance, api_glpk_compare_test.go:216
        assert np.all(r.x >= -1e-9) and np.max(np.abs(A @ r.x - b)) < 1e-8
        checked += 1
    c, A, b = raw_lp(rng, 4, 9, 40)
    for i in range(40):
        r = oracle.simplex(c[i], A[i], b[i])
        h = linprog(c[i], A_eq=A[i], b_eq=b[i], bounds=(0, None), method="highs")
        if r.status == S.GM_OK:
            assert h.status == 0 and abs(r.optF - h.fun) <= 0.005
        elif r.status == S.GM_ERR_INFEASIBLE:
            assert h.status == 2
        elif r.status == S.GM_ERR_UNBOUNDED:
            assert h.status in (3, 2) or h.status == 4
    assert checked == 60


def test_oracle_linear_algebra_against_lapack():
    """The restated Gonum/LAPACK pieces behind every SolveVec / mat.Cond (lu.go:63-84,293-325, matrix.go:284-322):
    solutions agree with LAPACK (numpy), condition estimates bracket the exact 1-norm condition number from below
    (Hager's estimator never overestimates) and flag singular systems the way LU.Solve does."""
    rng = np.random.default_rng(2024)
    for n in (1, 2, 5, 17, 64, 90):
        a = rng.standard_normal((n, n))
        b = rng.standard_normal(n)
        for tr in (False, True):
            rc, x, cond = oracle.solve_vec(a, b, transpose=tr)
            want = np.linalg.solve(a.T if tr else a, b)
            assert rc == 0 and np.allclose(x, want, rtol=1e-9, atol=1e-11)
        exact = np.linalg.cond(a, 1)
        est = oracle.cond1(a)
        assert 0.1 * exact <= est <= exact * (1 + 1e-9)
        # tall slices, as findLinearlyIndependent sees them: cond_1 of the R factor
        if n >= 5:
            k = n // 2
            r = np.linalg.qr(a[:, :k], mode="r")
            exact_r = np.linalg.cond(r, 1)
            est_r = oracle.cond1(a[:, :k])
            assert 0.1 * exact_r <= est_r <= exact_r * (1 + 1e-9)
    sing = np.array([[1.0, 2.0, 3.0], [2.0, 4.0, 6.0], [1.0, 0.0, 1.0]])
    rc, _, _ = oracle.solve_vec(sing, np.ones(3))
    assert rc != 0
    assert oracle.cond1(sing) > 1e15
