"""GPU (-m gpu): the CUDA path, called through the C ABI of libgomilp_b200.so, against the CPU oracle on the
same seeded inputs and against the reference's golden vectors; at BASELINE.json's full C2 size through
size-independent properties (primal feasibility, x >= 0, dual feasibility / optimality certificate).
Tolerance: 1e-9 relative on objectives and primal values (BASELINE.json north_star)."""
import numpy as np
import pytest

import gomilp_b200 as gm
import oracle
from gomilp_b200 import status as S
from problems import (feasible_bounded_lp, knapsack, random_milp, raw_lp, reference_pins, standard_form)

pytestmark = pytest.mark.gpu
PINS = reference_pins()
RTOL = 1e-9
MILP_STATUS = {"OK": S.GM_MILP_OK, "DEADLINE": S.GM_MILP_DEADLINE_EXCEEDED,
               "NO_INTEGER_FEASIBLE_SOLUTION": S.GM_MILP_NO_INTEGER_FEASIBLE_SOLUTION}


def _close(a, b, tol=RTOL):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))) <= tol


def test_native_library_is_the_one_loaded():
    assert gm.device_count() >= 1
    gm.init(0)
    with open("/proc/self/maps") as f:
        assert "libgomilp_b200.so" in f.read()


@pytest.mark.parametrize("case", PINS["milp"], ids=[c["src"].split(" ")[0] for c in PINS["milp"]])
def test_reference_milp_pins(case):
    r = gm.milp_solve(case["c"], case["A"], case["b"], case["G"], case["h"], case["integrality"], node_limit=400)
    assert r.status == MILP_STATUS[case["want_status"]]
    if case["want_status"] == "OK":
        assert _close(r.x, case["want_x"], 1e-12) and abs(r.z - case["want_z"]) <= 1e-12


def test_reference_status_pins():
    p = PINS["singular_16x14"]
    assert gm.simplex(p["c"], p["A"], p["b"]).status == S.GM_ERR_SINGULAR
    p = PINS["api_end_to_end"]
    r = gm.milp_solve(p["c"], p["A"], p["b"], p["G"], p["h"], p["integrality"])
    assert r.status == S.GM_MILP_OK and _close(r.x, p["want_x"], 1e-12)


def test_single_lp_entry_point_and_status_taxonomy():
    r = gm.simplex([-1, -2, 0, 0], [[-1, 2, 1, 0], [3, 1, 0, 1]], [4, 9])
    assert r.status == S.GM_OK and _close(r.optF, -8.0, 1e-12) and _close(r.x, [2, 3, 0, 0], 1e-12) and r.pivots == 2
    A = np.array([[1.0, 2.0, 3.0], [0.0, 0.0, 0.0]])
    assert gm.simplex([1, 1, 1], A, [1, 1]).status == S.GM_ERR_INFEASIBLE
    assert gm.simplex([1, 1, 1], A, [1, 0]).status == S.GM_ERR_ZERO_ROW
    A = np.array([[1.0, 0.0, 3.0], [2.0, 0.0, 1.0]])
    assert gm.simplex([1, -1, 1], A, [1, 1]).status == S.GM_ERR_UNBOUNDED
    assert gm.simplex([1, 1, 1], A, [1, 1]).status == S.GM_ERR_ZERO_COLUMN
    r = gm.simplex([-1.0, 0.0], [[1.0, -1.0]], [1.0])
    assert r.status == S.GM_ERR_UNBOUNDED and r.optF == -np.inf and r.x is None
    r = gm.simplex([1.0, 1.0], [[2.0, 1.0], [1.0, 3.0]], [3.0, 4.0])  # m == n
    assert r.status == S.GM_OK and _close(r.x, [1, 1]) and _close(r.optF, 2.0)
    rng = np.random.default_rng(3)
    c, A, b = feasible_bounded_lp(rng, 9, 20)
    o = oracle.simplex(c, A, b)
    w = gm.simplex(c, A, b, initial_basic=o.basis)  # warm start from the optimal basis: no pivots
    assert w.status == S.GM_OK and w.pivots == 0 and _close(w.x, o.x) and _close(w.optF, o.optF)
    assert gm.simplex(c, A, b, initial_basic=[0] * 9).status == S.GM_PANIC_INITIAL_BASIC


@pytest.mark.parametrize("m,n,count", [(1, 2, 8), (2, 5, 64), (7, 19, 128), (16, 32, 256), (33, 70, 64),
                                        (64, 128, 192), (64, 65, 16), (100, 180, 24)])
def test_batch_matches_oracle(m, n, count):
    rng = np.random.default_rng(1000 * m + n)
    c, A, b = feasible_bounded_lp(rng, m, n, count)
    g = gm.simplex_batch(c, A, b)
    o = oracle.simplex_batch(c, A, b, threads=oracle.num_hw_threads())
    assert (g["status"] == o["status"]).all() and (o["status"] == S.GM_OK).all()
    assert _close(g["optF"], o["optF"]) and _close(g["x"], o["x"])
    same_basis = np.mean([set(g["basis"][i]) == set(o["basis"][i]) for i in range(count)])
    same_pivots = np.mean(g["pivots"] == o["pivots"])
    print(f"m={m} n={n}: identical final basis {same_basis:.3f}, identical pivot count {same_pivots:.3f}")
    if n >= 2 * m:
        assert same_basis >= 0.9


def test_status_parity_on_raw_gaussian_lps():
    rng = np.random.default_rng(77)
    agree = total = 0
    for (m, n, count, pz) in [(3, 7, 256, 0.0), (6, 10, 256, 0.3), (10, 24, 128, 0.0), (5, 5, 64, 0.0), (6, 4, 16, 0.0)]:
        c, A, b = raw_lp(rng, m, n, count, pz)
        g = gm.simplex_batch(c, A, b)
        o = oracle.simplex_batch(c, A, b, threads=oracle.num_hw_threads(), max_pivots=20000)
        total += count
        agree += int((g["status"] == o["status"]).sum())
        ok = (g["status"] == S.GM_OK) & (o["status"] == S.GM_OK)
        if ok.any():
            assert _close(g["optF"][ok], o["optF"][ok]) and _close(g["x"][ok], o["x"][ok])
    print(f"status agreement {agree}/{total}")
    assert agree >= total - 3  # continuous data: only noise-level ties (r_e = -1e-17 at tol 0) can differ


def test_every_tier_matches_oracle():
    """Tier 1 keeps the basis inverse in registers (m <= 64); tier 2 keeps it in shared memory; tiers 3 / 4
    run the same solver with W / the inverse (and the vectors) in an HBM workspace."""
    rng = np.random.default_rng(5)
    c, A, b = feasible_bounded_lp(rng, 150, 260, 6)
    g = gm.simplex_batch(c, A, b)
    assert gm.last_timing()["tier"] == 6   # 6 LPs on 148 SMs: the cooperative tier, 24 CTAs per LP
    o = oracle.simplex_batch(c, A, b, threads=oracle.num_hw_threads())
    assert (g["status"] == o["status"]).all()
    assert _close(g["optF"], o["optF"]) and _close(g["x"], o["x"])
    try:
        gm.set_options(force_tier=3)                     # basis inverse in shared memory, W in HBM, one CTA per LP
        g3 = gm.simplex_batch(c, A, b)
        assert gm.last_timing()["tier"] == 3
        assert (g3["status"] == o["status"]).all() and _close(g3["x"], o["x"]) and (g3["pivots"] == g["pivots"]).all()
        gm.set_options(force_tier=4)                     # W and Bi in HBM (m < 384: plain loads)
        g2 = gm.simplex_batch(c, A, b)
        assert (g2["status"] == o["status"]).all() and _close(g2["x"], o["x"])
    finally:
        gm.set_options()
    # long rows: tier 4 streams W / Bi through the TMA staging ring; same answers with the ring disabled, and the
    # objective HiGHS finds. Slack form [R I] x = b, b > 0: non-degenerate, the slack basis is feasible.
    from scipy.optimize import linprog
    m4, n4 = 400, 700
    A4 = np.zeros((2, m4, n4))
    A4[:, :, : n4 - m4] = rng.random((2, m4, n4 - m4))
    A4[:, :, n4 - m4:] = np.eye(m4)
    b4 = 1.0 + rng.random((2, m4))
    c4 = np.zeros((2, n4))
    c4[:, : n4 - m4] = -rng.random((2, n4 - m4))
    try:
        gm.set_options(force_tier=4)
        g4 = gm.simplex_batch(c4, A4, b4)
    finally:
        gm.set_options()
    assert gm.last_timing()["tier"] == 4 and (g4["status"] == S.GM_OK).all()
    assert np.abs(np.einsum("kij,kj->ki", A4, g4["x"]) - b4).max() < 1e-9 and g4["x"].min() >= -1e-12
    for k in range(2):
        hs = linprog(c4[k], A_eq=A4[k], b_eq=b4[k], bounds=(0, None), method="highs")
        assert hs.status == 0 and abs(hs.fun - g4["optF"][k]) <= 1e-7 * max(1.0, abs(hs.fun))
    try:
        gm.set_options(no_tma_ring=True, force_tier=4)
        g5 = gm.simplex_batch(c4, A4, b4)
    finally:
        gm.set_options()
    g6 = gm.simplex_batch(c4, A4, b4)                    # the shape's own choice: cooperative, 74 CTAs per LP
    assert gm.last_timing()["tier"] == 6 and (g6["status"] == S.GM_OK).all()
    assert _close(g6["optF"], g4["optF"]) and _close(g6["x"], g4["x"], 1e-7) and (g6["pivots"] == g4["pivots"]).all()
    assert (g5["status"] == S.GM_OK).all() and _close(g5["optF"], g4["optF"]) and _close(g5["x"], g4["x"], 1e-7)
    assert (g5["pivots"] == g4["pivots"]).all()
    c, A, b = feasible_bounded_lp(rng, 16, 40, 32)
    c2, A2, b2 = raw_lp(rng, 9, 17, 64, 0.2)
    o = oracle.simplex_batch(c, A, b)
    o2 = oracle.simplex_batch(c2, A2, b2, max_pivots=20000)
    try:
        for tier in (1, 2, 3, 4, 5, 6):
            gm.set_options(force_tier=tier, coop_group=4 if tier == 6 else 0)
            g = gm.simplex_batch(c, A, b)
            assert gm.last_timing()["tier"] == tier
            assert (g["status"] == o["status"]).all() and _close(g["x"], o["x"]) and _close(g["optF"], o["optF"])
            g2 = gm.simplex_batch(c2, A2, b2)
            assert (g2["status"] == o2["status"]).sum() >= 63
    finally:
        gm.set_options()


def test_wave_matches_oracle_children():
    rng = np.random.default_rng(9)
    p = random_milp(rng, 12, 6)
    c0, A0, b0 = standard_form(p)
    m0, n0 = A0.shape
    L, nodes = 3, 64
    bvar = rng.integers(0, 12, size=(nodes, L)).astype(np.int32)
    bsign = rng.choice([-1.0, 1.0], size=(nodes, L))
    brhs = np.where(bsign > 0, rng.integers(0, 4, size=(nodes, L)), -rng.integers(1, 3, size=(nodes, L))).astype(float)
    root = gm.upload_root(c0, A0, b0)
    try:
        w = gm.solve_wave(root, n0, m0, bvar, bsign, brhs)
    finally:
        gm.free_root(root)
    agree = 0
    for k in range(nodes):
        A = np.zeros((m0 + L, n0 + L))
        A[:m0, :n0] = A0
        for l in range(L):
            A[m0 + l, bvar[k, l]] = bsign[k, l]
            A[m0 + l, n0 + l] = 1.0
        o = oracle.simplex(np.concatenate([c0, np.zeros(L)]), A, np.concatenate([b0, brhs[k]]))
        agree += int(w.status[k] == o.status)
        if o.status == S.GM_OK and w.status[k] == S.GM_OK:
            assert _close(w.z[k], o.optF) and _close(w.x[k], o.x[:n0])
    assert agree >= nodes - 1


@pytest.mark.timeout(300)
def test_milp_objective_matches_oracle_and_highs():
    from scipy.optimize import Bounds, LinearConstraint, milp
    rng = np.random.default_rng(155)
    same_tree = solved = 0
    cases = 12
    for _ in range(cases):
        p = random_milp(rng, int(rng.integers(3, 8)), int(rng.integers(1, 4)))
        g = gm.milp_solve(p["c"], p["A"], p["b"], p["G"], p["h"], p["integrality"], mode=S.GM_BNB_FIXED,
                          node_limit=300)
        # the oracle replays the reference's arithmetic, whose last-bit noise can make the tree non-terminating
        # (3.9999999999999996 is "fractional" for tree.go:290-297): keep its budget small
        o = oracle.bnb_solve(p["c"], p["A"], p["b"], p["G"], p["h"], p["integrality"], mode=S.GM_BNB_FIXED,
                             node_limit=120)
        hs = milp(p["c"], constraints=LinearConstraint(p["G"], -np.inf, p["h"]), integrality=p["integrality"],
                  bounds=Bounds(0, np.inf), options={"time_limit": 10})
        assert hs.status == 0
        if g.status == S.GM_MILP_OK:
            solved += 1
            assert abs(g.z - hs.fun) <= 0.005           # the reference's GLPK tolerance
            if o.status == S.GM_MILP_OK:
                assert abs(g.z - o.z) <= RTOL * max(1.0, abs(o.z))
        same_tree += int(o.nodes == g.nodes and o.status == g.status)
    print(f"solved {solved}/{cases}; identical node count as the oracle replay: {same_tree}/{cases}")
    assert solved >= cases - 2


def test_full_size_c2_batch_properties():
    """BASELINE.json configs[1]: 4096 LPs, m=64, n=128. Oracle parity on a slice, optimality certificate on all."""
    rng = np.random.default_rng(1234)
    m, n, count = 64, 128, 4096
    c, A, b = feasible_bounded_lp(rng, m, n, count)
    g = gm.simplex_batch(c, A, b)
    assert (g["status"] == S.GM_OK).all()
    x = g["x"]
    assert x.min() >= -1e-9
    res = np.abs(np.einsum("kij,kj->ki", A, x) - b).max(axis=1)
    assert res.max() <= 1e-8 * max(1.0, np.abs(b).max())
    assert _close(g["optF"], np.einsum("kj,kj->k", c, x))
    assert ((x != 0).sum(axis=1) <= m).all()  # basic solutions
    # dual certificate: y = B^-T c_B gives reduced costs >= -1e-7 on every column
    for k in range(0, count, 97):
        B = A[k][:, g["basis"][k]]
        y = np.linalg.solve(B.T, c[k][g["basis"][k]])
        assert (c[k] - A[k].T @ y).min() >= -1e-7
    sl = slice(0, 128)
    o = oracle.simplex_batch(c[sl], A[sl], b[sl], threads=oracle.num_hw_threads())
    assert _close(g["optF"][sl], o["optF"]) and _close(g["x"][sl], o["x"])


@pytest.mark.timeout(300)
def test_knapsack_wave_root_and_first_levels():
    rng = np.random.default_rng(7)
    p = knapsack(rng, 60, 8)
    g = gm.milp_solve(p["c"], None, None, p["G"], p["h"], p["integrality"], mode=S.GM_BNB_FIXED, node_limit=64)
    o = oracle.bnb_solve(p["c"], None, None, p["G"], p["h"], p["integrality"], mode=S.GM_BNB_FIXED, node_limit=1)
    assert g.log[0][3] == S.GM_OK and abs(g.log[0][4] - o.log["z"][0]) <= RTOL * abs(o.log["z"][0])
    assert g.status in (S.GM_MILP_OK, S.GM_MILP_DEADLINE_EXCEEDED)


@pytest.mark.timeout(300)
def test_sharded_driver_single_rank():
    """The multi-rank scheduler (gomilp_b200/sharded.py) over the C ABI with one rank == gm_milp_solve."""
    from gomilp_b200.sharded import gpu_wave_solver, milp_solve_sharded
    for case in PINS["milp"]:
        r = milp_solve_sharded(case["c"], case["A"], case["b"], case["G"], case["h"], case["integrality"],
                               solve_wave=gpu_wave_solver(), node_limit=200)
        g = gm.milp_solve(case["c"], case["A"], case["b"], case["G"], case["h"], case["integrality"], node_limit=200)
        assert r.status == g.status == MILP_STATUS[case["want_status"]] and r.nodes == g.nodes
        assert [d[5] for d in r.decisions] == [l[5] for l in g.log]
        if case["want_status"] == "OK":
            assert _close(r.x, case["want_x"], 1e-12) and abs(r.z - case["want_z"]) <= 1e-12
    rng = np.random.default_rng(8)
    p = knapsack(rng, 24, 4)
    r = milp_solve_sharded(p["c"], None, None, p["G"], p["h"], p["integrality"], solve_wave=gpu_wave_solver(),
                           mode=S.GM_BNB_FIXED, heuristic=S.GM_BRANCH_MOST_INFEASIBLE, node_limit=500)
    g = gm.milp_solve(p["c"], None, None, p["G"], p["h"], p["integrality"], mode=S.GM_BNB_FIXED,
                      heuristic=S.GM_BRANCH_MOST_INFEASIBLE, node_limit=500)
    assert r.status == g.status and r.nodes == g.nodes and (r.z == g.z or (r.x is None and g.x is None))


@pytest.mark.timeout(300)
def test_warm_started_children_reach_the_cold_optimum():
    rng = np.random.default_rng(41)
    ratio = []
    for _ in range(8):
        p = random_milp(rng, int(rng.integers(4, 10)), int(rng.integers(2, 5)))
        cold = gm.milp_solve(p["c"], None, None, p["G"], p["h"], p["integrality"], mode=S.GM_BNB_FIXED, heuristic=1,
                             node_limit=2000)
        warm = gm.milp_solve(p["c"], None, None, p["G"], p["h"], p["integrality"],
                             mode=S.GM_BNB_FIXED | S.GM_BNB_WARM_START, heuristic=1, node_limit=2000)
        assert warm.status == cold.status
        if cold.status == S.GM_MILP_OK:
            assert abs(warm.z - cold.z) <= RTOL * max(1.0, abs(cold.z))
        if warm.nodes == cold.nodes and cold.nodes > 1:
            ratio.append(cold.pivots / max(1, warm.pivots))
    print("cold / warm pivots on identical trees:", [round(r, 2) for r in ratio])
    assert ratio and min(ratio) >= 1.0


def test_edge_cases_empty_ragged_and_strided():
    """Empty batch, 1x2 and 1x1 LPs, a wave of depth 0 (the root), row stride > n (mat.Dense views), odd shapes
    on every tier boundary."""
    import ctypes as C
    # empty batch: nothing to do, GM_OK
    e = gm.simplex_batch(np.zeros((0, 3)), np.zeros((0, 2, 3)), np.zeros((0, 2)))
    assert e["status"].shape == (0,)
    # 1 x 1 (m == n path) and 1 x 2
    r = gm.simplex([2.0], [[4.0]], [2.0])
    assert r.status == S.GM_OK and _close(r.x, [0.5]) and _close(r.optF, 1.0)
    r = gm.simplex([1.0, -1.0], [[1.0, 1.0]], [3.0])
    assert r.status == S.GM_OK and _close(r.x, [0.0, 3.0]) and _close(r.optF, -3.0)
    # root wave (L = 0) equals the single-LP entry point
    rng = np.random.default_rng(12)
    c, A, b = feasible_bounded_lp(rng, 7, 15)
    root = gm.upload_root(c, A, b)
    try:
        w = gm.solve_wave(root, 15, 7, np.zeros((1, 0), dtype=np.int32), np.zeros((1, 0)), np.zeros((1, 0)))
    finally:
        gm.free_root(root)
    s1 = gm.simplex(c, A, b)
    assert w.status[0] == s1.status == S.GM_OK and _close(w.x[0], s1.x) and _close(w.z[0], s1.optF)
    # strided A (lda > n): a column slice of a wider row-major matrix, passed without copying
    wide = np.zeros((7, 20))
    wide[:, :15] = A
    L = gm.capi.lib()
    optF = C.c_double(0.0)
    x = np.zeros(15)
    st = L.gm_simplex(c.ctypes.data_as(C.c_void_p), wide.ctypes.data_as(C.c_void_p), 20, b.ctypes.data_as(C.c_void_p),
                      7, 15, 0.0, None, C.byref(optF), x.ctypes.data_as(C.c_void_p), None, None)
    assert st == S.GM_OK and _close(x, s1.x) and _close(optF.value, s1.optF)
    # shapes straddling the tier boundaries, checked against the oracle
    for (m, n, k) in [(64, 200, 4), (65, 131, 4), (63, 64, 4), (97, 99, 2)]:
        c, A, b = feasible_bounded_lp(rng, m, n, k)
        g = gm.simplex_batch(c, A, b)
        o = oracle.simplex_batch(c, A, b, threads=oracle.num_hw_threads(), max_pivots=20000)
        ok = o["status"] == S.GM_OK
        assert (g["status"][ok] == S.GM_OK).all()
        assert _close(g["optF"][ok], o["optF"][ok]) and _close(g["x"][ok], o["x"][ok], 1e-8)
    # bad arguments come back as codes, never as exceptions across the C boundary
    assert L.gm_simplex(c.ctypes.data_as(C.c_void_p), wide.ctypes.data_as(C.c_void_p), 3, b.ctypes.data_as(C.c_void_p),
                        7, 15, 0.0, None, C.byref(optF), x.ctypes.data_as(C.c_void_p), None, None) == S.GM_ERR_BAD_SHAPE
    assert L.gm_free_root(123456) == S.GM_ERR_BAD_HANDLE


def test_results_are_bitwise_reproducible_run_to_run():
    """compute-sanitizer is closed on this GPU pool, so shared-memory races cannot be checked directly; a race would
    show as run-to-run differences (the persistent CTAs pick LPs in a different order every launch)."""
    rng = np.random.default_rng(99)
    c, A, b = feasible_bounded_lp(rng, 48, 100, 600)
    c2, A2, b2 = raw_lp(rng, 20, 45, 600, 0.1)
    try:
        for tier in (1, 2, 3, 4):
            gm.set_options(force_tier=tier)
            runs = [gm.simplex_batch(c, A, b) for _ in range(3)]
            runs2 = [gm.simplex_batch(c2, A2, b2) for _ in range(2)]
            for r in runs[1:]:
                assert np.array_equal(r["status"], runs[0]["status"]) and np.array_equal(r["x"], runs[0]["x"])
                assert np.array_equal(r["optF"], runs[0]["optF"]) and np.array_equal(r["basis"], runs[0]["basis"])
                assert np.array_equal(r["pivots"], runs[0]["pivots"])
            assert np.array_equal(runs2[1]["status"], runs2[0]["status"])
            assert np.array_equal(runs2[1]["x"], runs2[0]["x"])
    finally:
        gm.set_options()


def test_streamed_pinned_batch_equals_the_sliced_and_the_plain_path():
    """Large host batches: pinned inputs go through ONE launch whose CTAs wait on the copy stream's arrival counter
    (run_host_batch_streamed); pageable inputs through one launch per slice. Same LPs, same kernel, same results."""
    import torch
    rng = np.random.default_rng(1234)
    m, n, count = 64, 128, 2048
    c, A, b = feasible_bounded_lp(rng, m, n, count)
    plain = gm.simplex_batch(c, A, b)                       # pageable numpy memory: sliced launches
    pc = torch.from_numpy(c).pin_memory().numpy()
    pA = torch.from_numpy(A).pin_memory().numpy()
    pb = torch.from_numpy(b).pin_memory().numpy()
    streamed = gm.simplex_batch(pc, pA, pb)                 # pinned: one gated launch
    tm = gm.last_timing()
    assert tm["launches"] == 1 and tm["lps"] == count
    try:
        gm.set_options(no_streamed_batch=True)
        sliced = gm.simplex_batch(pc, pA, pb)
        assert gm.last_timing()["launches"] > 1
    finally:
        gm.set_options()
    for other in (streamed, sliced):
        assert np.array_equal(other["status"], plain["status"]) and (plain["status"] == S.GM_OK).all()
        assert np.array_equal(other["optF"], plain["optF"]) and np.array_equal(other["x"], plain["x"])
        assert np.array_equal(other["basis"], plain["basis"]) and np.array_equal(other["pivots"], plain["pivots"])
