"""CPU: the kernel SOURCE that nvcc compiles for sm_100a (gomilp_b200/csrc/simplex_cta.cuh) and the
product's B&B host (csrc/bnb_host.cpp), executed by the fiber CTA emulator (tests/emu) and compared with
the oracle. This is a control-flow check of the device code where there is no GPU; the parity tests proper
are the -m gpu ones, which go through the C ABI of the built library."""
import numpy as np
import pytest

import emu_harness as E
import oracle
from gomilp_b200 import status as S
from problems import feasible_bounded_lp, random_milp, raw_lp, reference_pins, standard_form

PINS = reference_pins()
MILP_STATUS = {"OK": S.GM_MILP_OK, "DEADLINE": S.GM_MILP_DEADLINE_EXCEEDED,
               "NO_INTEGER_FEASIBLE_SOLUTION": S.GM_MILP_NO_INTEGER_FEASIBLE_SOLUTION}


def _close(a, b, tol=1e-9):
    return np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(1.0, np.abs(np.asarray(b)))) <= tol


@pytest.mark.parametrize("T,reg", [(32, False), (64, False), (256, True)])
def test_emulated_kernel_matches_oracle_on_feasible_lps(T, reg):
    rng = np.random.default_rng(11 + T)
    for (m, n, k) in [(1, 3, 4), (3, 7, 6), (6, 13, 6), (12, 25, 4), (20, 40, 2)]:
        c, A, b = feasible_bounded_lp(rng, m, n, k)
        g = E.simplex_batch(c, A, b, T=T, shuffle_order=True, reg=reg)
        o = oracle.simplex_batch(c, A, b)
        assert (g["status"] == o["status"]).all() and (o["status"] == S.GM_OK).all()
        assert _close(g["optF"], o["optF"]) and _close(g["x"], o["x"])
        assert np.max(np.abs(np.einsum("kij,kj->ki", A, g["x"]) - b)) < 1e-9


def test_emulated_quad_mapped_tier_matches_oracle():
    """Tiers 2-3 main loop (basis inverse in shared memory, quad per row, REDUX first-minima)."""
    rng = np.random.default_rng(78)
    for (m, n, k, T) in [(3, 7, 6, 64), (17, 40, 4, 64), (70, 140, 1, 128)]:
        c, A, b = feasible_bounded_lp(rng, m, n, k)
        g = E.simplex_batch(c, A, b, T=T, quad=True, shuffle_order=True)
        o = oracle.simplex_batch(c, A, b)
        assert (g["status"] == o["status"]).all() and _close(g["optF"], o["optF"]) and _close(g["x"], o["x"])
    c, A, b = raw_lp(rng, 6, 11, 24, 0.3)
    g = E.simplex_batch(c, A, b, T=64, quad=True)
    o = oracle.simplex_batch(c, A, b, max_pivots=5000)
    assert (g["status"] == o["status"]).sum() >= 23


def test_emulated_tma_streaming_tier_matches_oracle():
    """The HBM tier's main loop (TMA bulk-copy ring emulated as memcpy + byte-counting mbarrier)."""
    rng = np.random.default_rng(77)
    for (m, n, k) in [(4, 9, 6), (13, 30, 4), (24, 50, 2)]:
        c, A, b = feasible_bounded_lp(rng, m, n, k)
        g = E.simplex_batch(c, A, b, T=64, ring_stages=3, ring_stage_bytes=1024)
        o = oracle.simplex_batch(c, A, b)
        assert (g["status"] == o["status"]).all() and _close(g["optF"], o["optF"]) and _close(g["x"], o["x"])
    c, A, b = raw_lp(rng, 6, 11, 24, 0.3)
    g = E.simplex_batch(c, A, b, T=64, ring_stages=2, ring_stage_bytes=1024)
    o = oracle.simplex_batch(c, A, b, max_pivots=5000)
    assert (g["status"] == o["status"]).sum() >= 23


@pytest.mark.parametrize("reg", [False, True])
def test_emulated_kernel_status_parity_on_raw_lps(reg):
    rng = np.random.default_rng(5)
    agree = total = 0
    for trial in range(12):
        m = int(rng.integers(1, 9))
        n = int(rng.integers(m, 13))
        c, A, b = raw_lp(rng, m, n, 12, p_zero=0.4 if trial % 3 == 0 else 0.0)
        g = E.simplex_batch(c, A, b, T=64, reg=reg)
        o = oracle.simplex_batch(c, A, b, max_pivots=10000)
        for i in range(12):
            total += 1
            agree += int(g["status"][i] == o["status"][i])
            if g["status"][i] == o["status"][i] == S.GM_OK:
                assert _close(g["optF"][i], o["optF"][i]) and _close(g["x"][i], o["x"][i])
    # continuous random data: statuses can only differ through noise-level ties
    assert agree >= total - 1


@pytest.mark.parametrize("reg", [False, True])
def test_emulated_kernel_square_tall_and_initial_basic(reg):
    p = PINS["singular_16x14"]
    g = E.simplex_batch(np.array([p["c"]], float), np.array([p["A"]], float), np.array([p["b"]], float), T=32, reg=reg)
    assert g["status"][0] == S.GM_ERR_SINGULAR
    A = np.array([[[2.0, 1.0], [1.0, 3.0]]])
    g = E.simplex_batch(np.array([[1.0, 1.0]]), A, np.array([[3.0, 4.0]]), T=32, reg=reg)   # m == n, simplex.go:103-119
    assert g["status"][0] == S.GM_OK and _close(g["x"][0], [1.0, 1.0]) and _close(g["optF"][0], 2.0)
    g = E.simplex_batch(np.array([[1.0, 1.0]]), A, np.array([[-3.0, 4.0]]), T=32, reg=reg)
    assert g["status"][0] == S.GM_ERR_INFEASIBLE
    # warm start from the optimal basis: zero pivots (simplex.go:147-160)
    rng = np.random.default_rng(3)
    c, A, b = feasible_bounded_lp(rng, 5, 11, 1)
    o = oracle.simplex(c[0], A[0], b[0])
    g = E.simplex_batch(c, A, b, initial_basic=o.basis[None, :], T=32, reg=reg)
    assert g["status"][0] == S.GM_OK and g["stats"][0, 0] + g["stats"][0, 1] == 0 and _close(g["x"][0], o.x)
    bad = np.array([[0, 0, 1, 2, 3]])  # repeated column: singular -> the reference panics
    assert E.simplex_batch(c, A, b, initial_basic=bad, T=32, reg=reg)["status"][0] == S.GM_PANIC_INITIAL_BASIC


@pytest.mark.parametrize("reg", [False, True])
def test_emulated_wave_matches_materialised_children(reg):
    """Branch rows synthesised on the fly == convertToEqualities materialised (subproblem.go:81-139)."""
    rng = np.random.default_rng(9)
    p = random_milp(rng, 5, 3)
    c0, A0, b0 = standard_form(p)
    m0, n0 = A0.shape
    L, nodes = 2, 6
    bvar = rng.integers(0, 5, size=(nodes, L)).astype(np.int32)
    bsign = rng.choice([-1.0, 1.0], size=(nodes, L))
    brhs = np.where(bsign > 0, rng.integers(0, 4, size=(nodes, L)), -rng.integers(1, 3, size=(nodes, L))).astype(float)
    g = E.simplex_batch(c0, A0, b0, bvar=bvar, bsign=bsign, brhs=brhs, shared_root=True, T=64, reg=reg)
    for k in range(nodes):
        A = np.zeros((m0 + L, n0 + L))
        A[:m0, :n0] = A0
        for l in range(L):
            A[m0 + l, bvar[k, l]] = bsign[k, l]
            A[m0 + l, n0 + l] = 1.0
        o = oracle.simplex(np.concatenate([c0, np.zeros(L)]), A, np.concatenate([b0, brhs[k]]))
        assert g["status"][k] == o.status
        if o.status == S.GM_OK:
            assert _close(g["optF"][k], o.optF) and _close(g["x"][k], o.x[:n0])


@pytest.mark.parametrize("reg", [False, True])
@pytest.mark.parametrize("case", PINS["milp"], ids=[c["src"].split(" ")[0] for c in PINS["milp"]])
def test_emulated_bnb_reproduces_reference_pins(case, reg):
    r = E.milp_solve(case["c"], case["A"], case["b"], case["G"], case["h"], case["integrality"],
                     node_limit=100 if reg else 400, T=32, reg=reg)
    assert r["rc"] == 0 and r["status"] == MILP_STATUS[case["want_status"]]
    if case["want_status"] == "OK":
        assert _close(r["x"], case["want_x"], 1e-12) and abs(r["z"] - case["want_z"]) <= 1e-12


def test_emulated_bnb_replays_oracle_decisions():
    """Same decisions / node counts as the oracle's FIFO replay, except where the reference's exact
    `x == trunc(x)` integrality test (tree.go:290-297) is decided by the last bit of a fresh-LU solve
    (e.g. 3.9999999999999996 vs 4): those trees legitimately differ ("Parity")."""
    rng = np.random.default_rng(21)
    same = 0
    for trial in range(6):
        p = random_milp(rng, int(rng.integers(3, 6)), 2)
        o = oracle.bnb_solve(p["c"], p["A"], p["b"], p["G"], p["h"], p["integrality"], mode=1, node_limit=60)
        g = E.milp_solve(p["c"], p["A"], p["b"], p["G"], p["h"], p["integrality"], mode=1, node_limit=60, T=32)
        if [(l[6], l[7]) for l in g["log"]] == list(zip(o.log["branch_var"].tolist(), o.log["branch_floor"].tolist())):
            same += 1  # same branching sequence => everything else must replay
            assert g["status"] == o.status and g["nodes"] == o.nodes
            assert [l[5] for l in g["log"]] == o.log["decision"].tolist()
        if g["status"] == S.GM_MILP_OK and o.status == S.GM_MILP_OK:
            assert abs(g["z"] - o.z) <= 1e-9 * max(1.0, abs(o.z))
    assert same >= 3


@pytest.mark.parametrize("reg", [False, True])
def test_emulated_warm_started_children_reach_the_cold_optimum(reg):
    """GM_BNB_WARM_START: children continue from [B 0; g 1]^-1 built from the parent's inverse. Same optimum,
    same tree where the LP optima are unique, fewer pivots."""
    rng = np.random.default_rng(33)
    fewer = 0
    for _ in range(4):
        p = random_milp(rng, int(rng.integers(3, 7)), 3)
        cold = E.milp_solve(p["c"], None, None, p["G"], p["h"], p["integrality"], mode=1, heuristic=1, node_limit=60,
                            T=64, reg=reg)
        warm = E.milp_solve(p["c"], None, None, p["G"], p["h"], p["integrality"], mode=1 | S.GM_BNB_WARM_START,
                            heuristic=1, node_limit=60, T=64, reg=reg)
        assert warm["status"] == cold["status"]
        if cold["status"] == S.GM_MILP_OK:
            assert abs(warm["z"] - cold["z"]) <= 1e-9 * max(1.0, abs(cold["z"]))
            assert _close(warm["x"], cold["x"])
        if warm["nodes"] == cold["nodes"] and cold["nodes"] > 1:
            fewer += int(warm["pivots"] < cold["pivots"])
    assert fewer >= 1
