"""Seeded synthetic problem generators shared by tests/ and bench.py.

Distributions follow the reference's own generators where it has one: getRandomMILP (ilp_test.go:370-429:
A, G, b, h, c ~ N(0,1), integrality Bernoulli(1/2)); the feasible+bounded LP generator and the knapsack are
this repo's (the reference has no benchmark inputs).
"""
from __future__ import annotations

import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_pins() -> dict:
    with open(os.path.join(_HERE, "golden", "reference_pins.json")) as f:
        return json.load(f)


def feasible_bounded_lp(rng: np.random.Generator, m: int, n: int, count: int | None = None):
    """Standard-form LPs with a finite optimum: b = A x0 (x0 >= 0), c = A'y0 + s0 (s0 >= 0)."""
    k = 1 if count is None else count
    A = rng.standard_normal((k, m, n))
    x0 = rng.random((k, n)) * (rng.random((k, n)) < 0.5)
    b = np.einsum("kij,kj->ki", A, x0)
    y0 = rng.standard_normal((k, m))
    s0 = rng.random((k, n))
    c = np.einsum("kij,ki->kj", A, y0) + s0
    if count is None:
        return c[0], A[0], b[0]
    return c, A, b


def raw_lp(rng: np.random.Generator, m: int, n: int, count: int, p_zero: float = 0.0):
    """getRandomMILP-style raw N(0,1) data: mostly infeasible / unbounded, for status parity."""
    A = rng.standard_normal((count, m, n))
    if p_zero > 0:
        A[rng.random(A.shape) < p_zero] = 0.0
    return rng.standard_normal((count, n)), A, rng.standard_normal((count, m))


def random_milp(rng: np.random.Generator, n: int, m: int, bounded: bool = True):
    """Small MILP in GoMILP's numeric form: min c'x, Ax=b?, Gx<=h, x>=0, some x integer.

    bounded=True keeps the relaxation bounded and feasible (positive G, positive h, box rows) so that
    branch-and-bound has something to do; the raw getRandomMILP distribution is almost always infeasible
    or unbounded at the root (which the reference turns into a panic)."""
    if not bounded:
        return dict(c=rng.standard_normal(n), A=rng.standard_normal((m, n)), b=rng.standard_normal(m),
                    G=rng.standard_normal((m, n)), h=rng.standard_normal(m),
                    integrality=(rng.random(n) < 0.5).astype(np.uint8))
    G = np.vstack([rng.random((m, n)) + 0.1, np.eye(n)])
    h = np.concatenate([G[:m].sum(axis=1) * (1.5 + rng.random(m)), np.full(n, 4.0)])
    c = -(rng.random(n) + 0.1)
    integ = (rng.random(n) < 0.6).astype(np.uint8)
    if not integ.any():
        integ[0] = 1
    return dict(c=c, A=None, b=None, G=G, h=h, integrality=integ)


def knapsack(rng: np.random.Generator, n: int, m: int):
    """0-1 multidimensional knapsack (config C3): max p'x s.t. Wx <= cap, x binary, as min -p'x."""
    W = rng.integers(1, 1001, size=(m, n)).astype(np.float64)
    p = W.sum(axis=0) / m + rng.integers(0, 501, size=n)
    cap = np.floor(0.5 * W.sum(axis=1))
    G = np.vstack([W, np.eye(n)])  # capacity rows, then x_j <= 1 rows (api.go:245-272 builds bounds as G rows)
    h = np.concatenate([cap, np.ones(n)])
    return dict(c=-p, A=None, b=None, G=G, h=h, integrality=np.ones(n, dtype=np.uint8))


def standard_form(p: dict):
    """[A 0; G I] of toInitialSubproblem / convertToEqualities (ilp.go:43-71, subproblem.go:81-139)."""
    c = np.asarray(p["c"], dtype=np.float64)
    nvar = c.shape[0]
    A = None if p.get("A") is None else np.asarray(p["A"], dtype=np.float64).reshape(-1, nvar)
    G = None if p.get("G") is None else np.asarray(p["G"], dtype=np.float64).reshape(-1, nvar)
    meq = 0 if A is None else A.shape[0]
    nineq = 0 if G is None else G.shape[0]
    A0 = np.zeros((meq + nineq, nvar + nineq))
    b0 = np.zeros(meq + nineq)
    if meq:
        A0[:meq, :nvar] = A
        b0[:meq] = p["b"]
    if nineq:
        A0[meq:, :nvar] = G
        A0[meq:, nvar:] = np.eye(nineq)
        b0[meq:] = p["h"]
    c0 = np.concatenate([c, np.zeros(nineq)])
    return c0, A0, b0


def c5_general_integer(n: int):
    """Config C5: general-integer MILP, m_ineq = n/2 dense rows G ~ U(0,1), h = G u / 2, bounds
    0 <= x <= u = 10 materialised as rows like api.go:245-272 does, c ~ -U(0,1) (min), seed 100 + n."""
    rng = np.random.default_rng(100 + n)
    mi = n // 2
    Gd = rng.random((mi, n))
    u = 10.0
    G = np.vstack([Gd, np.eye(n)])
    h = np.concatenate([0.5 * Gd @ np.full(n, u), np.full(n, u)])
    return dict(c=-rng.random(n), A=None, b=None, G=G, h=h, integrality=np.ones(n, dtype=np.uint8))


def node_lp(c0, A0, b0, bvar, bsign, brhs):
    """The explicit LP of one B&B node: [A0 0; G I] x = [b0; h] with one row per bnbConstraint
    (combineInequalities + convertToEqualities, subproblem.go:55-139)."""
    m0, n0 = A0.shape
    L = len(bvar)
    A = np.zeros((m0 + L, n0 + L))
    A[:m0, :n0] = A0
    for l in range(L):
        A[m0 + l, int(bvar[l])] = bsign[l]
        A[m0 + l, n0 + l] = 1.0
    return np.concatenate([c0, np.zeros(L)]), A, np.concatenate([b0, np.asarray(brhs, dtype=np.float64)])


def most_infeasible(x, integ) -> int:
    """FIXED-mode most-infeasible branching point (bnb_host.cpp fixed_point): the integer-flagged variable whose
    fractional part is closest to 1/2, first one on ties; -1 when x is integer feasible (exact x == trunc(x))."""
    best, bestv = -1, -1.0
    for i in range(len(x)):
        if not integ[i] or x[i] == np.trunc(x[i]):
            continue
        f = x[i] - np.floor(x[i])
        score = 0.5 - abs(0.5 - f)
        if score > bestv:
            best, bestv = i, score
    return best
