#!/usr/bin/env python
"""Generates the oracle fixtures for BASELINE.json's configurations — TEST INFRASTRUCTURE.

    python tests/golden/make_config_fixtures.py [c1] [c3] [c4] [c5] [forced]      (default: all)

Every fixture is produced by running the CPU oracle (oracle/: restatement of Gonum's lp.Simplex and of GoMILP's
branch-and-bound, pinned by tests/test_oracle_golden.py) ONCE here, on inputs drawn from the seeded generators of
tests/problems.py, and is committed under tests/golden/ as a compressed .npz so that the `-m gpu` parity tests
can compare the CUDA path with the oracle at sizes the oracle needs minutes to hours for
(a 700 x 1200 relaxation is ~1000 pivots x 3 fresh LU factorisations of a 700 x 700 basis).

  c1_milps.npz       100 small MILPs, seed 155 (50 getRandomMILP-style dense + 50 bounded), COMPAT and FIXED
                     B&B replays under equal node budgets: status, z, x, node count, full decision log
  c3_knapsack.npz    0-1 knapsack n=500 m=200 seed 7 (standard form 700 x 1200): root + the first 3 waves of
                     children (15 node LPs), per node status / z / x / pivot counts / Bland calls / basis
  c4_large_lp.npz    single dense LP 1024 x 2048 seed 42: optimal basis found by HiGHS, the oracle's
                     (Gonum-arithmetic) z / x on that basis, and the oracle's own first pivots (capped trace)
  c5_general_n*.npz  general-integer MILPs n = 50, 100, 200 (400: root only), bounds as rows, root + children
  forced_paths.npz   small degenerate LPs on which the oracle runs replaceBland, the artificial-still-basic
                     repair loop and rejects columns in the findLinearlyIndependent scan, with pivot traces
"""
from __future__ import annotations

import hashlib
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from problems import (c5_general_integer, feasible_bounded_lp, knapsack, random_milp, standard_form,  # noqa: E402
                      node_lp, most_infeasible)

WORKERS = int(os.environ.get("FIXTURE_WORKERS", "6"))
TRACE_CAP = 64


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def _solve_node(args):
    c0, A0, b0, bvar, bsign, brhs, max_pivots = args
    c, A, b = node_lp(c0, A0, b0, bvar, bsign, brhs)
    t0 = time.time()
    o = oracle.simplex(c, A, b, trace_cap=TRACE_CAP, max_pivots=max_pivots)
    return o, time.time() - t0


def wave_fixture(name, c0, A0, b0, integ, depth, max_nodes, extra=None, max_pivots=0):
    """Root + `depth` waves of children (most-infeasible branching on the oracle's own x, no pruning)."""
    m0, n0 = A0.shape
    waves = [[(np.zeros(0, np.int32), np.zeros(0), np.zeros(0))]]
    recs = []
    with Pool(WORKERS) as pool:
        for L in range(depth + 1):
            nodes = waves[L]
            if not nodes or len(recs) + len(nodes) > max_nodes:
                break
            res = pool.map(_solve_node, [(c0, A0, b0, bv, bs, br, max_pivots) for (bv, bs, br) in nodes])
            nxt = []
            for (bv, bs, br), (o, dt) in zip(nodes, res):
                recs.append((L, bv, bs, br, o))
                print(f"  {name} depth {L}: status {o.status} z {o.optF!r} pivots {o.pivots_phase1}+{o.pivots_phase2} "
                      f"bland {o.bland_calls} repair {o.repair_trials} ({dt:.1f}s)", flush=True)
                if o.status != 0 or o.x is None:
                    continue
                j = most_infeasible(o.x[:n0], integ)
                if j < 0:
                    continue
                fl = np.floor(o.x[j])
                for sg, rh in ((1.0, fl), (-1.0, -(fl + 1.0))):
                    nxt.append((np.append(bv, np.int32(j)).astype(np.int32), np.append(bs, sg), np.append(br, rh)))
            waves.append(nxt)
    N = len(recs)
    Lmax = max(r[0] for r in recs)
    out = dict(L=np.array([r[0] for r in recs], np.int32),
               bvar=np.full((N, max(Lmax, 1)), -1, np.int32), bsign=np.zeros((N, max(Lmax, 1))),
               brhs=np.zeros((N, max(Lmax, 1))),
               status=np.array([r[4].status for r in recs], np.int32),
               z=np.array([r[4].optF for r in recs]),
               x=np.zeros((N, n0)), has_x=np.zeros(N, np.uint8),
               piv1=np.array([r[4].pivots_phase1 for r in recs], np.int64),
               piv2=np.array([r[4].pivots_phase2 for r in recs], np.int64),
               bland=np.array([r[4].bland_calls for r in recs], np.int64),
               repair=np.array([r[4].repair_trials for r in recs], np.int64),
               used_p1=np.array([r[4].used_phase1 for r in recs], np.int32),
               basis=np.full((N, m0 + max(Lmax, 1)), -1, np.int64),
               trace=np.full((N, TRACE_CAP, 4), -1, np.int32), trace_len=np.zeros(N, np.int32),
               m0=np.int64(m0), n0=np.int64(n0), A0_sha=np.array(sha(A0)), integ=np.asarray(integ, np.uint8))
    for k, (L, bv, bs, br, o) in enumerate(recs):
        out["bvar"][k, :L] = bv
        out["bsign"][k, :L] = bs
        out["brhs"][k, :L] = br
        if o.x is not None:
            out["x"][k] = o.x[:n0]
            out["has_x"][k] = 1
        if o.basis is not None:
            out["basis"][k, : m0 + L] = o.basis
        if o.trace is not None:
            out["trace"][k, : len(o.trace)] = o.trace
            out["trace_len"][k] = len(o.trace)
    if extra:
        out.update(extra)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"wrote {name}.npz: {N} node LPs", flush=True)


def make_c3():
    p = knapsack(np.random.default_rng(7), 500, 200)
    c0, A0, b0 = standard_form(p)
    wave_fixture("c3_knapsack", c0, A0, b0, np.concatenate([p["integrality"], np.zeros(A0.shape[1] - 500, np.uint8)]),
                 depth=3, max_nodes=15)


def make_c5():
    for n, depth, max_nodes in ((50, 3, 15), (100, 3, 15), (200, 2, 7), (400, 0, 1)):
        p = c5_general_integer(n)
        c0, A0, b0 = standard_form(p)
        integ = np.concatenate([p["integrality"], np.zeros(A0.shape[1] - n, np.uint8)])
        extra = {}
        if n <= 100:  # full B&B replays under an equal node budget (decision logs), FIXED most-infeasible
            o = oracle.bnb_solve(p["c"], None, None, p["G"], p["h"], p["integrality"], heuristic=1, mode=1,
                                 node_limit=64 if n == 50 else 32)
            extra = dict(bnb_status=np.int32(o.status), bnb_z=np.float64(o.z), bnb_nodes=np.int64(o.nodes),
                         bnb_log_id=o.log["id"], bnb_log_parent=o.log["parent"], bnb_log_lp_status=o.log["lp_status"],
                         bnb_log_z=o.log["z"], bnb_log_decision=o.log["decision"],
                         bnb_log_branch_var=o.log["branch_var"], bnb_log_branch_floor=o.log["branch_floor"],
                         bnb_x=(o.x if o.x is not None else np.zeros(0)))
            print(f"  c5 n={n}: oracle B&B status {o.status} nodes {o.nodes} z {o.z!r}", flush=True)
        wave_fixture(f"c5_general_n{n}", c0, A0, b0, integ, depth=depth, max_nodes=max_nodes, extra=extra)


def make_c4():
    from scipy.optimize import linprog
    m, n = 1024, 2048
    c, A, b = feasible_bounded_lp(np.random.default_rng(42), m, n)
    t0 = time.time()
    hs = linprog(c, A_eq=A, b_eq=b, bounds=(0, None), method="highs-ds")
    assert hs.status == 0
    # optimal basis: the m largest entries of a vertex solution (non-degenerate for continuous random data)
    basis = np.sort(np.argsort(-hs.x)[:m]).astype(np.int64)
    assert hs.x[basis].min() > 1e-9 and np.delete(hs.x, basis).max() < 1e-9
    print(f"  c4: HiGHS optimum {hs.fun!r} in {time.time() - t0:.1f}s", flush=True)
    # the oracle's own arithmetic on that basis: initialBasic = optimal basis -> zero pivots, fresh-LU x and z
    t0 = time.time()
    o = oracle.simplex(c, A, b, initial_basic=basis)
    assert o.status == 0 and o.pivots == 0, (o.status, o.pivots)
    print(f"  c4: oracle on the optimal basis z {o.optF!r} ({time.time() - t0:.1f}s)", flush=True)
    # the oracle's own cold start: the basis scan and its first pivots (capped: a full solve is ~hours of LU)
    t0 = time.time()
    cap = int(os.environ.get("C4_TRACE_PIVOTS", "24"))
    oc = oracle.simplex(c, A, b, trace_cap=cap, max_pivots=cap)
    print(f"  c4: oracle cold start, {cap} pivots: status {oc.status} used_p1 {oc.used_phase1} "
          f"({time.time() - t0:.1f}s)", flush=True)
    np.savez_compressed(os.path.join(HERE, "c4_large_lp.npz"), m=np.int64(m), n=np.int64(n), seed=np.int64(42),
                        A_sha=np.array(sha(A)), highs_z=np.float64(hs.fun), opt_basis=basis, z=np.float64(o.optF),
                        x=o.x, cold_status=np.int32(oc.status), cold_used_p1=np.int32(oc.used_phase1),
                        cold_trace=oc.trace if oc.trace is not None else np.zeros((0, 4), np.int32),
                        cold_cap=np.int64(cap))
    print("wrote c4_large_lp.npz", flush=True)


def make_c1():
    rng = np.random.default_rng(155)
    cases = []
    for k in range(100):
        n = int(rng.integers(2, 12))
        m = int(rng.integers(1, n))
        p = random_milp(rng, n, m, bounded=(k % 2 == 1))
        cases.append(p)
    budget = 400
    out = dict(budget=np.int64(budget), count=np.int64(len(cases)))
    for k, p in enumerate(cases):
        out[f"p{k}_c"] = p["c"]
        out[f"p{k}_G"] = p["G"]
        out[f"p{k}_h"] = p["h"]
        out[f"p{k}_integ"] = np.asarray(p["integrality"], np.uint8)
        if p["A"] is not None:
            out[f"p{k}_A"] = p["A"]
            out[f"p{k}_b"] = p["b"]
        for mode in (0, 1):
            o = oracle.bnb_solve(p["c"], p["A"], p["b"], p["G"], p["h"], p["integrality"], heuristic=1, mode=mode,
                                 node_limit=budget)
            pre = f"p{k}_m{mode}_"
            out[pre + "status"] = np.int32(o.status)
            out[pre + "lp_status"] = np.int32(o.lp_status)
            out[pre + "z"] = np.float64(o.z)
            out[pre + "x"] = o.x if o.x is not None else np.zeros(0)
            out[pre + "nodes"] = np.int64(o.nodes)
            for key in ("id", "parent", "lp_status", "z", "decision", "branch_var", "branch_floor"):
                out[pre + "log_" + key] = o.log[key]
    np.savez_compressed(os.path.join(HERE, "c1_milps.npz"), **out)
    st = [int(out[f"p{k}_m1_status"]) for k in range(len(cases))]
    print("wrote c1_milps.npz; FIXED-mode statuses:", {s: st.count(s) for s in sorted(set(st))}, flush=True)


def make_forced():
    """Search seeded small degenerate LPs for instances where the oracle takes the rarely-used paths."""
    rng = np.random.default_rng(2024)
    found = {"bland": [], "repair": [], "scan": []}
    want = 6
    tries = 0
    while min(len(v) for v in found.values()) < want and tries < 20000:
        tries += 1
        kind = tries % 3
        if kind == 0:      # knapsack children: integer data, active bound rows -> degenerate vertices
            n, m = int(rng.integers(6, 14)), int(rng.integers(2, 5))
            p = knapsack(rng, n, m)
            c0, A0, b0 = standard_form(p)
            L = int(rng.integers(1, 4))
            bv = rng.integers(0, n, size=L).astype(np.int32)
            bs = rng.choice([-1.0, 1.0], size=L)
            br = np.where(bs > 0, 0.0, -1.0)
            c, A, b = node_lp(c0, A0, b0, bv, bs, br)
        elif kind == 1:    # small-integer equality systems: dependent columns in the basis scan, degenerate Phase I
            m, n = int(rng.integers(3, 8)), int(rng.integers(8, 16))
            A = rng.integers(-2, 3, size=(m, n)).astype(float)
            x0 = rng.integers(0, 3, size=n).astype(float) * (rng.random(n) < 0.4)
            b = A @ x0
            c = rng.integers(-3, 4, size=n).astype(float)
        else:              # duplicated / scaled columns at the end: the reverse scan must reject some
            m, n = int(rng.integers(3, 7)), int(rng.integers(8, 14))
            A = rng.standard_normal((m, n))
            k = int(rng.integers(1, 3))
            for q in range(k):
                A[:, n - 1 - q] = A[:, n - 2 - k] * (q + 2.0)
            x0 = rng.random(n) * (rng.random(n) < 0.5)
            b = A @ x0
            c = rng.random(n)
        o = oracle.simplex(c, A, b, trace_cap=256, max_pivots=2000)
        if o.status not in (0, 1, 2):
            continue
        rec = dict(c=c, A=A, b=b, status=o.status, z=o.optF, x=o.x if o.x is not None else np.zeros(A.shape[1]),
                   has_x=o.x is not None, piv1=o.pivots_phase1, piv2=o.pivots_phase2, bland=o.bland_calls,
                   repair=o.repair_trials, trace=o.trace, basis=o.basis if o.basis is not None else np.zeros(0, np.int64))
        # scan rejection: the accepted basis of a cold start is not simply the last m columns
        m = A.shape[0]
        scan_rejects = False
        if o.status == 0:
            st = oracle.initial_basis(A)
            scan_rejects = st is not None and list(st) != list(range(A.shape[1] - 1, A.shape[1] - 1 - m, -1))
        if o.bland_calls > 0 and len(found["bland"]) < want and o.status == 0:
            found["bland"].append(rec)
        elif o.repair_trials > 0 and len(found["repair"]) < want and o.status == 0:
            found["repair"].append(rec)
        elif scan_rejects and len(found["scan"]) < want:
            found["scan"].append(rec)
    out = {}
    for kind, lst in found.items():
        out[kind + "_count"] = np.int64(len(lst))
        for k, r in enumerate(lst):
            for key, v in r.items():
                out[f"{kind}{k}_{key}"] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "forced_paths.npz"), **out)
    print("wrote forced_paths.npz:", {k: len(v) for k, v in found.items()}, f"after {tries} candidates", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "forced", "c5", "c4", "c3"]
    oracle.build()
    for w in which:
        t0 = time.time()
        {"c1": make_c1, "c3": make_c3, "c4": make_c4, "c5": make_c5, "forced": make_forced}[w]()
        print(f"{w}: {time.time() - t0:.1f}s", flush=True)
