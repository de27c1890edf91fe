"""GPU probe (not a test): timings of the cooperative tier on BASELINE's shapes. One JSON line per case."""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import gomilp_b200 as gm  # noqa: E402
from problems import c5_general_integer, feasible_bounded_lp, knapsack, standard_form  # noqa: E402

gm.init(0)


def bpp(m, n):
    return 8 * (3 * m * m + m * (n - m))


def run(label, c, A, b, **opt):
    gm.set_options(**opt)
    try:
        gm.profile_arm()
        t0 = time.perf_counter()
        g = gm.simplex_batch(c, A, b, want_basis=False)
        wall = time.perf_counter() - t0
        tm = gm.last_timing()
        prof = gm.profile_fetch(1)
        pr = prof[0].astype(float) if len(prof) else None
        piv = int(g["pivots"].sum())
        m, n = A.shape[1:]
        s = g["stats"] if "stats" in g else None
        print(json.dumps({"case": label, "count": int(A.shape[0]), "m": m, "n": n, "tier": tm["tier"], "grid": tm["grid"],
                          "status": sorted(set(int(v) for v in g["status"])), "pivots": piv, "kernel_ms": tm["kernel_ms"],
                          "us_per_pivot_round": 1e3 * tm["kernel_ms"] / max(1, piv / A.shape[0]),
                          "alg_GBps": piv * bpp(m, n) / (tm["kernel_ms"] * 1e-3) / 1e9, "wall_s": wall,
                          "inversions": None if s is None else int(s[:, 3].sum()),
                          "bland": None if s is None else int(s[:, 2].sum()),
                          "leader_cycles_pct": None if pr is None or pr[0] == 0 else {
                              "main_loop": round(100 * pr[1] / pr[0], 1), "inversion": round(100 * pr[2] / pr[0], 1),
                              "polish": round(100 * pr[3] / pr[0], 1), "leader_bland": round(100 * pr[4] / pr[0], 1),
                              "refactor": round(100 * pr[5] / pr[0], 1), "entries": int(pr[6]), "polishes": int(pr[7]),
                              "solve_Mcyc": round(pr[0] / 1e6, 1)}}), flush=True)
        return g
    finally:
        gm.set_options()


rng = np.random.default_rng(1)
c, A, b = feasible_bounded_lp(rng, 40, 90, 4)
run("small forced tier 6", c, A, b, force_tier=6, coop_group=4)
c, A, b = feasible_bounded_lp(rng, 150, 300, 1)
run("150x300 x1 auto", c, A, b)
run("150x300 x1 tier 3", c, A, b, force_tier=3)
c, A, b = feasible_bounded_lp(rng, 150, 300, 16)
run("150x300 x16 auto", c, A, b)
c, A, b = feasible_bounded_lp(rng, 400, 800, 1)
run("400x800 x1 auto", c, A, b)
lps = [feasible_bounded_lp(np.random.default_rng(42 + k), 1024, 2048) for k in range(16)]
c = np.stack([l[0] for l in lps]); A = np.stack([l[1] for l in lps]); b = np.stack([l[2] for l in lps])
run("C4 single", c[:1], A[:1], b[:1])
run("C4 single G=64", c[:1], A[:1], b[:1], coop_group=64)
run("C4 single G=32", c[:1], A[:1], b[:1], coop_group=32)
run("C4 batch16", c, A, b)
run("C4 batch4", c[:4], A[:4], b[:4])
# C3 root + one child through the wave API
p = knapsack(np.random.default_rng(7), 500, 200)
c0, A0, b0 = standard_form(p)
root = gm.upload_root(c0, A0, b0)
for L, bv, bs, br in ((0, np.zeros((1, 0), np.int32), np.zeros((1, 0)), np.zeros((1, 0))),):
    t0 = time.perf_counter()
    w = gm.solve_wave(root, A0.shape[1], A0.shape[0], bv, bs, br)
    tm = gm.last_timing()
    print(json.dumps({"case": "C3 root wave", "tier": tm["tier"], "grid": tm["grid"], "status": int(w.status[0]),
                      "z": float(w.z[0]), "pivots": int(w.stats[0, 0] + w.stats[0, 1]), "bland": int(w.stats[0, 2]),
                      "kernel_ms": tm["kernel_ms"], "wall_s": time.perf_counter() - t0}), flush=True)
x = w.x[0]
frac = np.abs(x[:500] - np.round(x[:500]))
j = int(np.argmax(frac))
bv = np.array([[j], [j]], np.int32); bs = np.array([[1.0], [-1.0]]); br = np.array([[np.floor(x[j])], [-(np.floor(x[j]) + 1)]])
t0 = time.perf_counter()
w = gm.solve_wave(root, A0.shape[1], A0.shape[0], bv, bs, br)
tm = gm.last_timing()
print(json.dumps({"case": "C3 depth-1 wave (2 nodes)", "tier": tm["tier"], "grid": tm["grid"], "status": w.status.tolist(),
                  "z": w.z.tolist(), "pivots": (w.stats[:, 0] + w.stats[:, 1]).tolist(), "bland": w.stats[:, 2].tolist(),
                  "kernel_ms": tm["kernel_ms"], "wall_s": time.perf_counter() - t0}), flush=True)
gm.free_root(root)

# wide waves of HBM-resident LPs: one CTA per LP through the TMA-ring tier 4 vs the cooperative kernel with groups of 1
for (mm, nn_, cap) in ((1024, 2048, 60), (700, 1200, 60)):
    base = 4
    rng2 = np.random.default_rng(42)
    A4 = np.zeros((base, mm, nn_)); A4[:, :, : nn_ - mm] = rng2.random((base, mm, nn_ - mm)); A4[:, :, nn_ - mm:] = np.eye(mm)
    b4 = 1.0 + rng2.random((base, mm)); c4 = np.zeros((base, nn_)); c4[:, : nn_ - mm] = -rng2.random((base, nn_ - mm))
    reps = 148 // base
    c4, A4, b4 = np.tile(c4, (reps, 1)), np.tile(A4, (reps, 1, 1)), np.tile(b4, (reps, 1))
    run(f"148 slack-form LPs {mm}x{nn_}, {cap} pivots, tier 4 (TMA ring)", c4, A4, b4, force_tier=4, max_pivots=cap, refactor_period=100000)
    run(f"148 slack-form LPs {mm}x{nn_}, {cap} pivots, tier 6 G=1", c4, A4, b4, force_tier=6, coop_group=1, max_pivots=cap, refactor_period=100000)
    run(f"148 slack-form LPs {mm}x{nn_}, {cap} pivots, tier 6 G=2 (74 groups)", c4[:74], A4[:74], b4[:74], force_tier=6, coop_group=2, max_pivots=cap, refactor_period=100000)

# wide batches of mid-size LPs: the one-CTA-per-LP tiers 2 / 3 vs the cooperative kernel with groups of 1
for (mm, nn_, cnt, tier_) in ((150, 300, 296, 3), (100, 200, 296, 3), (70, 140, 592, 2), (150, 250, 2048, 3)):
    cB, AB, bB = feasible_bounded_lp(np.random.default_rng(5), mm, nn_, cnt)
    run(f"{cnt} LPs {mm}x{nn_} tier {tier_}", cB, AB, bB, force_tier=tier_)
    run(f"{cnt} LPs {mm}x{nn_} tier 6 G=1", cB, AB, bB, force_tier=6, coop_group=1)

# which B&B workloads run through without a solver-failure panic (reference semantics, tree.go:272)
from gomilp_b200 import status as S  # noqa: E402
for label, prob, lim in (("c5 n=50", c5_general_integer(50), 4095), ("c5 n=100", c5_general_integer(100), 2047),
                         ("c5 n=100 deep", c5_general_integer(100), 16383),
                         ("c5 n=200", c5_general_integer(200), 511), ("knapsack 60x10", knapsack(np.random.default_rng(7), 60, 10), 4095),
                         ("knapsack 120x40", knapsack(np.random.default_rng(7), 120, 40), 1023),
                         ("knapsack 500x200 (C3)", knapsack(np.random.default_rng(7), 500, 200), 63)):
    t0 = time.perf_counter()
    r = gm.milp_solve(prob["c"], None, None, prob["G"], prob["h"], prob["integrality"], mode=S.GM_BNB_FIXED | S.GM_BNB_DEVICE_SCAN,
                      heuristic=1, node_limit=lim, keep_log=False)
    dt = time.perf_counter() - t0
    print(json.dumps({"case": "bnb " + label, "budget": lim, "status": r.status, "lp_status": r.lp_status, "nodes": r.nodes,
                      "waves": r.waves, "pivots": r.pivots, "wall_s": dt, "nodes_per_sec": r.nodes / dt}), flush=True)
