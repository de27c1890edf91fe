"""Small launches through every tier, for compute-sanitizer (racecheck / memcheck) on the GPU box."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import gomilp_b200 as gm
from problems import feasible_bounded_lp, raw_lp, knapsack
from gomilp_b200 import status as S

gm.init(0)
rng = np.random.default_rng(3)
c, A, b = feasible_bounded_lp(rng, 12, 30, 6)
c2, A2, b2 = raw_lp(rng, 7, 15, 6, 0.2)
for tier in (1, 2, 3, 4, 5):
    gm.set_options(force_tier=tier)
    g = gm.simplex_batch(c, A, b)
    g2 = gm.simplex_batch(c2, A2, b2)
    print("tier", tier, g["status"].tolist(), g2["status"].tolist())
gm.set_options()
# TMA ring tier on a long-row problem (few pivots)
m, n = 400, 520
A4 = np.zeros((2, m, n)); A4[:, :, : n - m] = rng.random((2, m, n - m)); A4[:, :, n - m:] = np.eye(m)
gm.set_options(max_pivots=12)
g4 = gm.simplex_batch(np.concatenate([-rng.random((2, n - m)), np.zeros((2, m))], axis=1), A4, 1.0 + rng.random((2, m)))
print("tier", gm.last_timing()["tier"], g4["status"].tolist())
gm.set_options()
p = knapsack(rng, 8, 2)
for mode in (1, 5):
    r = gm.milp_solve(p["c"], None, None, p["G"], p["h"], p["integrality"], mode=mode, heuristic=1, node_limit=40)
    print("milp mode", mode, r.status, r.nodes)
