"""One warm-up + one measured launch of a mid/large batch, for ncu attribution of tiers 3 / 4.
usage: python tests/ncu_tier_probe.py m n count cap"""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import gomilp_b200 as gm
from hbm_probe import slack_form

m, n, count, cap = (int(v) for v in sys.argv[1:5])
gm.init(0)
rng = np.random.default_rng(42)
c, A, b = slack_form(rng, m, n, min(count, 8))
reps = (count + c.shape[0] - 1) // c.shape[0]
c, A, b = np.tile(c, (reps, 1))[:count], np.tile(A, (reps, 1, 1))[:count], np.tile(b, (reps, 1))[:count]
gm.set_options(max_pivots=cap, refactor_period=100000)
for _ in range(2):
    g = gm.simplex_batch(c, A, b)
print(gm.last_timing(), int(g["pivots"].sum()))
