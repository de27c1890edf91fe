"""ncu target (not a test): 296 LPs 150 x 300 (the shape of a wide C5 wave) through one tier. python tests/ncu_mid_probe.py <tier> [G]"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import gomilp_b200 as gm  # noqa: E402
from problems import feasible_bounded_lp  # noqa: E402

tier = int(sys.argv[1])
G = int(sys.argv[2]) if len(sys.argv) > 2 else 0
gm.init(0)
c, A, b = feasible_bounded_lp(np.random.default_rng(5), 150, 300, 296)
gm.set_options(force_tier=tier, coop_group=G)
g = gm.simplex_batch(c, A, b, want_basis=False)
tm = gm.last_timing()
print({"tier": tm["tier"], "grid": tm["grid"], "pivots": int(g["pivots"].sum()), "kernel_ms": tm["kernel_ms"],
       "status": sorted(set(int(v) for v in g["status"]))})
