"""ncu target (not a test): ONE launch of the cooperative kernel on a batch of 16 dense LPs 1024 x 2048 (BASELINE config
4's shape, working set > L2), pivot-capped so that ~40 ncu replays stay short. A tiny warm-up launch loads the module.
   python tests/ncu_c4_probe.py [max_pivots]"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import gomilp_b200 as gm  # noqa: E402
from problems import feasible_bounded_lp  # noqa: E402

cap = int(sys.argv[1]) if len(sys.argv) > 1 else 400
gm.init(0)
lps = [feasible_bounded_lp(np.random.default_rng(42 + k), 1024, 2048) for k in range(16)]
c = np.stack([l[0] for l in lps]); A = np.stack([l[1] for l in lps]); b = np.stack([l[2] for l in lps])
gm.set_options(max_pivots=cap)
g = gm.simplex_batch(c, A, b, want_basis=False)
tm = gm.last_timing()
piv = int(g["pivots"].sum())
print({"tier": tm["tier"], "grid": tm["grid"], "pivots": piv, "kernel_ms": tm["kernel_ms"],
       "alg_GBps": piv * 8 * (3 * 1024 * 1024 + 1024 * 1024) / (tm["kernel_ms"] * 1e-3) / 1e9,
       "status": sorted(set(int(v) for v in g["status"]))})
