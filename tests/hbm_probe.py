"""GPU box: throughput of the HBM-resident tiers (3/4) on larger dense LPs; pivots are capped so that the probe is
bounded. Prints pivots/s and the algorithmic GB/s (8*(3m^2 + m(n-m)) B per pivot)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import gomilp_b200 as gm
from problems import feasible_bounded_lp

gm.init(0)
for (m, n, count, cap) in [(96, 192, 296, 0), (150, 300, 296, 0), (256, 512, 148, 300), (512, 1024, 148, 200), (1024, 2048, 16, 100),
                           (1024, 2048, 148, 100)]:
    rng = np.random.default_rng(42)
    c, A, b = feasible_bounded_lp(rng, m, n, min(count, 16))
    reps = (count + c.shape[0] - 1) // c.shape[0]
    c, A, b = np.tile(c, (reps, 1))[:count], np.tile(A, (reps, 1, 1))[:count], np.tile(b, (reps, 1))[:count]
    gm.set_options(max_pivots=cap)
    g = gm.simplex_batch(c, A, b)   # warm
    t0 = time.perf_counter()
    g = gm.simplex_batch(c, A, b)
    dt = time.perf_counter() - t0
    tm = gm.last_timing()
    piv = int(g["pivots"].sum())
    bpp = 8 * (3 * m * m + m * (n - m))
    print(json.dumps({"m": m, "n": n, "count": count, "tier": tm["tier"], "grid": tm["grid"], "kernel_ms": tm["kernel_ms"],
                      "pivots": piv, "pivots_per_s": piv / (tm["kernel_ms"] * 1e-3), "us_per_pivot_per_cta":
                      tm["kernel_ms"] * 1e3 / max(1, piv / tm["grid"]), "alg_GBps": piv * bpp / (tm["kernel_ms"] * 1e-3) / 1e9,
                      "inversions": int(g["stats"][:, 3].sum()), "status": np.bincount(g["status"], minlength=1).tolist()[:3]}))
gm.set_options()
