"""GPU box: pivot-loop throughput of the HBM-resident tier on larger dense LPs in slack form [A I] x = b, b > 0 (the
initial slack basis is feasible, so no Phase I and no initial inversion dilute the measurement). Pivots are capped.
Prints pivots/s and the algorithmic GB/s (8*(3m^2 + m(n-m)) B per pivot), with and without the
TMA staging ring."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import gomilp_b200 as gm


def slack_form(rng, m, n, count):
    ns = n - m
    A = np.zeros((count, m, n))
    A[:, :, :ns] = rng.random((count, m, ns))
    A[:, :, ns:] = np.eye(m)
    b = 1.0 + rng.random((count, m))
    c = np.zeros((count, n))
    c[:, :ns] = -rng.random((count, ns))
    return c, A, b


def main():
    gm.init(0)
    for (m, n, count, cap) in [(70, 140, 592, 100), (100, 200, 296, 100), (150, 300, 296, 150), (256, 512, 296, 200), (512, 1024, 148, 200),
                               (700, 1200, 148, 150), (1024, 2048, 148, 100)]:
        rng = np.random.default_rng(42)
        base = min(count, 8)
        c, A, b = slack_form(rng, m, n, base)
        reps = (count + base - 1) // base
        c, A, b = np.tile(c, (reps, 1))[:count], np.tile(A, (reps, 1, 1))[:count], np.tile(b, (reps, 1))[:count]
        for no_ring in (False, True):
            gm.set_options(max_pivots=cap, no_tma_ring=no_ring, refactor_period=100000)
            g = gm.simplex_batch(c, A, b)   # warm
            g = gm.simplex_batch(c, A, b)
            tm = gm.last_timing()
            piv = int(g["pivots"].sum())
            bpp = 8 * (3 * m * m + m * (n - m))
            print(json.dumps({"m": m, "n": n, "count": count, "tma_ring": not no_ring, "tier": tm["tier"], "grid": tm["grid"],
                              "kernel_ms": round(tm["kernel_ms"], 2), "pivots": piv,
                              "us_per_pivot_per_cta": round(tm["kernel_ms"] * 1e3 / max(1, piv / tm["grid"]), 1),
                              "alg_GBps": round(piv * bpp / (tm["kernel_ms"] * 1e-3) / 1e9, 1),
                              "inversions": int(g["stats"][:, 3].sum()), "status": np.unique(g["status"]).tolist()}))
    gm.set_options()


if __name__ == "__main__":
    main()
