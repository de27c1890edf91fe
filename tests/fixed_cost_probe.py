"""GPU probe (not a test): per-LP fixed cost (setup, polish, results) vs per-pivot cost on a wide wave of C5 nodes."""
import json, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import gomilp_b200 as gm
from problems import c5_general_integer, standard_form
gm.init(0)
p = c5_general_integer(100)
c0, A0, b0 = standard_form(p)
m0, n0 = A0.shape
root = gm.upload_root(c0, A0, b0)
nodes = 1184
L = 3
rng = np.random.default_rng(1)
bvar = rng.integers(0, 100, size=(nodes, L)).astype(np.int32)
bsign = np.ones((nodes, L)); brhs = rng.integers(3, 8, size=(nodes, L)).astype(float)
for tier, G in ((3, 0), (6, 1)):
    for cap in (1, 4, 16, 64, 0):
        gm.set_options(force_tier=tier, coop_group=G, max_pivots=cap)
        gm.solve_wave(root, n0, m0, bvar, bsign, brhs)
        t0 = time.perf_counter()
        w = gm.solve_wave(root, n0, m0, bvar, bsign, brhs)
        tm = gm.last_timing()
        piv = int((w.stats[:, 0] + w.stats[:, 1]).sum())
        print(json.dumps({"tier": tm["tier"], "cap": cap, "kernel_ms": tm["kernel_ms"], "pivots": piv,
                          "us_per_lp_per_cta": 1e3 * tm["kernel_ms"] * 148 / nodes, "status": sorted(set(int(v) for v in w.status))}), flush=True)
gm.set_options()
gm.free_root(root)
