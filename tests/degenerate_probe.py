"""GPU probe (not a test): the degenerate 105 x 137 LP of the cross-tier test under every tier / group size."""
import json, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import gomilp_b200 as gm
gm.init(0)
z = np.load("tests/golden/degenerate_105x137.npz"); c, A, b = z["c"], z["A"], z["b"]
for (tier, G, robust) in ((5, 0, False), (3, 0, False), (6, 1, False), (6, 2, False), (6, 5, False), (6, 37, False), (6, 1, True), (5, 0, True)):
    gm.set_options(force_tier=tier, coop_group=G, robust=robust)
    g = gm.simplex_batch(c, A, b)
    s = g["stats"][0]
    print(json.dumps({"tier": tier, "G": G, "robust": robust, "status": int(g["status"][0]), "z": float(g["optF"][0]), "piv": int(s[0] + s[1]),
                      "bland": int(s[2]), "inv": int(s[3]), "flags": int(s[5]), "repair": int(s[6])}), flush=True)
gm.set_options()
