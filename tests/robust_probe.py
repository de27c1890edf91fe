import json, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import gomilp_b200 as gm
from gomilp_b200 import status as S
from problems import knapsack
gm.init(0)
for (n, m, lim) in ((120, 40, 1023), (500, 200, 63), (500, 200, 255)):
    p = knapsack(np.random.default_rng(7), n, m)
    t0 = time.perf_counter()
    r = gm.milp_solve(p["c"], None, None, p["G"], p["h"], p["integrality"], mode=S.GM_BNB_FIXED | S.GM_BNB_DEVICE_SCAN | S.GM_BNB_ROBUST,
                      heuristic=1, node_limit=lim, keep_log=False)
    print(json.dumps({"case": f"robust bnb knapsack {n}x{m}", "budget": lim, "status": r.status, "lp_status": r.lp_status, "nodes": r.nodes,
                      "waves": r.waves, "pivots": r.pivots, "wall_s": time.perf_counter() - t0, "z": r.z if r.x is not None else None}), flush=True)
