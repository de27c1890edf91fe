"""Timing probe (GPU box): B&B node throughput through gm_milp_solve on knapsack / general-integer instances."""
import sys, time, json
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import gomilp_b200 as gm
from problems import knapsack
from gomilp_b200 import status as S

def run(n, m, node_limit, seed=7):
    rng = np.random.default_rng(seed)
    p = knapsack(rng, n, m)
    t0 = time.perf_counter()
    r = gm.milp_solve(p["c"], None, None, p["G"], p["h"], p["integrality"], mode=S.GM_BNB_FIXED, heuristic=S.GM_BRANCH_MOST_INFEASIBLE, node_limit=node_limit)
    dt = time.perf_counter() - t0
    print(json.dumps({"knapsack": [n, m], "std_form": [m + n, m + 2 * n], "status": r.status, "lp_status": r.lp_status, "last_log": r.log[-1] if r.log else None, "nodes": r.nodes, "waves": r.waves,
                      "pivots": r.pivots, "wall_s": dt, "device_ms": r.device_ms, "nodes_per_s": r.nodes / dt,
                      "waves_log": r.waves_log[:12]}))

if __name__ == "__main__":
    gm.init(0)
    for (n, m, lim) in [(30, 5, 2000), (60, 10, 2000), (120, 30, 1000)]:
        run(n, m, lim)
