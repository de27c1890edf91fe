import sys, time, json
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import gomilp_b200 as gm
from problems import knapsack
gm.init(0)
p = knapsack(np.random.default_rng(7), 30, 5)
for mode in (1, 5, 1, 5):
    t0 = time.perf_counter()
    r = gm.milp_solve(p["c"], None, None, p["G"], p["h"], p["integrality"], mode=mode, heuristic=1, node_limit=8192)
    dt = time.perf_counter() - t0
    print(mode, r.status, r.nodes, r.pivots, round(dt*1e3,1), 'ms', round(r.device_ms,2), [(w[1], w[2], round(w[3],3)) for w in r.waves_log])
