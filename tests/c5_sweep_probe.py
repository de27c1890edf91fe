"""GPU probe (not a test): BASELINE config 5, the general-integer MILP sweep n = 50 .. 400 on one GPU: B&B nodes/s with
cold and with warm-started children (device-side scan), beside the serial C++ oracle on one host core for a bounded
time. Writes one JSON line per n."""
import json, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import gomilp_b200 as gm
import oracle
from problems import c5_general_integer
gm.init(0)
BUDGET = {50: 16383, 100: 16383, 200: 2047, 400: 255}
ns = [int(a) for a in sys.argv[1:]] or [50, 100, 200, 400]
for n in ns:
    p = c5_general_integer(n)
    def run(mode, limit):
        t0 = time.perf_counter()
        r = gm.milp_solve(p["c"], None, None, p["G"], p["h"], p["integrality"], mode=mode, heuristic=1, node_limit=limit,
                          keep_log=False)
        dt = time.perf_counter() - t0
        return {"nodes": r.nodes, "waves": r.waves, "pivots": r.pivots, "status": r.status, "lp_status": r.lp_status,
                "z": r.z, "wall_s": dt, "device_ms": r.device_ms, "nodes_per_sec": r.nodes / dt}
    run(1 | 8, 63); run(1 | 4 | 8, 63)  # loads the kernels
    row = {"n": n, "lp_shape": [n // 2 + n, n // 2 + 2 * n], "node_budget": BUDGET[n]}
    row["cold"] = min((run(1 | 8, BUDGET[n]) for _ in range(2)), key=lambda r: r["wall_s"])
    row["warm"] = min((run(1 | 4 | 8, BUDGET[n]) for _ in range(2)), key=lambda r: r["wall_s"])
    if n > 100:  # one LP of these sizes takes the serial oracle longer than the whole bounded sample
        print(json.dumps(row), flush=True)
        continue
    t0 = time.perf_counter()
    o = oracle.bnb_solve(p["c"], None, None, p["G"], p["h"], p["integrality"], heuristic=1, mode=1, node_limit=BUDGET[n],
                         time_limit_s=6.0, log_cap=1)
    dt = time.perf_counter() - t0
    row["oracle_1_core"] = {"nodes": o.nodes, "wall_s": dt, "nodes_per_sec": o.nodes / dt, "status": o.status}
    print(json.dumps(row), flush=True)
