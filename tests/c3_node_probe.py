"""GPU probe (not a test): isolate the C3 node on which the search stops and look at its pivot trace."""
import json, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import gomilp_b200 as gm
from gomilp_b200 import status as S
from problems import knapsack, standard_form
gm.init(0)
p = knapsack(np.random.default_rng(7), 500, 200)
r = gm.milp_solve(p["c"], None, None, p["G"], p["h"], p["integrality"], mode=S.GM_BNB_FIXED | S.GM_BNB_ROBUST, heuristic=1,
                  node_limit=int(sys.argv[1]) if len(sys.argv) > 1 else 7)
print("replay:", r.status, r.lp_status, r.nodes, r.pivots)
# rebuild the descriptors of the children of the logged nodes
nodes = {0: []}
order = []
nxt = 0
for (id_, parent, depth, st, z, dec, bv, bf) in r.log:
    order.append(id_)
    if dec == 4:
        for ch in range(2):
            nxt += 1
            nodes[nxt] = nodes[id_] + [(bv, 1.0 if ch == 0 else -1.0, bf if ch == 0 else -(bf + 1.0))]
print("logged", len(r.log), "nodes; children known:", sorted(nodes))
c0, A0, b0 = standard_form(p)
root = gm.upload_root(c0, A0, b0)
gm.set_options(max_pivots=12000)
for nid in sorted(nodes):
    d = nodes[nid]
    if nid < (int(sys.argv[2]) if len(sys.argv) > 2 else 7) or nid > (int(sys.argv[3]) if len(sys.argv) > 3 else 14):
        continue
    bvar = np.array([[t[0] for t in d]], np.int32); bs = np.array([[t[1] for t in d]]); br = np.array([[t[2] for t in d]])
    for robust in ((False, True) if len(sys.argv) <= 1 else (True,)):
        gm.set_options(max_pivots=0 if robust else 12000, robust=robust)
        gm.trace_arm(0, 12000)
        t0 = time.perf_counter()
        w = gm.solve_wave(root, A0.shape[1], A0.shape[0], bvar, bs, br)
        tr = gm.trace_fetch(12000)
        s = w.stats[0]
        per = None
        if len(tr) > 2000:   # period of the (enter, leave) sequence at the end of the trace
            tail = [tuple(x[1:3]) for x in tr[-1500:]]
            for P_ in range(1, 600):
                if all(tail[i] == tail[i - P_] for i in range(len(tail) - 600, len(tail))):
                    per = P_
                    break
        print(json.dumps({"node": nid, "robust": robust, "status": int(w.status[0]), "z": float(w.z[0]), "piv1": int(s[0]), "piv2": int(s[1]),
                          "bland": int(s[2]), "inv": int(s[3]), "flags": int(s[5]), "repair": int(s[6]), "trace_len": len(tr), "period": per,
                          "bland_in_last_1000": int(tr[-1000:, 3].sum()) if len(tr) else 0, "phase_last": int(tr[-1, 0]) if len(tr) else 0,
                          "ms": 1e3 * (time.perf_counter() - t0)}), flush=True)
gm.set_options()
gm.free_root(root)
