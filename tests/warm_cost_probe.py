"""GPU probe (not a test): what a warm-started wave of C5 nodes spends (stats per node)."""
import json, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import gomilp_b200 as gm
from problems import c5_general_integer, standard_form, most_infeasible
gm.init(0)
p = c5_general_integer(100)
c0, A0, b0 = standard_form(p)
m0, n0 = A0.shape
integ = np.concatenate([p["integrality"], np.zeros(n0 - 100, np.uint8)])
root = gm.upload_root(c0, A0, b0)
waves = [(np.zeros((1, 0), np.int32), np.zeros((1, 0)), np.zeros((1, 0)), None)]
for L in range(0, 11):
    bv, bs, br, par = waves[-1]
    t0 = time.perf_counter()
    gm.profile_arm()
    w = gm.solve_wave(root, n0, m0, bv, bs, br, parent=par, warm=True)
    tm = gm.last_timing()
    pr = gm.profile_fetch(len(bv)).astype(float)
    prm = pr.mean(axis=0) if len(pr) else np.zeros(16)
    s = w.stats
    print(json.dumps({"L": L, "nodes": len(bv), "tier": tm["tier"], "kernel_ms": tm["kernel_ms"], "launches": tm["launches"],
                      "piv_per_node": float((s[:, 0] + s[:, 1]).mean()), "p1_per_node": float(s[:, 0].mean()), "inv_per_node": float(s[:, 3].mean()),
                      "bland_per_node": float(s[:, 2].mean()), "repair_per_node": float(s[:, 6].mean()), "used_p1": float(s[:, 4].mean()),
                      "ok": int((w.status == 0).sum()), "leader_kcyc": {"solve": round(prm[0] / 1e3), "main": round(prm[1] / 1e3), "inv": round(prm[2] / 1e3),
                      "polish": round(prm[3] / 1e3), "bland": round(prm[4] / 1e3), "entries": round(prm[6], 1), "polishes": round(prm[7], 1), "checks": round(prm[8] / 1e3),
                      "warm_start": round(prm[9] / 1e3), "basis_setup": round(prm[10] / 1e3), "repair": round(prm[11] / 1e3), "results": round(prm[12] / 1e3),
                      "loop_entry_exit": round(prm[13] / 1e3)}, "us_per_node_per_cta": 1e3 * tm["kernel_ms"] * min(148, len(bv)) / len(bv)}), flush=True)
    nb, ns, nr, npar = [], [], [], []
    for k in range(len(bv)):
        if w.status[k] != 0:
            continue
        j = most_infeasible(w.x[k], integ)
        if j < 0:
            continue
        fl = np.floor(w.x[k][j])
        for sg, rh in ((1.0, fl), (-1.0, -(fl + 1))):
            nb.append(np.append(bv[k], j)); ns.append(np.append(bs[k], sg)); nr.append(np.append(br[k], rh)); npar.append(k)
    if not nb:
        break
    waves.append((np.array(nb, np.int32), np.array(ns), np.array(nr), np.array(npar, np.int32)))
gm.free_root(root)
