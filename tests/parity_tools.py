"""Decision-level parity tools (TEST INFRASTRUCTURE): compare two pivot traces of the same LP and classify the first
divergence. BASELINE.json asks for "identical node counts and branching sequences where pivots tie-break
identically": a trace row is (phase, entering variable, leaving variable, bland) - what simplex.go:233-293 decides
per iteration. Two implementations that make Gonum's decisions on different arithmetic (three fresh LUs per pivot
there, a product-form inverse here) can only part ways where the quantities compared by floats.MinIdx are equal up
to rounding: reduced costs r_e (simplex.go:244) or ratios xb_i / |d_i| (:268, :336-340). `classify_divergence`
recomputes those quantities in numpy at the last common basis and says which it was - or "REAL" when the two choices
are not a near-tie, which would be a bug.
"""
from __future__ import annotations

import numpy as np

TIE = 1e-9


def first_divergence(tr_a: np.ndarray, tr_b: np.ndarray) -> int:
    """Index of the first row where the traces differ in (phase, enter, leave); -1 when one is a prefix of the other
    and they have the same length."""
    k = min(len(tr_a), len(tr_b))
    for i in range(k):
        if not np.array_equal(tr_a[i, :3], tr_b[i, :3]):
            return i
    return -1 if len(tr_a) == len(tr_b) else k


def _phase1_start(A, b, basis0):
    """Basis set and artificial column after simplex.go:529-553 (the artificial replaces position argmin xb)."""
    B0 = A[:, basis0]
    xb = np.linalg.solve(B0, b)
    j = int(np.argmin(xb))
    art = b - (B0.sum(axis=1) - B0[:, j])
    S = [int(v) for v in basis0]
    S[j] = A.shape[1]
    return S, art, bool((xb >= -1e-13).all())


def classify_divergence(c, A, b, tr_ref, tr_got, k, basis0, final_basis_ref=None) -> str:
    """Why do the traces differ at row k? Returns 'entering-near-tie', 'ratio-near-tie', 'bland', 'length' or 'REAL'."""
    c, A, b = (np.asarray(v, dtype=np.float64) for v in (c, A, b))
    m, n = A.shape
    if k >= len(tr_ref) or k >= len(tr_got):
        return "length"
    if tr_ref[k, 3] or tr_got[k, 3]:
        return "bland"   # replaceBland only runs on a degenerate step (move <= 0): the vertex is degenerate
    phase = int(tr_ref[k, 0])
    if int(tr_got[k, 0]) != phase:
        return "phase-boundary"
    S, art, feasible0 = _phase1_start(A, b, basis0)
    if feasible0:
        S = [int(v) for v in basis0]
    # walk forward through the common prefix
    fwd_ok = True
    for i in range(k):
        ph, e, l, _ = (int(v) for v in tr_ref[i])
        if i > 0 and ph != int(tr_ref[i - 1, 0]):
            fwd_ok = False  # Phase I -> II: the repair loop may have changed the basis without a trace row
            break
        if l not in S:
            fwd_ok = False
            break
        S[S.index(l)] = e
    if not fwd_ok:
        if final_basis_ref is None or len(tr_ref) == 0:
            return "unclassified"
        Sset = set(int(v) for v in final_basis_ref)
        for i in range(len(tr_ref) - 1, k - 1, -1):
            ph, e, l, _ = (int(v) for v in tr_ref[i])
            Sset = (Sset - {e}) | {l}
        S = sorted(Sset)
    Afull = np.hstack([A, art[:, None]])
    cost = np.zeros(n + 1)
    if phase == 1:
        cost[n] = 1.0
    else:
        cost[:n] = c
    ncols = n + 1 if phase == 1 else n
    B = Afull[:, S]
    try:
        y = np.linalg.solve(B.T, cost[S])
        xb = np.linalg.solve(B, b)
    except np.linalg.LinAlgError:
        return "unclassified"
    nonbasic = [j for j in range(ncols) if j not in set(S)]
    r = {j: cost[j] - Afull[:, j] @ y for j in nonbasic}
    e_ref, e_got = int(tr_ref[k, 1]), int(tr_got[k, 1])
    scale = max(1.0, float(np.abs(cost).max()))
    if e_ref != e_got:
        if e_ref in r and e_got in r and abs(r[e_ref] - r[e_got]) <= TIE * scale:
            return "entering-near-tie"
        return "REAL"
    d = np.linalg.solve(B, Afull[:, e_ref])
    l_ref, l_got = int(tr_ref[k, 2]), int(tr_got[k, 2])
    if l_ref not in S or l_got not in S:
        return "REAL"
    i_ref, i_got = S.index(l_ref), S.index(l_got)
    if d[i_ref] <= 0 or d[i_got] <= 0:
        return "REAL"
    q_ref, q_got = xb[i_ref] / d[i_ref], xb[i_got] / d[i_got]
    if abs(q_ref - q_got) <= TIE * max(1.0, abs(q_ref)):
        return "ratio-near-tie"
    return "REAL"


def compare_bnb_logs(ref_log: dict, got_log: list, rtol: float = 1e-9):
    """Node-by-node comparison of two decision logs (FIFO order). ref_log: oracle.bnb_solve(...).log (dict of arrays),
    got_log: MilpResult.log rows (id, parent, depth, lp_status, z, decision, branch_var, branch_floor).
    Returns (identical, first_divergent_node, why)."""
    nref = len(ref_log["id"])
    for k in range(min(nref, len(got_log))):
        g = got_log[k]
        if int(ref_log["id"][k]) != g[0] or int(ref_log["parent"][k]) != g[1]:
            return False, k, "ids"
        if int(ref_log["lp_status"][k]) != g[3]:
            return False, k, f"lp_status {int(ref_log['lp_status'][k])} vs {g[3]}"
        zr, zg = float(ref_log["z"][k]), g[4]
        if g[3] == 0 and not (abs(zr - zg) <= rtol * max(1.0, abs(zr))):
            return False, k, f"z {zr!r} vs {zg!r}"
        if int(ref_log["decision"][k]) != g[5]:
            return False, k, f"decision {int(ref_log['decision'][k])} vs {g[5]} (z {zr!r} vs {zg!r})"
        if g[5] == 4 and (int(ref_log["branch_var"][k]) != g[6] or float(ref_log["branch_floor"][k]) != g[7]):
            return False, k, f"branch ({int(ref_log['branch_var'][k])}, {float(ref_log['branch_floor'][k])}) vs ({g[6]}, {g[7]})"
    if nref != len(got_log):
        return False, min(nref, len(got_log)), f"length {nref} vs {len(got_log)}"
    return True, -1, ""
