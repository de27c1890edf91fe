"""Multi-GPU branch-and-bound: one process per GPU, the frontier wave sharded in contiguous FIFO blocks.

Mirror of ``enumerationTree.startSearch`` / ``checkSolution`` (tree.go:66-123, 207-263) for N ranks:

* every rank holds the whole tree state (node descriptors are tiny: L x 20 B per node) and runs the same
  deterministic scheduler, so no rank is a master;
* per wave (= one BFS level = one FIFO segment) rank r solves nodes
  ``fifo_block(count, world, r)`` on its own GPU through ``gm_solve_wave`` — the data path has no collective;
* each rank then reduces its block to one small record per node — (lp status, z, integer-feasible?, branch
  variable, floor) — and the ranks exchange ONLY those records (``all_gather``; 32 B per node, latency
  bound over NVLink/NVSwitch) plus, when the incumbent improves, the incumbent's x from its owner;
* ``checkSolution`` is then replayed over the full wave in FIFO order on every rank, which keeps pruning
  global and reproduces the 1-worker reference order: node i is pruned against the best integer-feasible z
  among the incumbent and nodes < i of the wave (an exclusive prefix-min — a wave-global min would prune
  nodes the reference branches on).

The collective runs on whatever ``torch.distributed`` backend the group has: NCCL on the GPU box
(``device`` = the rank's cuda device), gloo in the CPU tests.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field

import hashlib

import numpy as np

from . import status as S


def fifo_block(count: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous FIFO block [lo, hi) of a wave of ``count`` nodes owned by ``rank`` (sizes differ by <= 1)."""
    base, extra = divmod(count, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def feasible_for_ip(integ: np.ndarray, x: np.ndarray) -> bool:
    """tree.go:276-297: exact x == trunc(x) on the integer-flagged entries."""
    xi = x[integ.astype(bool)]
    return bool(np.all(xi == np.trunc(xi)))


def maxfun_point(c: np.ndarray, integ: np.ndarray) -> int:
    """branching.go:54-72 as it behaves: the last integer-flagged index (candidateValue is never updated)."""
    idx = np.nonzero(integ.astype(bool) & (np.abs(c) >= 0))[0]
    return int(idx[-1]) if idx.size else 0


def fixed_point(heuristic: int, c: np.ndarray, x: np.ndarray, integ: np.ndarray, last_var: int) -> int:
    """FIXED-mode heuristics (same rules as csrc/bnb_host.cpp)."""
    n = x.shape[0]
    frac = integ.astype(bool) & (x != np.trunc(x))
    if not frac.any():
        return -1
    if heuristic == S.GM_BRANCH_NAIVE:
        start = 0 if last_var < 0 else (last_var + 1) % n
        order = (start + np.arange(n)) % n
        return int(order[np.nonzero(frac[order])[0][0]])
    if heuristic == S.GM_BRANCH_MOST_INFEASIBLE:
        f = x - np.floor(x)
        score = 0.5 - np.abs(0.5 - f)
    else:
        score = np.abs(c)
    score = np.where(frac, score, -1.0)
    return int(np.argmax(score))  # first maximum


@dataclass
class ShardedResult:
    status: int
    lp_status: int
    x: np.ndarray | None
    z: float
    nodes: int
    waves: int
    pivots: int
    device_ms: float
    decisions: list = field(default_factory=list)   # (id, parent, depth, lp_status, z, decision, branch_var, floor)
    exchange_bytes: int = 0


def milp_solve_sharded(c, A, b, G, h, integrality, *, solve_wave, group=None, device=None, heuristic=0, mode=0,
                       node_limit=0, time_limit_s=0.0) -> ShardedResult:
    """``milpProblem.solve`` (ilp.go:75-116) over the ranks of ``group``.

    ``solve_wave(c0, A0, b0, bvar, bsign, brhs) -> (status, z, x, pivots, kernel_ms)`` solves a block of nodes of
    one depth on this rank's device (``gomilp_b200.sharded.gpu_wave_solver()`` builds it from the C ABI).
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    dev = device if device is not None else torch.device("cpu")

    c = np.asarray(c, dtype=np.float64)
    nvar = c.shape[0]
    A = None if A is None else np.asarray(A, dtype=np.float64).reshape(-1, nvar)
    G = None if G is None else np.asarray(G, dtype=np.float64).reshape(-1, nvar)
    meq = 0 if A is None else A.shape[0]
    nineq = 0 if G is None else G.shape[0]
    m0, n0 = meq + nineq, nvar + nineq
    # toInitialSubproblem + convertToEqualities (ilp.go:43-71, subproblem.go:81-139)
    A0 = np.zeros((m0, n0))
    b0 = np.zeros(m0)
    c0 = np.concatenate([c, np.zeros(nineq)])
    if meq:
        A0[:meq, :nvar] = A
        b0[:meq] = b
    if nineq:
        A0[meq:, :nvar] = G
        A0[meq:, nvar:] = np.eye(nineq)
        b0[meq:] = h
    integ = np.concatenate([np.asarray(integrality, dtype=np.uint8), np.zeros(nineq, dtype=np.uint8)])

    t0 = time.perf_counter()
    res = ShardedResult(S.GM_MILP_OK, 0, None, 0.0, 0, 0, 0, 0.0)
    inc_z, inc_x = math.inf, None
    ids, parents = [0], [0]
    bvar = np.zeros((1, 0), dtype=np.int32)
    bsign = np.zeros((1, 0))
    brhs = np.zeros((1, 0))
    next_id, depth, timed_out = 0, 0, False
    REC = 6  # status, z, feasible, branch var, floor, pivots

    while len(ids) > 0:
        count = len(ids)
        if node_limit > 0:
            left = node_limit - res.nodes
            if left <= 0:
                timed_out = True
                break
            if count > left:
                count, timed_out = left, True
        if time_limit_s > 0 and depth > 0:
            stop = torch.tensor([1.0 if time.perf_counter() - t0 > time_limit_s else 0.0], device=dev)
            if world > 1:
                dist.all_reduce(stop, op=dist.ReduceOp.MAX, group=group)  # every rank must take the same branch
            if stop.item() > 0:
                timed_out = True
                break
        lo, hi = fifo_block(count, world, rank)
        rec = np.zeros((hi - lo, REC))
        xs = np.zeros((hi - lo, n0))
        kms = 0.0
        if hi > lo:
            st, z, xs, piv, kms = solve_wave(c0, A0, b0, bvar[lo:hi], bsign[lo:hi], brhs[lo:hi])
            # per-node record, vectorised over the block: integer-feasible? else branch variable and floor
            imask = integ.astype(bool)
            okm = st == S.GM_OK
            frac = (xs != np.trunc(xs)) & imask[None, :]
            feas = okm & ~frac.any(axis=1)
            need = okm & ~feas
            bvv = np.full(hi - lo, -1.0)
            flv = np.zeros(hi - lo)
            if need.any():
                if mode == S.GM_BNB_COMPAT:
                    on = np.full(hi - lo, maxfun_point(c0, integ))
                elif heuristic == S.GM_BRANCH_NAIVE:
                    on = np.array([fixed_point(heuristic, c0, xs[k], integ, int(bvar[lo + k, -1]) if depth > 0 else -1)
                                   if need[k] else 0 for k in range(hi - lo)])
                else:
                    if heuristic == S.GM_BRANCH_MOST_INFEASIBLE:
                        f = xs - np.floor(xs)
                        score = 0.5 - np.abs(0.5 - f)
                    else:
                        score = np.broadcast_to(np.abs(c0), xs.shape)
                    on = np.argmax(np.where(frac, score, -1.0), axis=1)  # first maximum, like fixed_point()
                idx = np.nonzero(need)[0]
                bvv[idx] = on[idx]
                flv[idx] = np.floor(xs[idx, on[idx].astype(np.int64)])
            rec = np.stack([st.astype(np.float64), z, feas.astype(np.float64), bvv, flv, piv.astype(np.float64)], axis=1)
        # ---- the only exchange of the wave: one 48-byte record per node -------------------------------
        if world > 1:
            sizes = [fifo_block(count, world, r) for r in range(world)]
            maxlen = max(h_ - l_ for l_, h_ in sizes)
            mine = torch.zeros(maxlen, REC, dtype=torch.float64, device=dev)
            if hi > lo:
                mine[: hi - lo] = torch.from_numpy(rec).to(dev)
            gathered = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(gathered, mine, group=group)
            full = np.concatenate([gathered[r][: sizes[r][1] - sizes[r][0]].cpu().numpy() for r in range(world)])
            res.exchange_bytes += maxlen * REC * 8 * world
            t = torch.tensor([kms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            kms = float(t.item())
        else:
            full = rec
        res.waves += 1
        res.device_ms += kms

        # ---- checkSolution over the whole wave in FIFO order, identically on every rank ---------------
        nid, npar, nbv, nbs, nbr = [], [], [], [], []
        new_inc_owner = -1
        panic = 0
        root_done = False
        for k in range(count):
            st, zk, feas, bv, fl, piv = full[k]
            st = int(st)
            res.nodes += 1
            res.pivots += int(piv)
            if depth == 0:
                if st != S.GM_OK:  # subproblem.go:173-176
                    panic, res.lp_status = S.GM_MILP_PANIC_ROOT, st
                    break
                if feas:  # tree.go:88-92
                    res.decisions.append((0, 0, 0, st, zk, S.GM_DEC_INITIAL_RX_FEASIBLE_FOR_IP, -1, 0.0))
                    inc_z, new_inc_owner, root_done = zk, k, True
                    break
            decision, dbv, dfl = S.GM_DEC_NONE, -1, 0.0
            if st != S.GM_OK:
                if st == S.GM_ERR_INFEASIBLE:
                    decision = S.GM_DEC_SUBPROBLEM_IS_DEGENERATE
                elif st == S.GM_ERR_SINGULAR:
                    decision = S.GM_DEC_SUBPROBLEM_NOT_FEASIBLE
                else:
                    panic, res.lp_status = S.GM_MILP_PANIC_SOLVER_FAILURE, st
                    break
            elif inc_z <= zk:
                decision = S.GM_DEC_WORSE_THAN_INCUMBENT
            elif inc_z > zk:
                if feas:
                    inc_z, new_inc_owner = zk, k
                    decision = S.GM_DEC_BETTER_THAN_INCUMBENT_FEASIBLE
                else:
                    on = int(bv)
                    for child in range(2):
                        next_id += 1
                        nid.append(next_id)
                        npar.append(ids[k])
                        nbv.append(np.concatenate([bvar[k], [on]]))
                        nbs.append(np.concatenate([bsign[k], [1.0 if child == 0 else -1.0]]))
                        nbr.append(np.concatenate([brhs[k], [fl if child == 0 else -(fl + 1.0)]]))
                    decision, dbv, dfl = S.GM_DEC_BETTER_THAN_INCUMBENT_BRANCHING, on, fl
            else:
                panic, res.lp_status = S.GM_MILP_PANIC_UNEXPECTED_CASE, st
                break
            res.decisions.append((ids[k], parents[k], depth, st, zk, decision, dbv, dfl))
        # the incumbent's x travels only when it changed in this wave, from the rank that solved that node
        if new_inc_owner >= 0:
            owner = next(r for r in range(world) if fifo_block(count, world, r)[0] <= new_inc_owner
                         < fifo_block(count, world, r)[1])
            xt = torch.zeros(n0, dtype=torch.float64, device=dev)
            if rank == owner:
                xt = torch.from_numpy(np.ascontiguousarray(xs[new_inc_owner - lo])).to(dev)
            if world > 1:
                dist.broadcast(xt, src=dist.get_global_rank(group, owner) if group is not None else owner, group=group)
                res.exchange_bytes += n0 * 8
            inc_x = xt.cpu().numpy()
        if panic:
            res.status = panic
            getattr(solve_wave, "close", lambda: None)()
            return res
        if root_done or timed_out:
            break
        ids, parents = nid, npar
        depth += 1
        bvar = np.array(nbv, dtype=np.int32).reshape(len(nid), depth)
        bsign = np.array(nbs, dtype=np.float64).reshape(len(nid), depth)
        brhs = np.array(nbr, dtype=np.float64).reshape(len(nid), depth)

    getattr(solve_wave, "close", lambda: None)()
    if timed_out:  # ilp.go:92-99: the incumbent is returned as is, slack entries included
        res.status = S.GM_MILP_DEADLINE_EXCEEDED
        if inc_x is not None:
            res.x, res.z = inc_x.copy(), inc_z
        return res
    if inc_x is None:
        res.status = S.GM_MILP_NO_INTEGER_FEASIBLE_SOLUTION
        return res
    res.x, res.z = inc_x[:nvar].copy(), inc_z  # ilp.go:111-112
    return res


def gpu_wave_solver():
    """``solve_wave`` backed by the C ABI on the calling process's GPU; the root is uploaded once and cached."""
    from . import capi
    cache = {}

    def solve(c0, A0, b0, bvar, bsign, brhs):
        # keyed on the contents, not on id(): an id can be reused by another array after the first one is collected
        key = (A0.shape, hashlib.sha256(np.ascontiguousarray(A0).tobytes()).digest(),
               hashlib.sha256(np.ascontiguousarray(b0).tobytes() + np.ascontiguousarray(c0).tobytes()).digest())
        if key not in cache:
            for h in cache.values():      # a new root replaces the cached one: release it on the device first
                capi.free_root(h)
            cache.clear()
            cache[key] = capi.upload_root(c0, A0, b0)
        m0, n0 = A0.shape
        w = capi.solve_wave(cache[key], n0, m0, bvar, bsign, brhs)
        tm = capi.last_timing()
        return w.status, w.z, w.x, w.stats[:, 0] + w.stats[:, 1], tm["kernel_ms"]

    def close():
        for h in cache.values():
            capi.free_root(h)
        cache.clear()

    solve.close = close
    return solve
