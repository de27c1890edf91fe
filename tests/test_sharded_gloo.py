"""CPU, world_size 2, gloo: the multi-rank branch-and-bound scheduler (gomilp_b200/sharded.py) — FIFO block
partition, the per-wave record exchange, incumbent broadcast — replays the oracle's 1-worker decisions.
The LP solves of each rank's block are done by the oracle here (no GPU in this container); on the GPU box
the same scheduler runs over NCCL with gm_solve_wave (tests/test_gpu_parity.py::test_sharded_driver_single_rank)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def oracle_wave_solver():
    import oracle

    def solve(c0, A0, b0, bvar, bsign, brhs):
        m0, n0 = A0.shape
        k, L = bvar.shape
        st = np.zeros(k, dtype=np.int32)
        z = np.zeros(k)
        x = np.zeros((k, n0))
        piv = np.zeros(k, dtype=np.int64)
        for i in range(k):
            A = np.zeros((m0 + L, n0 + L))
            A[:m0, :n0] = A0
            for l in range(L):
                A[m0 + l, bvar[i, l]] = bsign[i, l]
                A[m0 + l, n0 + l] = 1.0
            r = oracle.simplex(np.concatenate([c0, np.zeros(L)]), A, np.concatenate([b0, brhs[i]]))
            st[i], z[i], piv[i] = r.status, r.optF, r.pivots
            if r.x is not None:
                x[i] = r.x[:n0]
        return st, z, x, piv, 0.0

    return solve


def _worker(rank, world, port, problems, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gomilp_b200.sharded import milp_solve_sharded
    res = []
    for p in problems:
        r = milp_solve_sharded(p["c"], p["A"], p["b"], p["G"], p["h"], p["integrality"], solve_wave=oracle_wave_solver(),
                               mode=p["mode"], heuristic=p["heuristic"], node_limit=p["node_limit"])
        res.append((r.status, r.z, None if r.x is None else r.x.tolist(), r.nodes, r.waves,
                    [(d[0], d[1], d[5], d[6], d[7]) for d in r.decisions], r.exchange_bytes))
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


def test_fifo_blocks_are_a_contiguous_partition():
    from gomilp_b200.sharded import fifo_block
    for count in (0, 1, 2, 7, 8, 9, 1000):
        for world in (1, 2, 3, 8):
            blocks = [fifo_block(count, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == count
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            sizes = [h - l for l, h in blocks]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_bnb_replays_the_one_worker_reference_order():
    import oracle
    from problems import random_milp, reference_pins
    problems = []
    for case in reference_pins()["milp"]:
        problems.append(dict(c=case["c"], A=case["A"], b=case["b"], G=case["G"], h=case["h"],
                             integrality=case["integrality"], mode=0, heuristic=0, node_limit=40))
    rng = np.random.default_rng(4)
    for _ in range(3):
        p = random_milp(rng, 4, 2)
        problems.append(dict(c=p["c"], A=None, b=None, G=p["G"], h=p["h"], integrality=p["integrality"], mode=1,
                             heuristic=1, node_limit=40))
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(2, port, problems, out), nprocs=2, join=True)
    assert out[0] == out[1]  # both ranks took identical decisions
    for p, got in zip(problems, out[0]):
        o = oracle.bnb_solve(p["c"], p["A"], p["b"], p["G"], p["h"], p["integrality"], heuristic=p["heuristic"],
                             mode=p["mode"], node_limit=p["node_limit"])
        status, z, x, nodes, waves, decisions, xbytes = got
        assert status == o.status and nodes == o.nodes
        want = list(zip(o.log["id"].tolist(), o.log["parent"].tolist(), o.log["decision"].tolist(),
                        o.log["branch_var"].tolist(), o.log["branch_floor"].tolist()))
        assert [tuple(d) for d in decisions] == want
        if o.x is not None:
            assert z == o.z and x == o.x.tolist()  # same arithmetic (the oracle) => bit-identical replay
        assert xbytes > 0 or nodes <= 1
