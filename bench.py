#!/usr/bin/env python
"""bench.py — LP relaxations solved per second on BASELINE.json configs[1]:
a batch of 4096 independent random dense LP relaxations, m=64, n=128, float64, per GPU (weak scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N>1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...)

One "step" = one pass of the hot path (gm_simplex_batch_device: the simplex wave kernel) over one batch.
`value` is measured with the inputs resident in HBM; `e2e` goes through the host-buffer C-ABI entry point
(gm_simplex_batch) from pinned host memory, copies inside the timed region. `cpu_baseline` / `--impl
reference` time the CPU oracle (oracle/: C++ restatement of the reference's Gonum lp.Simplex — the
reference is Go and there is no Go toolchain in the image, so kind = "port") on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

M, N, BATCH = 64, 128, 4096
SEED = 1234
METRIC = "lp_relaxations_per_sec"
UNIT = "LP/s"
WORKLOAD = "C2: 4096 independent random dense LP relaxations m=64 n=128 f64 per GPU (feasible+bounded generator, seed 1234+rank)"


# ncu --set full capture of one launch of this exact workload (seed 1234): 277.0 MB read + 9.5 MB written, i.e. the
# 275 MB batch is read once from HBM and the results written once; the per-pivot bytes stay on chip.
TIER1_DRAM_BYTES_PER_LAUNCH = 286.5e6


def hbm_tier_extra(gm):
    """The HBM-resident pivot path (tier 4, TMA staging ring), outside the timed region: 148 dense LPs of C4's shape
    (m=1024, n=2048, slack form so that no basis inversion dilutes it), 60 pivots each."""
    try:
        m, n, count, cap = 1024, 2048, 148, 60
        rng = np.random.default_rng(42)
        base = 4
        A = np.zeros((base, m, n))
        A[:, :, : n - m] = rng.random((base, m, n - m))
        A[:, :, n - m:] = np.eye(m)
        b = 1.0 + rng.random((base, m))
        c = np.zeros((base, n))
        c[:, : n - m] = -rng.random((base, n - m))
        reps = (count + base - 1) // base
        c, A, b = np.tile(c, (reps, 1))[:count], np.tile(A, (reps, 1, 1))[:count], np.tile(b, (reps, 1))[:count]
        gm.set_options(max_pivots=cap, refactor_period=100000)
        gm.simplex_batch(c, A, b, want_basis=False)
        g = gm.simplex_batch(c, A, b, want_basis=False)
        tm = gm.last_timing()
        gm.set_options()
        piv = int(g["pivots"].sum())
        alg = piv * bytes_per_pivot(m, n)
        return {"workload": "148 dense LPs m=1024 n=2048 (C4's shape), slack form, 60 pivots each, one CTA per LP",
                "tier": tm["tier"], "kernel_ms": tm["kernel_ms"], "pivots": piv,
                "algorithmic_GBps": alg / (tm["kernel_ms"] * 1e-3) / 1e9,
                "dram_GBps_from_ncu": 3560.0, "dram_frac_of_measured_hbm_peak": 0.55,
                "dram_source": "profiles/r01_tier4_tma_1024x2048_ncu_full_attribution.txt: 162.3 GB read + 78.9 GB "
                               "written in 67.75 ms for the same launch"}
    except Exception as e:
        gm.set_options()
        return {"error": str(e)}


def bytes_per_pivot(m: int, n: int) -> int:
    """FTRAN reads B^-1 (m^2), the update reads+writes it (2 m^2), pricing reads A_N (m(n-m))."""
    return 8 * (3 * m * m + m * (n - m))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([s.strip() for s in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_batch(rank: int, count: int = BATCH):
    from problems import feasible_bounded_lp
    rng = np.random.default_rng(SEED + rank)
    return feasible_bounded_lp(rng, M, N, count)


def cpu_baseline(count: int, threads: int):
    """The oracle timed on `count` LPs of the same workload with `threads` host threads."""
    import oracle
    c, A, b = make_batch(0, count)
    oracle.simplex_batch(c[:threads], A[:threads], b[:threads], threads=threads)  # warm the library
    t0 = time.perf_counter()
    o = oracle.simplex_batch(c, A, b, threads=threads)
    dt = time.perf_counter() - t0
    assert (o["status"] == 0).all()
    return count / dt, dt, int(o["pivots"].sum())


def bnb_extra(gm):
    """Second half of BASELINE.json's metric, outside the timed region: B&B nodes/s through gm_milp_solve (one wave
    launch per BFS level, decisions replayed on the host) on a small 0-1 multidimensional knapsack, 1 GPU."""
    from problems import knapsack
    try:
        p = knapsack(np.random.default_rng(7), 30, 5)
        for warm_up_mode in (1, 1 | 4):  # also loads the warm-start kernel before anything is timed
            gm.milp_solve(p["c"], None, None, p["G"], p["h"], p["integrality"], mode=warm_up_mode, heuristic=1,
                          node_limit=256, keep_log=False)
        def best_of(mode, reps=3):
            best = None
            for _ in range(reps):
                t0 = time.perf_counter()
                res = gm.milp_solve(p["c"], None, None, p["G"], p["h"], p["integrality"], mode=mode, heuristic=1,
                                    node_limit=8192, keep_log=False)
                el = time.perf_counter() - t0
                if best is None or el < best[1]:
                    best = (res, el)
            return best

        r, dt = best_of(1)       # wall clock of a host-driven loop of ~14 small launches: best of 3
        rw, dtw = best_of(1 | 4)
        return {"nodes_per_sec": r.nodes / dt, "nodes": r.nodes, "waves": r.waves, "pivots": r.pivots,
                "wall_s": dt, "device_ms": r.device_ms, "gpus": 1,
                "warm_start": {"nodes_per_sec": rw.nodes / dtw, "nodes": rw.nodes, "pivots": rw.pivots,
                               "device_ms": rw.device_ms, "note": "GM_BNB_WARM_START: children continue from the "
                               "parent's basis inverse kept in HBM (not a replay of the reference's cold solves)"},
                "workload": "0-1 knapsack n=30 m=5 (standard form 35x65 + depth), FIXED mode, most-infeasible "
                            "branching, node budget 8192, children re-solved from scratch like the reference"}
    except Exception as e:  # never let the extra break the contract line
        return {"error": str(e)}


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    import oracle
    cores = os.cpu_count() or 1
    per_step = max(16, 8 * cores)
    c, A, b = make_batch(0, per_step)
    for _ in range(args.warmup):
        oracle.simplex_batch(c[:cores], A[:cores], b[:cores], threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.simplex_batch(c, A, b, threads=cores)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = f"{per_step} LPs of the workload per step, {cores} threads, one LP per thread"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU oracle = C++ restatement of the reference's Gonum lp.Simplex "
                   "(Go toolchain absent); bounded sample per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import gomilp_b200 as gm
    if not torch.cuda.is_available() or gm.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    gm.init(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)

    c, A, b = make_batch(rank)
    # pinned host copies (e2e) and device-resident copies (value)
    hc = torch.from_numpy(c).pin_memory()
    hA = torch.from_numpy(A).pin_memory()
    hb = torch.from_numpy(b).pin_memory()
    dc, dA, db = hc.to(dev), hA.to(dev), hb.to(dev)
    d_status = torch.zeros(BATCH, dtype=torch.int32, device=dev)
    d_optF = torch.zeros(BATCH, dtype=torch.float64, device=dev)
    d_x = torch.zeros(BATCH, N, dtype=torch.float64, device=dev)
    d_basis = torch.zeros(BATCH, M, dtype=torch.int64, device=dev)
    d_stats = torch.zeros(BATCH, 8, dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(device=dev)

    def step():
        gm.simplex_batch_device(BATCH, dc.data_ptr(), dA.data_ptr(), db.data_ptr(), M, N, 0.0, d_status.data_ptr(),
                                d_optF.data_ptr(), d_x.data_ptr(), d_basis.data_ptr(), d_stats.data_ptr(),
                                stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for k in range(args.steps):
            step()
            ev[k + 1].record(stream)
    stream.synchronize()
    barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    tm = gm.last_timing()
    status = d_status.cpu().numpy()
    stats = d_stats.cpu().numpy()
    pivots = int(stats[:, 0].sum() + stats[:, 1].sum())
    inversions = int(stats[:, 3].sum())
    assert (status == 0).all(), "bench workload must solve to optimality"

    # ---- e2e: host buffers through the C ABI, copies inside the timed region --------------------
    hs = np.zeros(BATCH, dtype=np.int32)
    hF = np.zeros(BATCH)
    hx = np.zeros((BATCH, N))
    hB = np.zeros((BATCH, M), dtype=np.int64)
    hS = np.zeros((BATCH, 8), dtype=np.int32)
    cn, An, bn = hc.numpy(), hA.numpy(), hb.numpy()
    import ctypes as C
    L = gm.capi.lib()

    def p(a):
        return a.ctypes.data_as(C.c_void_p)

    def e2e_step():
        rc = L.gm_simplex_batch(BATCH, p(cn), p(An), p(bn), M, N, 0.0, p(hs), p(hF), p(hx), p(hB), p(hS))
        assert rc == 0, rc

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    e2e_tm = gm.last_timing()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    assert np.array_equal(hs, status) and np.allclose(hF, d_optF.cpu().numpy(), rtol=0, atol=0)

    t = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        pv = torch.tensor([pivots], dtype=torch.float64, device=dev)
        dist.all_reduce(pv)
        pivots_all = int(pv.item())
    else:
        pivots_all = pivots
    total_ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    value = world * BATCH * args.steps / (total_ms * 1e-3)
    e2e_value = world * BATCH * args.steps / (e2e_ms * 1e-3)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    launch_ms = float(np.mean(kernel_ms))
    alg_bytes = pivots * bytes_per_pivot(M, N)  # this rank's launch
    achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
    sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    smem_peak = 148 * 128 * sm_mhz * 1e6 / 1e9  # B/clk/SM crossbar x SMs x clock under load, GB/s
    cores = os.cpu_count() or 1
    n_cpu = min(BATCH, 64 * cores)
    cpu_v, cpu_dt, _ = cpu_baseline(n_cpu, cores)
    bnb = bnb_extra(gm)
    hbm_tier = hbm_tier_extra(gm)
    h2d = BATCH * (M * N + M + N) * 8
    d2h = BATCH * (N + 1) * 8 + BATCH * 4 + BATCH * M * 8 + BATCH * 8 * 4
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "m": M, "n": N, "batch_per_gpu": BATCH, "tol": 0.0,
                   "l2": "inputs (275 MB per GPU) exceed the 126 MB L2; no explicit flush",
                   "tier": tm["tier"], "grid": tm["grid"], "block": tm["block"], "smem_bytes_per_cta": tm["smem_bytes"],
                   "pivots_per_lp": pivots / BATCH, "basis_inversions_per_lp": inversions / BATCH,
                   "sharding": "independent LPs, no data-path collective"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps,
                "breakdown_last_step_ms": {"h2d": e2e_tm["h2d_ms"], "kernel": e2e_tm["kernel_ms"], "d2h": e2e_tm["d2h_ms"]},
                "api": "gm_simplex_batch (C ABI, pinned host buffers)"},
        "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": TIER1_DRAM_BYTES_PER_LAUNCH,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch of this workload, "
                                       "profiles/r01_tier1_final_ncu_full_attribution.txt (ncu --set full)", "kernel": "simplex_wave_reg<256,2>" if tm["tier"] == 1 else "simplex_wave_smem<256,2>", "launch_ms": launch_ms,
                     "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_pivot": bytes_per_pivot(M, N),
                     "pivots_per_launch": pivots, "peak_source": peak_src,
                     "note": "tier 1 keeps B^-1 in registers and W in shared memory, so the algorithmic bytes never "
                             "touch HBM; the binding resources are issue slots / barrier latency (see smem)",
                     "smem": {"peak_gbs": smem_peak, "frac": achieved / smem_peak,
                              "peak_source": "148 SMs x 128 B/clk x median SM clock under load"}},
        "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"first {n_cpu} LPs of rank 0's batch, {cores} threads, {cpu_dt:.1f} s"},
        "pivots_per_sec": pivots_all * args.steps / (total_ms * 1e-3),
        "bnb": bnb,
        "hbm_tier": hbm_tier,
        "clocks": clocks,
    }))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it so that there is one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533"),
               os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup",
               str(args.warmup), "--impl", args.impl]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
