"""ctypes binding of libgomilp_b200.so — one Python function per C-ABI entry point (include/gomilp_b200.h).

The names and argument meaning follow the reference interface each entry point replaces:
``simplex`` = ``lp.Simplex(c, A, b, tol, initialBasic)`` (vendor/gonum.org/v1/gonum/optimize/convex/lp/simplex.go:88),
``solve_wave`` = ``subProblem.solve`` over a FIFO wave (subproblem.go:141-187), ``milp_solve`` =
``milpProblem.solve`` (ilp.go:75-116). Nothing here computes: the library is required and a missing GPU
surfaces as ``EngineError(GM_ERR_NO_DEVICE)``.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import status as S

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libgomilp_b200.so")
_lib = None


class EngineError(RuntimeError):
    def __init__(self, code: int, msg: str = ""):
        self.code = code
        super().__init__(f"gomilp_b200 engine error {code} ({S.STATUS_NAMES.get(code, '?')}) {msg}")


class gm_timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("kernel_ms", C.c_double), ("d2h_ms", C.c_double), ("lps", C.c_int64),
                ("launches", C.c_int64), ("smem_bytes", C.c_int64), ("tier", C.c_int32), ("grid", C.c_int32),
                ("block", C.c_int32)]


class gm_options(C.Structure):
    _fields_ = [("max_pivots", C.c_int32), ("refactor_period", C.c_int32), ("force_tier", C.c_int32),
                ("reserved", C.c_int32), ("coop_group", C.c_int32), ("reserved2", C.c_int32), ("robust", C.c_int32),
                ("reserved3", C.c_int32)]


class gm_milp_result(C.Structure):
    _fields_ = [("status", C.c_int32), ("lp_status", C.c_int32), ("z", C.c_double), ("x_len", C.c_int64),
                ("nodes", C.c_int64), ("waves", C.c_int64), ("pivots", C.c_int64), ("device_ms", C.c_double)]


DECISION_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_double, C.c_int32,
                          C.c_int32, C.c_double)
WAVE_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_double)

EXPORTS = ["gm_device_count", "gm_init", "gm_shutdown", "gm_last_error", "gm_last_timing", "gm_set_options",
           "gm_simplex", "gm_simplex_batch", "gm_simplex_batch_device", "gm_upload_root", "gm_free_root",
           "gm_solve_wave", "gm_solve_wave_warm", "gm_milp_solve", "gm_trace_arm", "gm_trace_fetch",
           "gm_milp_solve_device", "gm_thread_robust", "gm_microbench_smem_gbs", "gm_profile_arm", "gm_profile_fetch", "gm_comm_unique_id", "gm_comm_init", "gm_comm_destroy"]


def lib():
    """Loads the in-tree library; raises if it has not been built (there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -m gomilp_b200.build` "
                          "(the engine has no CPU or PyTorch fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, f64 = C.c_void_p, C.c_int64, C.c_int32, C.c_double
    L.gm_device_count.restype = C.c_int
    L.gm_init.argtypes = [C.c_int]
    L.gm_last_error.restype = C.c_char_p
    L.gm_last_timing.argtypes = [C.POINTER(gm_timing)]
    L.gm_set_options.argtypes = [C.POINTER(gm_options)]
    L.gm_simplex.argtypes = [vp, vp, i64, vp, i64, i64, f64, vp, C.POINTER(f64), vp, vp, C.POINTER(i64)]
    L.gm_simplex_batch.argtypes = [i64, vp, vp, vp, i64, i64, f64, vp, vp, vp, vp, vp]
    L.gm_simplex_batch_device.argtypes = [i64, vp, vp, vp, i64, i64, f64, vp, vp, vp, vp, vp, vp]
    L.gm_upload_root.argtypes = [vp, vp, i64, vp, i64, i64, C.POINTER(i64)]
    L.gm_free_root.argtypes = [i64]
    L.gm_solve_wave.argtypes = [i64, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp]
    L.gm_solve_wave_warm.argtypes = [i64, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.gm_milp_solve.argtypes = [i64, vp, i64, vp, vp, i64, vp, vp, vp, i32, i32, i64, f64, vp,
                                C.POINTER(gm_milp_result), DECISION_CB, WAVE_CB, vp]
    L.gm_milp_solve_device.argtypes = L.gm_milp_solve.argtypes
    L.gm_thread_robust.argtypes = [C.c_int]
    L.gm_microbench_smem_gbs.argtypes = [C.POINTER(f64)]
    L.gm_comm_unique_id.argtypes = [vp]
    L.gm_comm_init.argtypes = [i32, i32, vp]
    L.gm_trace_arm.argtypes = [i64, i64]
    L.gm_trace_fetch.argtypes = [vp, i64]
    for name in EXPORTS:
        if name != "gm_last_error":
            getattr(L, name).restype = C.c_int
    L.gm_trace_fetch.restype = C.c_int64
    L.gm_profile_fetch.argtypes = [vp, i64]
    L.gm_profile_fetch.restype = C.c_int64
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _check(rc: int):
    if rc != S.GM_OK:
        raise EngineError(rc, lib().gm_last_error().decode() if rc == S.GM_ERR_CUDA else "")


def device_count() -> int:
    return lib().gm_device_count()


def init(device: int = 0):
    _check(lib().gm_init(device))


def set_options(max_pivots: int = 0, refactor_period: int = 0, force_tier: int = 0, no_tma_ring: bool = False,
                coop_group: int = 0, no_streamed_batch: bool = False, robust: bool = False):
    o = gm_options(max_pivots, refactor_period, force_tier, 1 if no_tma_ring else 0, coop_group,
                   1 if no_streamed_batch else 0, 1 if robust else 0, 0)
    _check(lib().gm_set_options(C.byref(o)))


def trace_arm(lp_index: int = 0, cap: int = 256):
    """The next host-buffer compute call on this thread records the first `cap` pivots of LP `lp_index`."""
    _check(lib().gm_trace_arm(lp_index, cap))


def trace_fetch(cap: int = 256) -> np.ndarray:
    """Rows (phase, entering variable, leaving variable, bland) of the armed call; unused rows trimmed."""
    rows = np.full((cap, 4), -1, dtype=np.int32)
    k = lib().gm_trace_fetch(_p(rows), cap)
    rows = rows[:k]
    return rows[rows[:, 0] >= 0]


def last_timing() -> dict:
    t = gm_timing()
    lib().gm_last_timing(C.byref(t))
    return {k: getattr(t, k) for k, _ in gm_timing._fields_}


@dataclass
class LPResult:
    """Return triple of lp.Simplex plus what the internal simplex() also returns (simplex.go:93,301)."""
    status: int
    optF: float
    x: np.ndarray | None
    basis: np.ndarray | None
    pivots: int


def simplex(c, A, b, tol: float = 0.0, initial_basic=None) -> LPResult:
    """lp.Simplex(c, A, b, tol, initialBasic) — simplex.go:88-91 — on the GPU engine."""
    A = _f64(A)
    c = _f64(c)
    b = _f64(b)
    m, n = A.shape
    if c.shape != (n,) or b.shape != (m,):
        raise ValueError("lp: c / b vector incorrect length")  # the reference panics, simplex.go:387-398
    optF = C.c_double(0.0)
    x = np.zeros(n)
    basis = np.full(m, -1, dtype=np.int64)
    piv = C.c_int64(0)
    ib = None if initial_basic is None else np.ascontiguousarray(initial_basic, dtype=np.int64)
    if ib is not None and ib.shape != (m,):
        raise ValueError("lp: initialBasic incorrect length")
    st = lib().gm_simplex(_p(c), _p(A), n, _p(b), m, n, float(tol), _p(ib), C.byref(optF), _p(x), _p(basis),
                          C.byref(piv))
    if st >= S.GM_ERR_BAD_SHAPE and st not in (S.GM_ERR_ITERATION_LIMIT,):
        _check(st)
    has_x = st not in (S.GM_ERR_INFEASIBLE, S.GM_ERR_UNBOUNDED, S.GM_ERR_SINGULAR, S.GM_ERR_ZERO_ROW,
                       S.GM_ERR_ZERO_COLUMN, S.GM_PANIC_INITIAL_BASIC) and st < S.GM_ERR_PHASE1_WRAPPED
    return LPResult(st, optF.value, x if has_x else None, basis if (has_x and basis[0] >= 0) else None, piv.value)


def simplex_batch(c, A, b, tol: float = 0.0, want_basis: bool = True, want_stats: bool = True) -> dict:
    """`count` independent lp.Simplex calls of one shape in one launch; HOST arrays in and out."""
    A = _f64(A)
    c = _f64(c)
    b = _f64(b)
    count, m, n = A.shape
    assert c.shape == (count, n) and b.shape == (count, m)
    status = np.zeros(count, dtype=np.int32)
    optF = np.zeros(count)
    x = np.zeros((count, n))
    basis = np.zeros((count, m), dtype=np.int64) if want_basis else None
    stats = np.zeros((count, 8), dtype=np.int32) if want_stats else None
    _check(lib().gm_simplex_batch(count, _p(c), _p(A), _p(b), m, n, float(tol), _p(status), _p(optF), _p(x),
                                  _p(basis), _p(stats)))
    out = {"status": status, "optF": optF, "x": x, "basis": basis, "stats": stats}
    if stats is not None:
        out["pivots"] = stats[:, 0].astype(np.int64) + stats[:, 1]
    return out


def simplex_batch_device(count, d_c, d_A, d_b, m, n, tol, d_status, d_optF, d_x, d_basis=0, d_stats=0, stream=0):
    """Device-pointer form (ints = CUDA device addresses, e.g. torch ``tensor.data_ptr()``), asynchronous."""
    _check(lib().gm_simplex_batch_device(count, d_c, d_A, d_b, m, n, float(tol), d_status, d_optF, d_x,
                                         d_basis or None, d_stats or None, stream or None))


def upload_root(c0, A0, b0) -> int:
    """Device copy of the root standard form every node shares (toInitialSubproblem, ilp.go:43-71)."""
    A0 = _f64(A0)
    c0 = _f64(c0)
    b0 = _f64(b0)
    m0, n0 = A0.shape
    h = C.c_int64(0)
    _check(lib().gm_upload_root(_p(c0), _p(A0), n0, _p(b0), m0, n0, C.byref(h)))
    return h.value


def free_root(h: int):
    _check(lib().gm_free_root(h))


@dataclass
class WaveResult:
    status: np.ndarray
    z: np.ndarray
    x: np.ndarray
    basis: np.ndarray
    stats: np.ndarray


def solve_wave(root: int, n0: int, m0: int, bvar, bsign, brhs, parent=None, warm: bool = False) -> WaveResult:
    """subProblem.solve (subproblem.go:141-187) for every node of a wave; node k has L branch rows.

    warm=True goes through gm_solve_wave_warm: the engine keeps this wave's bases / inverses in HBM, and
    ``parent[k]`` (index into the previous warm wave on this root, -1 = cold) lets node k start from there."""
    bvar = np.ascontiguousarray(bvar, dtype=np.int32)
    nodes, L = bvar.shape
    bsign = _f64(bsign).reshape(nodes, L)
    brhs = _f64(brhs).reshape(nodes, L)
    status = np.zeros(nodes, dtype=np.int32)
    z = np.zeros(nodes)
    x = np.zeros((nodes, n0))
    basis = np.zeros((nodes, m0 + L), dtype=np.int64)
    stats = np.zeros((nodes, 8), dtype=np.int32)
    if warm:
        par = None if parent is None else np.ascontiguousarray(parent, dtype=np.int32)
        _check(lib().gm_solve_wave_warm(root, nodes, L, _p(bvar), _p(bsign), _p(brhs), _p(par), _p(status), _p(z),
                                        _p(x), _p(basis), _p(stats)))
    else:
        _check(lib().gm_solve_wave(root, nodes, L, _p(bvar), _p(bsign), _p(brhs), _p(status), _p(z), _p(x),
                                   _p(basis), _p(stats)))
    return WaveResult(status, z, x, basis, stats)


@dataclass
class MilpResult:
    status: int
    lp_status: int
    x: np.ndarray | None
    z: float
    nodes: int
    waves: int
    pivots: int
    device_ms: float
    log: list = field(default_factory=list)    # (id, parent, depth, lp_status, z, decision, branch_var, branch_floor)
    waves_log: list = field(default_factory=list)  # (wave, nodes, pivots, kernel_ms)


def milp_solve(c, A=None, b=None, G=None, h=None, integrality=None, heuristic: int = 0, mode: int = 0,
               node_limit: int = 0, time_limit_s: float = 0.0, keep_log: bool = True) -> MilpResult:
    """milpProblem.solve (ilp.go:75-116): wavefront branch-and-bound, every relaxation on the GPU."""
    c = _f64(c)
    nvar = c.shape[0]
    meq = 0 if A is None else np.asarray(A).shape[0]
    nineq = 0 if G is None else np.asarray(G).shape[0]
    Aa = _f64(A).reshape(meq, nvar) if meq else None
    ba = _f64(b) if meq else None
    Ga = _f64(G).reshape(nineq, nvar) if nineq else None
    ha = _f64(h) if nineq else None
    integ = np.ascontiguousarray(integrality, dtype=np.uint8)
    assert integ.shape == (nvar,)
    x = np.zeros(nvar + nineq + 1)
    res = gm_milp_result()
    log, wlog = [], []

    def on_dec(_u, id_, parent, depth, lp_status, z, decision, bvar, bfloor):
        log.append((id_, parent, depth, lp_status, z, decision, bvar, bfloor))

    def on_wave(_u, wave, nodes, pivots, ms):
        wlog.append((wave, nodes, pivots, ms))

    cb1 = DECISION_CB(on_dec) if keep_log else C.cast(None, DECISION_CB)
    cb2 = WAVE_CB(on_wave) if keep_log else C.cast(None, WAVE_CB)
    rc = lib().gm_milp_solve(nvar, _p(c), meq, _p(Aa), _p(ba), nineq, _p(Ga), _p(ha), _p(integ), heuristic, mode,
                             node_limit, float(time_limit_s), _p(x), C.byref(res), cb1, cb2, None)
    if rc != S.GM_OK:
        _check(rc)
    xl = int(res.x_len)
    return MilpResult(res.status, res.lp_status, x[:xl].copy() if xl else None, res.z, res.nodes, res.waves,
                      res.pivots, res.device_ms, log, wlog)


def comm_unique_id() -> bytes:
    """128-byte NCCL id created by rank 0; hand it to the other ranks (torch.distributed, a file, a channel)."""
    buf = C.create_string_buffer(128)
    _check(lib().gm_comm_unique_id(buf))
    return buf.raw


def comm_init(rank: int, world: int, uid: bytes):
    """Joins the calling thread (bound to its device by init()) to the communicator of `world` ranks."""
    assert len(uid) == 128
    _check(lib().gm_comm_init(rank, world, C.create_string_buffer(uid, 128)))


def comm_destroy():
    lib().gm_comm_destroy()


def microbench_smem_gbs() -> float:
    """Measured shared-memory bandwidth of the current device, GB/s."""
    v = C.c_double(0.0)
    _check(lib().gm_microbench_smem_gbs(C.byref(v)))
    return v.value


def profile_arm():
    """Cooperative tier: the next host-buffer compute call reports per-LP leader clock cycles (see gm_profile_arm)."""
    _check(lib().gm_profile_arm())


def profile_fetch(lps: int = 1) -> np.ndarray:
    out = np.zeros((lps, 16), dtype=np.int64)
    k = lib().gm_profile_fetch(_p(out), lps)
    return out[:k]
