"""CPU oracle — TEST INFRASTRUCTURE ONLY.

ctypes front end of ``oracle/_build/liboracle.so`` (built by ``oracle/Makefile`` from the C++
restatement of Gonum's ``lp.Simplex`` and of GoMILP's branch-and-bound; see ``oracle/README.md``).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package. Nothing under ``gomilp_b200/`` does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile the oracle with ``make -C oracle`` (gcc only; no GPU, no reference needed)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    else:
        # rebuild when a source is newer than the library (cheap make invocation)
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    L.orc_simplex.restype = C.c_int
    L.orc_simplex.argtypes = [_f64p, _f64p, C.c_int64, _f64p, C.c_int64, C.c_int64, C.c_double, C.c_void_p,
                              C.POINTER(C.c_double), _f64p, _i64p, _i64p, C.c_void_p, C.c_int64, C.c_int64]
    L.orc_simplex_batch.restype = C.c_int
    L.orc_simplex_batch.argtypes = [C.c_int64, _f64p, _f64p, _f64p, C.c_int64, C.c_int64, C.c_double, C.c_int,
                                    _i32p, _f64p, _f64p, _i64p, _i64p, C.c_int64]
    L.orc_bnb_solve.restype = C.c_int
    L.orc_bnb_solve.argtypes = [C.c_int64, _f64p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int64, C.c_double, C.c_int64, _f64p,
                                C.POINTER(C.c_double), _i64p, C.c_int64, _i64p, _i64p, _i32p, _i32p, _f64p, _i32p,
                                _i64p, _i32p, _f64p]
    L.orc_convert_to_equalities.restype = None
    L.orc_convert_to_equalities.argtypes = [C.c_int64, _f64p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, _f64p,
                                            _f64p, _f64p, _f64p, _f64p]
    L.orc_maxfun_branch_point.restype = C.c_int
    L.orc_maxfun_branch_point.argtypes = [C.c_int64, _f64p, _u8p]
    L.orc_most_infeasible_branch_point.restype = C.c_int
    L.orc_most_infeasible_branch_point.argtypes = [C.c_int64, _f64p, _u8p]
    L.orc_feasible_for_ip.restype = C.c_int
    L.orc_feasible_for_ip.argtypes = [C.c_int64, _u8p, _f64p]
    L.orc_cond1.restype = C.c_double
    L.orc_cond1.argtypes = [_f64p, C.c_int64, C.c_int64, C.c_int64]
    L.orc_solve_vec.restype = C.c_int
    L.orc_solve_vec.argtypes = [_f64p, C.c_int64, C.c_int, _f64p, _f64p, C.POINTER(C.c_double)]
    L.orc_num_hw_threads.restype = C.c_int
    L.orc_find_linearly_independent.restype = C.c_int
    L.orc_find_linearly_independent.argtypes = [_f64p, C.c_int64, C.c_int64, C.c_int64, _i64p]
    _lib = L
    return L


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


@dataclass
class SimplexOut:
    status: int
    optF: float
    x: np.ndarray | None
    basis: np.ndarray | None
    pivots_phase1: int = 0
    pivots_phase2: int = 0
    bland_calls: int = 0
    lu_factorizations: int = 0
    repair_trials: int = 0
    used_phase1: int = 0
    trace: np.ndarray | None = None  # rows (phase, enter, leave, bland)

    @property
    def pivots(self) -> int:
        return self.pivots_phase1 + self.pivots_phase2


def simplex(c, A, b, tol: float = 0.0, initial_basic=None, trace_cap: int = 0, max_pivots: int = 0) -> SimplexOut:
    """``lp.Simplex(c, A, b, tol, initialBasic)`` (simplex.go:88-91) on the CPU oracle."""
    A = _f64(A)
    c = _f64(c)
    b = _f64(b)
    m, n = A.shape
    assert c.shape == (n,) and b.shape == (m,)
    optF = C.c_double(0.0)
    x = np.zeros(n)
    basis = np.zeros(m, dtype=np.int64)
    stats = np.zeros(8, dtype=np.int64)
    ib = None
    ibp = None
    if initial_basic is not None:
        ib = np.ascontiguousarray(initial_basic, dtype=np.int64)
        ibp = ib.ctypes.data_as(C.c_void_p)
    tr = np.zeros((max(trace_cap, 1), 4), dtype=np.int32)
    trp = tr.ctypes.data_as(C.c_void_p) if trace_cap > 0 else None
    st = lib().orc_simplex(c, A, n, b, m, n, float(tol), ibp, C.byref(optF), x, basis, stats, trp, trace_cap,
                           max_pivots)
    has_x = bool(stats[7])
    return SimplexOut(st, optF.value, x if has_x else None, basis if (has_x and basis[0] >= 0) else None,
                      int(stats[0]), int(stats[1]), int(stats[2]), int(stats[3]), int(stats[4]), int(stats[5]),
                      tr[: min(trace_cap, int(stats[6]))].copy() if trace_cap > 0 else None)


def simplex_batch(c, A, b, tol: float = 0.0, threads: int = 1, max_pivots: int = 0):
    """Independent same-shape LPs: A [batch,m,n], c [batch,n], b [batch,m]. Returns dict of arrays."""
    A = _f64(A)
    c = _f64(c)
    b = _f64(b)
    batch, m, n = A.shape
    status = np.zeros(batch, dtype=np.int32)
    optF = np.zeros(batch)
    x = np.zeros((batch, n))
    basis = np.zeros((batch, m), dtype=np.int64)
    pivots = np.zeros(batch, dtype=np.int64)
    lib().orc_simplex_batch(batch, c, A, b, m, n, float(tol), int(threads), status, optF, x, basis, pivots,
                            max_pivots)
    return {"status": status, "optF": optF, "x": x, "basis": basis, "pivots": pivots}


@dataclass
class BnbOut:
    status: int
    lp_status: int
    x: np.ndarray | None
    z: float
    nodes: int
    pivots: int
    log: dict = field(default_factory=dict)


def bnb_solve(c, A=None, b=None, G=None, h=None, integrality=None, heuristic: int = 0, mode: int = 0,
              node_limit: int = 0, time_limit_s: float = 0.0, max_pivots_per_lp: int = 0,
              log_cap: int = 1 << 16) -> BnbOut:
    """``milpProblem.solve`` (ilp.go:75-116) replayed serially in 1-worker FIFO order."""
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
    z = C.c_double(0.0)
    sc = np.zeros(8, dtype=np.int64)
    cap = max(log_cap, 1)
    lid = np.zeros(cap, dtype=np.int64)
    lpar = np.zeros(cap, dtype=np.int64)
    ldep = np.zeros(cap, dtype=np.int32)
    lst = np.zeros(cap, dtype=np.int32)
    lz = np.zeros(cap)
    ldec = np.zeros(cap, dtype=np.int32)
    lpiv = np.zeros(cap, dtype=np.int64)
    lbv = np.zeros(cap, dtype=np.int32)
    lbf = np.zeros(cap)

    def vp(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    lib().orc_bnb_solve(nvar, c, meq, vp(Aa), vp(ba), nineq, vp(Ga), vp(ha), integ, heuristic, mode, node_limit,
                        float(time_limit_s), max_pivots_per_lp, x, C.byref(z), sc, log_cap, lid, lpar, ldep, lst, lz,
                        ldec, lpiv, lbv, lbf)
    k = min(log_cap, int(sc[5]))
    xl = int(sc[2])
    return BnbOut(int(sc[0]), int(sc[1]), x[:xl].copy() if xl else None, z.value, int(sc[3]), int(sc[4]),
                  {"id": lid[:k], "parent": lpar[:k], "depth": ldep[:k], "lp_status": lst[:k], "z": lz[:k],
                   "decision": ldec[:k], "pivots": lpiv[:k], "branch_var": lbv[:k], "branch_floor": lbf[:k],
                   "total": int(sc[5])})


def convert_to_equalities(c, A, b, G, h):
    c = _f64(c)
    nvar = c.shape[0]
    meq = 0 if A is None else np.asarray(A).shape[0]
    G = _f64(G)
    nineq = G.shape[0]
    Aa = _f64(A) if meq else None
    ba = _f64(b) if meq else None
    cN = np.zeros(nvar + nineq)
    aN = np.zeros((meq + nineq, nvar + nineq))
    bN = np.zeros(meq + nineq)
    lib().orc_convert_to_equalities(nvar, c, meq, None if Aa is None else Aa.ctypes.data_as(C.c_void_p),
                                    None if ba is None else ba.ctypes.data_as(C.c_void_p), nineq, G, _f64(h), cN, aN,
                                    bN)
    return cN, aN, bN


def maxfun_branch_point(c, integ) -> int:
    return lib().orc_maxfun_branch_point(len(c), _f64(c), np.ascontiguousarray(integ, dtype=np.uint8))


def most_infeasible_branch_point(c, integ) -> int:
    return lib().orc_most_infeasible_branch_point(len(c), _f64(c), np.ascontiguousarray(integ, dtype=np.uint8))


def feasible_for_ip(integ, x) -> bool:
    return bool(lib().orc_feasible_for_ip(len(x), np.ascontiguousarray(integ, dtype=np.uint8), _f64(x)))


def cond1(a) -> float:
    a = _f64(a)
    m, k = a.shape
    return lib().orc_cond1(a, k, m, k)


def solve_vec(a, b, transpose: bool = False):
    a = _f64(a)
    n = a.shape[0]
    x = np.zeros(n)
    cond = C.c_double(0.0)
    rc = lib().orc_solve_vec(a, n, int(transpose), _f64(b), x, C.byref(cond))
    return rc, x, cond.value


def initial_basis(A):
    """findLinearlyIndependent (simplex.go:611-637): accepted column indices in acceptance order, None if < m."""
    A = _f64(A)
    m, n = A.shape
    idx = np.zeros(m, dtype=np.int64)
    k = lib().orc_find_linearly_independent(A, n, m, n, idx)
    return idx if k == m else None


def num_hw_threads() -> int:
    return lib().orc_num_hw_threads()
