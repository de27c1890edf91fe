"""gomilp_b200 — B200-native LP-relaxation engine for GoMILP's branch-and-bound.

Thin ctypes binding of ``libgomilp_b200.so`` (the C ABI in ``include/gomilp_b200.h``) plus a host-side
mirror of the reference's operator interface for the relaxation path. All arithmetic happens in the
CUDA kernels of ``csrc/``; there is no CPU fallback: without the built library or without a GPU every
compute call raises.
"""
from . import capi  # noqa: F401
from .capi import (EngineError, LPResult, WaveResult, MilpResult, simplex, simplex_batch, simplex_batch_device,
                   upload_root, free_root, solve_wave, milp_solve, last_timing, set_options, device_count, init,
                   trace_arm, trace_fetch, profile_arm, profile_fetch)
from .status import *  # noqa: F401,F403

__all__ = ["EngineError", "LPResult", "WaveResult", "MilpResult", "simplex", "simplex_batch",
           "simplex_batch_device", "upload_root", "free_root", "solve_wave", "milp_solve", "last_timing",
           "set_options", "device_count", "init", "trace_arm", "trace_fetch", "profile_arm", "profile_fetch"]
