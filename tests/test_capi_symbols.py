"""CPU: the C-ABI library builds, loads, and exports every symbol include/gomilp_b200.h declares; without a
GPU the compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import gomilp_b200 as gm
from gomilp_b200 import build, capi, status as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_is_built_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    hdr = open(os.path.join(ROOT, "include", "gomilp_b200.h")).read()
    declared = set(re.findall(r"\b(gm_[a-z0-9_]+)\s*\(", hdr)) - {"gm_decision_cb", "gm_wave_cb"}
    assert declared == set(capi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_product_package_never_touches_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gomilp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "liboracle" not in src and "oracle/" not in src, f


def test_no_cpu_fallback_without_a_device():
    if gm.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(gm.EngineError) as e:
        gm.simplex_batch(np.ones((1, 3)), np.ones((1, 2, 3)), np.ones((1, 2)))
    assert e.value.code == S.GM_ERR_NO_DEVICE
    r = gm.simplex([1.0, 1.0, 1.0], np.ones((2, 3)), [1.0, 1.0]) if False else None
    assert r is None
