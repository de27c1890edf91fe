"""Builds libgomilp_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT_DIR = os.path.join(_HERE, "_build")
LIB_PATH = os.path.join(OUT_DIR, "libgomilp_b200.so")
SOURCES = ["engine.cu", "kernels_reg.cu", "kernels_generic.cu", "kernels_coop.cu", "bnb_device.cu", "microbench.cu", "bnb_host.cpp"]
HEADERS = ["simplex_cta.cuh", "cta_rt.cuh", "kernels.h", "engine.h", os.path.join("..", "..", "include", "gomilp_b200.h"),
           os.path.join("..", "..", "include", "gomilp_status.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--extended-lambda", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


INC = os.path.join("..", "..", "include")
DEPS = {
    "engine.cu": ["engine.h", "kernels.h", "simplex_cta.cuh", "cta_rt.cuh", os.path.join(INC, "gomilp_b200.h"),
                  os.path.join(INC, "gomilp_status.h")],
    "bnb_device.cu": ["engine.h", "kernels.h", "simplex_cta.cuh", "cta_rt.cuh", os.path.join(INC, "gomilp_b200.h"),
                      os.path.join(INC, "gomilp_status.h")],
    "microbench.cu": ["engine.h", "kernels.h", os.path.join(INC, "gomilp_b200.h"), os.path.join(INC, "gomilp_status.h")],
    "kernels_coop.cu": ["kernels.h", "simplex_cta.cuh", "cta_rt.cuh", os.path.join(INC, "gomilp_status.h")],
    "kernels_reg.cu": ["kernels.h", "simplex_cta.cuh", "cta_rt.cuh", os.path.join(INC, "gomilp_status.h")],
    "kernels_generic.cu": ["kernels.h", "simplex_cta.cuh", "cta_rt.cuh", os.path.join(INC, "gomilp_status.h")],
    "bnb_host.cpp": [os.path.join(INC, "gomilp_b200.h"), os.path.join(INC, "gomilp_status.h")],
}


def _obj_stale(src: str, obj: str) -> bool:
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in [src] + DEPS.get(src, []))


def _stale() -> bool:
    """True when any object is older than its source / headers (an edit made WHILE a build was running leaves a
    library that is newer than the source it lacks, so the library's own time says nothing) or the library is older
    than an object."""
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    for src in SOURCES:
        if not os.path.exists(os.path.join(CSRC, src)):
            continue
        obj = os.path.join(OUT_DIR, os.path.splitext(src)[0] + ".o")
        if _obj_stale(src, obj) or os.path.getmtime(obj) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source of the engine into gomilp_b200/_build/libgomilp_b200.so."""
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    import time
    t_start = time.time()  # objects are stamped with the START of the build: an edit made while it runs makes them stale
    # the two kernel translation units dominate (minutes of cicc each: the solver is one fully inlined function per
    # tier family), so every source is compiled to an object in its own nvcc process, in parallel, then linked
    procs, objs, logs = [], [], []
    for src in srcs:
        obj = os.path.join(OUT_DIR, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        if not force and not _obj_stale(src, obj):
            continue
        cmd = [nvcc] + NVCC_FLAGS + ["-c", "-o", obj, os.path.join(CSRC, src)]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    ok = True
    for cmd, pr in procs:
        out, _ = pr.communicate()
        logs.append(" ".join(cmd) + "\n" + out)
        ok = ok and pr.returncode == 0
        obj = cmd[cmd.index("-o") + 1]
        if pr.returncode == 0 and os.path.exists(obj):
            os.utime(obj, (t_start, t_start))
    if ok:
        cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-ldl"]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        logs.append(" ".join(cmd) + "\n" + res.stdout)
        ok = res.returncode == 0
    log = os.path.join(OUT_DIR, "build.log")
    with open(log, "w") as f:
        f.write("\n".join(logs))
    if verbose or not ok:
        print("\n".join(logs))
    if not ok:
        raise RuntimeError("nvcc failed; see " + log)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
