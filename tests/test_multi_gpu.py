"""GPU, needs >= 2 devices (skipped otherwise): the B&B wavefront sharded over GPUs through gm_comm_init +
gm_milp_solve_device must reproduce the 1-GPU search exactly (status, nodes, pivots, z, x, decision log)."""
import json
import os
import subprocess
import sys

import pytest

import gomilp_b200 as gm

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.timeout(900)
def test_sharded_bnb_is_identical_to_one_gpu():
    world = min(gm.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    for extra in (["--n", "24", "--m", "4", "--nodes", "2000"], ["--kind", "c5", "--n", "50", "--nodes", "256"]):
        res = subprocess.run([sys.executable, os.path.join(HERE, "multi_gpu_probe.py"), "--world", str(world), "--reps", "0"]
                             + extra, capture_output=True, text=True, timeout=800)
        assert res.returncode == 0, res.stdout + res.stderr
        line = json.loads(res.stdout.strip().splitlines()[-1])
        assert line["identical_to_1gpu"] and line["nodes"] > 1


@pytest.mark.timeout(600)
def test_one_host_thread_per_device_in_one_process():
    """tests/cpp/two_devices.cpp: the C ABI driven from C++ threads, one per GPU (ADVICE r1 / INTEGRATION.md)."""
    if gm.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(HERE)
    exe = os.path.join(HERE, "cpp", "_build", "two_devices")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    libdir = os.path.join(root, "gomilp_b200", "_build")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(root, "include"),
                    os.path.join(HERE, "cpp", "two_devices.cpp"), "-L" + libdir, "-lgomilp_b200", "-lpthread",
                    "-Wl,-rpath," + libdir, "-o", exe], check=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=500)
    assert res.returncode == 0, res.stdout + res.stderr
