"""The plan compiler checked without a GPU: tools/plan_check.cpp builds plans of the shipped circuits and of random
circuits under every schedule policy, several fan-in limits and hot caps (whole-life hot / cold, split live ranges) and
interprets each one the way the gate kernels do, with fingerprints in place of labels (every gate must find its input
wires, its static tweak and row; every node must produce its wire; no unordered step may read what it writes; no reload
may be queued before its value has reached the scratch)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import load_circuit, mixed_circuit

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def plan_check(tmp_path_factory):
    out = tmp_path_factory.mktemp("plan_check")
    exe = str(out / "plan_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "tools", "plan_check.cpp"), os.path.join(ROOT, "mpc_b200", "csrc", "plan.cpp")], check=True)
    return exe, out


def _dump(circ, path):
    with open(path, "wb") as f:
        f.write(np.array([circ.num_gates, circ.num_wires, circ.num_inputs, circ.num_outputs], dtype=np.uint32).tobytes())
        f.write(np.ascontiguousarray(circ.gates).tobytes())


@pytest.mark.parametrize("mode", ["0", "1"])
def test_plans_of_the_shipped_circuits_interpret_correctly(plan_check, mode):
    exe, out = plan_check
    files = []
    for name in ("add64", "mul64", "aes_128", "sha256", "sha256xor", "chacha20block"):
        files.append(str(out / f"{name}.gates"))
        _dump(load_circuit(name), files[-1])
    files.append(str(out / "mixed.gates"))
    _dump(mixed_circuit(3, 4000, 64, 32), files[-1])
    r = subprocess.run([exe] + files, capture_output=True, text=True, env=dict(os.environ, GCB_HOT_MODE=mode))
    assert r.returncode == 0, r.stderr[-2000:] + r.stdout
    assert " 0 errors" in r.stdout


def test_plans_of_random_circuits_interpret_correctly(plan_check):
    """Random circuits with all five gate types, narrow and deep to wide and shallow; among the caps tried are the
    degenerate ones just below a circuit's all-hot need, where only tiny gaps fit (the regime in which the compiler once
    queued reloads of values that had not been born yet)."""
    exe, _ = plan_check
    r = subprocess.run([exe, "--random", "40"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:] + r.stdout
    assert " 0 errors" in r.stdout
