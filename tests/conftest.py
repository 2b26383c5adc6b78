import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mpc_b200.circuit_io import Circuit, parse_bristol  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


_cache = {}


def load_circuit(name: str) -> Circuit:
    if name not in _cache:
        _cache[name] = Circuit.load_npz(os.path.join(GOLDEN, "circuits", name + ".npz"), name)
    return _cache[name]


@pytest.fixture(scope="session")
def circuits():
    return load_circuit


def mixed_circuit(seed: int = 1, ngates: int = 300, nin: int = 24, nout: int = 8) -> Circuit:
    """Random circuit using all five gate types (the shipped Bristol files have
    no OR/XNOR); the last `nout` wires are the outputs, as the format requires."""
    rng = np.random.default_rng(seed)
    lines = []
    nw = nin
    for g in range(ngates):
        op = ["XOR", "XNOR", "AND", "OR", "INV"][int(rng.integers(0, 5))]
        lo = max(0, nw - 40)
        a = int(rng.integers(lo, nw))
        b = int(rng.integers(0, nw))
        if op == "INV":
            lines.append(f"1 1 {a} {nw} INV")
        else:
            lines.append(f"2 1 {a} {b} {nw} {op}")
        nw += 1
    text = f"{ngates} {nw}\n2 {nin // 2} {nin - nin // 2}\n1 {nout}\n\n" + "\n".join(lines) + "\n"
    return parse_bristol(text, f"mixed{seed}")


def millionaire_circuit() -> Circuit:
    """Hand-written stand-in for apps/garbled/examples/millionaire.mpcl (which
    needs the Go MPCL compiler): signed int64 ``a > b`` built gate for gate the
    way compiler/circuits/circ_comparators.go:16-54 (intComparator) does --
    per bit XNOR(cin,y) XOR(cin,x) AND XOR(cin,.), then a sign MUX
    r = cout ^ (cond & (y63 ^ cout)), cond = x63 ^ y63.  2x64 inputs, 1 output."""
    n = 64
    lines = []
    nw = 2 * n
    lines.append(f"2 1 0 0 {nw} XOR")              # constant 0 = cin (tests x > y)
    cin = nw
    nw += 1
    for i in range(n):
        x, y = i, n + i
        lines.append(f"2 1 {cin} {y} {nw} XNOR")
        lines.append(f"2 1 {cin} {x} {nw + 1} XOR")
        lines.append(f"2 1 {nw} {nw + 1} {nw + 2} AND")
        lines.append(f"2 1 {cin} {nw + 2} {nw + 3} XOR")
        cin = nw + 3
        nw += 4
    xs, ys = n - 1, 2 * n - 1
    lines.append(f"2 1 {xs} {ys} {nw} XOR")        # cond
    lines.append(f"2 1 {ys} {cin} {nw + 1} XOR")   # y63 ^ cout
    lines.append(f"2 1 {nw} {nw + 1} {nw + 2} AND")
    lines.append(f"2 1 {cin} {nw + 2} {nw + 3} XOR")
    nw += 4
    text = f"{len(lines)} {nw}\n2 {n} {n}\n1 1\n\n" + "\n".join(lines) + "\n"
    return parse_bristol(text, "millionaire")


def bits_of(value: int, n: int):
    return [(value >> i) & 1 for i in range(n)]


def int_of(bits) -> int:
    return sum(int(b) << i for i, b in enumerate(bits))
