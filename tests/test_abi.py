"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads and
exports every symbol include/gcb200.h declares; the host-side plan compiler
agrees with the circuit's static row / tweak layout; and the product has no CPU
fallback (compute calls fail loudly without a device)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_circuit, millionaire_circuit, mixed_circuit
from mpc_b200 import _lib
from mpc_b200.circuit import GarbleEngine, hash_half
from mpc_b200.circuit_io import AND, INV, LABEL_DTYPE, OR

HEADER = os.path.join(ROOT, "include", "gcb200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gcb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"libgcb200.so does not export {n}"
    assert set(names) == set(_lib.SYMBOLS), set(names) ^ set(_lib.SYMBOLS)
    assert b"sm_100a" in L.gcb_version()


def test_library_has_sm100a_code_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib._build.SO], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


@pytest.mark.parametrize("name", ["and", "not", "add64", "sub64", "mul64", "aes_128", "sha256", "sha256xor",
                                  "chacha20block", "mixed", "millionaire"])
def test_plan_matches_static_layout(name):
    circ = {"mixed": lambda: mixed_circuit(3, 400, 20, 8), "millionaire": millionaire_circuit}.get(
        name, lambda: load_circuit(name))()
    eng = GarbleEngine(circ)
    i = eng.info
    ops = circ.gates["op"]
    n_and, n_or, n_inv = (int(np.count_nonzero(ops == o)) for o in (AND, OR, INV))
    assert (i.num_gates, i.num_wires, i.num_inputs, i.num_outputs) == (
        circ.num_gates, circ.num_wires, circ.num_inputs, circ.num_outputs)
    assert (i.num_and, i.num_or, i.num_inv) == (n_and, n_or, n_inv)
    assert i.num_rows == circ.num_rows == 2 * n_and + 3 * n_or + n_inv
    assert i.num_tweaks == 2 * n_and + n_or + n_inv
    assert i.garble_hashes == 4 * n_and + 4 * n_or + 2 * n_inv
    assert i.eval_hashes == 2 * n_and + n_or + n_inv
    assert np.array_equal(eng.row_off, circ.row_offsets())
    assert 1 <= i.teams_per_sm <= 32 and i.teams_per_sm * i.team_threads <= 1024
    assert i.num_slots >= 1 and i.num_steps >= 1


def test_survey_sizes_of_the_baseline_circuits():
    a, s = GarbleEngine(load_circuit("aes_128")).info, GarbleEngine(load_circuit("sha256")).info
    assert (a.num_gates, a.num_and, a.num_inv, a.num_rows) == (36663, 6400, 2087, 14887)
    assert (s.num_gates, s.num_and, s.num_inv, s.num_rows) == (135073, 22573, 1856, 47002)
    x = GarbleEngine(load_circuit("sha256xor")).info
    assert x.num_rows == 42914                       # sha2pc/params.go:26


def test_plan_rejects_bad_circuits():
    circ = load_circuit("add64")
    g = circ.gates.copy()
    g["op"][5] = 9
    h = C.c_void_p()
    rc = _lib.lib().gcb_plan_create(_lib.ptr(g), len(g), circ.num_wires, circ.num_inputs, circ.num_outputs, C.byref(h))
    assert rc == _lib.E_BADOP and b"invalid gate type" in _lib.lib().gcb_last_error()
    g = circ.gates.copy()
    g["in0"][0] = circ.num_wires + 7
    rc = _lib.lib().gcb_plan_create(_lib.ptr(g), len(g), circ.num_wires, circ.num_inputs, circ.num_outputs, C.byref(h))
    assert rc == _lib.E_WIRE
    g = circ.gates.copy()
    g["in0"][0] = circ.num_wires - 1                 # read before assignment
    rc = _lib.lib().gcb_plan_create(_lib.ptr(g), len(g), circ.num_wires, circ.num_inputs, circ.num_outputs, C.byref(h))
    assert rc == _lib.E_WIRE and b"not set" in _lib.lib().gcb_last_error()


def test_iknp_size_helpers():
    L = _lib.lib()
    for n, (u, adv) in {0: (0, 0), 1: (128, 1), 129: (17 * 128, 17), 512: (8192, 64), 513: (8192 + 128, 65),
                        1100: (2 * 8192 + 10 * 128, 138), 1 << 24: (1 << 28, 1 << 21)}.items():
        assert L.gcb_iknp_u_size(n) == u and L.gcb_iknp_stream_advance(n) == adv


def test_bad_key_length_is_rejected_before_the_device():
    x = np.zeros(4, dtype=LABEL_DTYPE)
    with pytest.raises(_lib.GcbError) as e:
        hash_half(b"\0" * 15, x)
    assert e.value.rc == _lib.E_KEYLEN and "invalid key size" in str(e.value)


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box WITHOUT a GPU")
def test_no_cpu_fallback_without_a_device():
    x = np.zeros(4, dtype=LABEL_DTYPE)
    with pytest.raises(_lib.GcbError) as e:
        hash_half(b"\0" * 16, x)
    assert e.value.rc == _lib.E_CUDA and "no CPU path" in str(e.value)
    eng = GarbleEngine(load_circuit("and"))
    with pytest.raises(_lib.GcbError) as e:
        eng.garble(b"\1" * 48, b"\0" * 16)
    assert e.value.rc == _lib.E_CUDA


def test_plan_limits_of_live_labels():
    """A circuit that keeps more labels live than shared memory holds still gets a plan (the excess is spilled
    to a global-memory scratch by a dedicated kernel variant); only the 16-bit slot numbering of the plan
    records is a hard limit: GCB_E_TOO_LARGE at plan creation (host-only check)."""
    import ctypes as C
    from mpc_b200._lib import PlanInfo
    from mpc_b200.circuit_io import parse_bristol

    def plan(n):
        lines = [f"2 1 {i} {n + i} {2 * n + i} AND" for i in range(n)]
        circ = parse_bristol(f"{n} {3 * n}\n2 {n} {n}\n1 {n}\n\n" + "\n".join(lines) + "\n", "wide")
        h = C.c_void_p()
        gates = np.ascontiguousarray(circ.gates)
        rc = _lib.lib().gcb_plan_create(_lib.ptr(gates), circ.num_gates, circ.num_wires, circ.num_inputs,
                                        circ.num_outputs, C.byref(h))
        return rc, h

    rc, h = plan(6000)                                   # 18,000 live labels: spilling variant, one team per SM
    assert rc == 0
    info = PlanInfo()
    _lib.check(_lib.lib().gcb_plan_get_info(h, C.byref(info)))
    assert info.num_slots == 18000 and info.teams_per_sm == 1
    _lib.lib().gcb_plan_destroy(h)
    rc, h = plan(22000)                                  # 66,000 > 65,535
    assert rc == _lib.E_TOO_LARGE and b"65535" in _lib.lib().gcb_last_error()


def test_plan_choice_by_batch():
    """gcb_plan_get_info_for_batch (host only): a deep, narrow circuit whose batch overflows the resident instances of
    its all-hot plan runs on the plan with split live ranges (16 instances per SM, a hot subset of the labels in shared
    memory); a batch that fits, a wide circuit and a circuit of which 16 instances already fit stay on the default plan."""
    eng = GarbleEngine(load_circuit("sha256"))
    base = eng.info
    assert base.teams_per_sm == 8 and base.num_hot_slots == base.num_slots
    fits = eng.info_for(8 * 148)
    assert (fits.teams_per_sm, fits.num_slots, fits.num_hot_slots) == (8, base.num_slots, base.num_slots)
    many = eng.info_for(8 * 148 + 1)
    assert many.teams_per_sm == 16 and many.team_threads == 32 and many.num_hot_slots < many.num_slots
    assert many.num_rows == base.num_rows and many.num_and == base.num_and        # the same tables whatever the plan
    aes = GarbleEngine(load_circuit("aes_128"))
    wide = aes.info_for(1 << 20)
    assert wide.teams_per_sm == aes.info.teams_per_sm and wide.num_hot_slots == wide.num_slots
    mul = GarbleEngine(load_circuit("mul64"))
    assert mul.info_for(1 << 20).num_hot_slots == mul.info.num_slots


def test_staging_copier_moves_every_byte():
    """The parallel staging copier behind pageable host buffers (async.hpp: 2 MB pieces pulled by a few workers and the
    caller): sizes around the piece and threshold boundaries, and four callers at once."""
    import ctypes as C
    import threading
    fn = _lib.lib().gcb_debug_host_copy
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    fn.restype = None
    rng = np.random.default_rng(5)
    for n in (0, 1, (8 << 20) - 1, 8 << 20, (8 << 20) + 1, (32 << 20) + 12345):
        src = rng.integers(0, 256, n, dtype=np.uint8)
        dst = np.zeros(n + 64, dtype=np.uint8)
        fn(dst.ctypes.data + 32, src.ctypes.data, n)
        assert np.array_equal(dst[32:32 + n], src) and not dst[:32].any() and not dst[32 + n:].any()
    srcs = [rng.integers(0, 256, (24 << 20) + 7 * k, dtype=np.uint8) for k in range(4)]
    dsts = [np.zeros_like(s) for s in srcs]
    ts = [threading.Thread(target=lambda s=s, d=d: [fn(d.ctypes.data, s.ctypes.data, s.size) for _ in range(3)]) for s, d in zip(srcs, dsts)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert all(np.array_equal(s, d) for s, d in zip(srcs, dsts))


def test_geometry_choices_of_the_shipped_circuits():
    """Team geometry the plan reports (host-only; `compute_geometry`): wide circuits take the table count that gives the
    most warps, then the most teams (aes_128 / aes_256: 8 two-warp teams on two tables, mul64: 16 one-warp teams);
    deep, narrow ones keep as many one-warp instances as fit (sha256 8, chacha20 10), sha512's 2,401 labels leave room
    for three 96-thread teams."""
    want = {"aes_128": (8, 64), "aes_256": (8, 64), "mul64": (16, 32), "sha256": (8, 32), "chacha20block": (10, 32),
            "sha512": (3, 96), "add64": (16, 32)}
    for name, geo in want.items():
        i = GarbleEngine(load_circuit(name)).info
        assert (i.teams_per_sm, i.team_threads) == geo, (name, i.teams_per_sm, i.team_threads)
        assert i.teams_per_sm * i.team_threads <= 512


def test_plan_create_survives_corrupted_gate_arrays():
    """gcb_plan_create on gate arrays with corrupted fields (wires out of range, reads before assignment, invalid ops)
    and wire / input / output counts up to 2^32 - 1: refused or planned, quickly -- never a crash or an allocation sized
    by the corrupted count (a wire count of 0xffffffff once cost 32 GB and a minute before it was refused)."""
    import ctypes as C
    import time
    base = mixed_circuit(5, 400, 20, 8)
    rng = np.random.default_rng(3)
    t0 = time.time()
    planned = 0
    for _ in range(1200):
        g = np.ascontiguousarray(base.gates).copy()
        nw, nin, nout = base.num_wires, base.num_inputs, base.num_outputs
        for _ in range(int(rng.integers(1, 5))):
            i = int(rng.integers(0, len(g)))
            f = ["in0", "in1", "out", "op"][int(rng.integers(0, 4))]
            g[f][i] = int(rng.integers(0, 9)) if f == "op" else int(rng.choice([0, 1, nw - 1, nw, nw + 5, 0xffffffff, 7, 255]))
        kind = int(rng.integers(0, 6))
        if kind == 0:
            nw = int(rng.choice([0, 1, nw - 1, nw + 1, 70000, 0xffffffff]))
        elif kind == 1:
            nin = int(rng.choice([0, nw, nw + 1, 0xffffffff]))
        elif kind == 2:
            nout = int(rng.choice([0, nw, nw + 1, 0xffffffff]))
        h = C.c_void_p()
        if _lib.lib().gcb_plan_create(_lib.ptr(g), len(g), nw, nin, nout, C.byref(h)) == 0:
            planned += 1
            _lib.lib().gcb_plan_destroy(h)
    assert planned > 0 and time.time() - t0 < 120
