"""The C++ circuit front end of the C ABI (gcb_circuit_*: Bristol / MPCLC parsers, AssignLevels statistics,
Circuit.Compute) against the Python mirror of the same reference code (mpc_b200/circuit_io.py) on the golden
circuits.  Host-only: runs without a GPU."""
import ctypes as C
import struct

import numpy as np
import pytest

from conftest import load_circuit, millionaire_circuit, mixed_circuit
from mpc_b200 import _lib
from mpc_b200._lib import GcbError, PlanInfo, check, ptr
from mpc_b200.circuit_io import GATE_DTYPE, INV, parse_bristol, parse_mpclc


class CircuitInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "num_gates", "num_wires", "num_inputs", "num_outputs", "num_input_args", "num_output_args",
        "num_xor", "num_xnor", "num_and", "num_or", "num_inv", "num_levels", "max_width")]


def c_parse(data: bytes, fmt: int):
    h = C.c_void_p()
    buf = np.frombuffer(data, dtype=np.uint8)
    check(_lib.lib().gcb_circuit_parse(ptr(buf) if len(buf) else None, len(buf), fmt, C.byref(h)))
    return h


def c_view(h):
    info = CircuitInfo()
    check(_lib.lib().gcb_circuit_get_info(h, C.byref(info)))
    gates = np.zeros(info.num_gates, dtype=GATE_DTYPE)
    check(_lib.lib().gcb_circuit_get_gates(h, ptr(gates) if info.num_gates else None))
    ins, outs = np.zeros(max(info.num_input_args, 1), np.uint32), np.zeros(max(info.num_output_args, 1), np.uint32)
    check(_lib.lib().gcb_circuit_get_io(h, ptr(ins), ptr(outs)))
    return info, gates, ins[: info.num_input_args].tolist(), outs[: info.num_output_args].tolist()


def to_mpclc(circ) -> bytes:
    """circuit/parser.go:71-211 layout (what Circuit.Marshal writes): header, IOArgs, gates."""
    def ioarg(name, bits):
        return (struct.pack(">I", len(name)) + name.encode() + struct.pack(">I", 5) + b"uint0" +
                struct.pack(">II", bits, 0))
    out = struct.pack(">5I", 0x63726300, circ.num_gates, circ.num_wires, len(circ.inputs), len(circ.outputs))
    for k, b in enumerate(circ.inputs):
        out += ioarg(f"i{k}", b)
    for k, b in enumerate(circ.outputs):
        out += ioarg(f"o{k}", b)
    g = circ.gates
    for a, b, c, op in zip(g["in0"].tolist(), g["in1"].tolist(), g["out"].tolist(), g["op"].tolist()):
        out += bytes([op]) + (struct.pack(">2I", a, c) if op == INV else struct.pack(">3I", a, b, c))
    return out


CIRCS = ["add64", "mul64", "aes_128", "mixed", "millionaire"]


def _get(name):
    return {"mixed": lambda: mixed_circuit(9, 800, 30, 10), "millionaire": millionaire_circuit}.get(
        name, lambda: load_circuit(name))()


@pytest.mark.parametrize("name", CIRCS)
@pytest.mark.parametrize("fmt", ["bristol", "mpclc"])
def test_parse_levels_compute_match_the_python_mirror(name, fmt):
    circ = _get(name)
    circ.assign_levels()
    data, code = (circ.to_bristol().encode(), 0) if fmt == "bristol" else (to_mpclc(circ), 1)
    # the mirror reads its own writer's output back (writer check), then the C++ parser reads the same bytes
    back = parse_bristol(data.decode()) if fmt == "bristol" else parse_mpclc(data)
    binary = circ.gates["op"] != INV
    assert all(np.array_equal(back.gates[f], circ.gates[f]) for f in ("in0", "out", "op"))
    assert np.array_equal(back.gates["in1"][binary], circ.gates["in1"][binary])
    h = c_parse(data, code)
    try:
        info, gates, ins, outs = c_view(h)
        assert (info.num_gates, info.num_wires, info.num_inputs, info.num_outputs) == (
            circ.num_gates, circ.num_wires, circ.num_inputs, circ.num_outputs)
        assert ins == list(circ.inputs) and outs == list(circ.outputs)
        for f in ("in0", "out", "op", "level"):
            assert np.array_equal(gates[f], circ.gates[f]), f
        assert np.array_equal(gates["in1"][binary], circ.gates["in1"][binary])
        assert [info.num_xor, info.num_xnor, info.num_and, info.num_or, info.num_inv] == [circ.count(k) for k in range(5)]
        assert (info.num_levels, info.max_width) == (circ.stats["levels"], circ.stats["width"])
        # Circuit.Compute on a batch of wire-bit vectors
        rng = np.random.default_rng(5)
        batch = 3
        bits = rng.integers(0, 2, (batch, circ.num_inputs)).astype(np.uint8)
        out = np.zeros((batch, circ.num_outputs), dtype=np.uint8)
        check(_lib.lib().gcb_circuit_compute(h, batch, ptr(bits), ptr(out)))
        for b in range(batch):
            assert np.array_equal(out[b], circ.compute_bits(bits[b].tolist()))
        # and the plan built from the parsed gates is the plan of the gate array
        ph = C.c_void_p()
        check(_lib.lib().gcb_circuit_plan(h, C.byref(ph)))
        pi = PlanInfo()
        check(_lib.lib().gcb_plan_get_info(ph, C.byref(pi)))
        assert (pi.num_rows, pi.num_and, pi.num_inv) == (circ.num_rows, circ.count(2), circ.count(4))
        _lib.lib().gcb_plan_destroy(ph)
    finally:
        _lib.lib().gcb_circuit_destroy(h)


def test_known_answers_through_the_c_front_end():
    """FIPS-197 C.1 through aes_128.circ and 750000 > 800000 through the millionaire comparator."""
    circ = load_circuit("aes_128")
    h = c_parse(circ.to_bristol().encode(), 0)
    key, pt = int.from_bytes(bytes(range(16)), "big"), int("00112233445566778899aabbccddeeff", 16)
    bits = np.array([[(key >> i) & 1 for i in range(128)] + [(pt >> i) & 1 for i in range(128)]], dtype=np.uint8)
    out = np.zeros((1, 128), dtype=np.uint8)
    check(_lib.lib().gcb_circuit_compute(h, 1, ptr(bits), ptr(out)))
    assert sum(int(b) << i for i, b in enumerate(out[0])) == int("69c4e0d86a7b0430d8cdb78070b4c55a", 16)
    _lib.lib().gcb_circuit_destroy(h)


@pytest.mark.parametrize("text,msg", [
    ("", "invalid 1st line"),
    ("1 3\n2 1 1\n1 1\n\n2 1 0 1 2 NAND\n", "invalid operation 'NAND'"),
    ("1 3\n2 1 1\n1 1\n\n2 1 0 5 2 AND\n", "input 5 of gate 0 not set"),
    ("2 4\n2 1 1\n1 1\n\n2 1 0 1 2 AND\n", "not enough gates: got 1, expected 2"),
    ("1 4\n2 1 1\n1 1\n\n2 1 0 1 2 AND\n", "wire 3 not assigned"),
    ("1 3\n2 1 1\n1 1\n\n1 1 0 1 2 AND\n", "invalid gate"),
    ("1 3\n1 0\n1 1\n\n2 1 0 1 2 AND\n", "no inputs defined"),
])
def test_parser_errors(text, msg):
    with pytest.raises(GcbError, match=msg):
        c_parse(text.encode(), 0)
    with pytest.raises(GcbError):
        c_parse(b"\x00\x01", 1)                           # truncated MPCLC header


REFERENCE = "/root/reference"
REF_FILES = {"aes_128": "pkg/crypto/aes/aes_128.circ", "sha256": "pkg/crypto/sha256/sha256.circ", "mul64": "pkg/math/mul64.circ",
             "sha256xor": "sha2pc/sha256xor.mpclc", "chacha20block": "pkg/crypto/chacha20/chacha20block.mpclc",
             "and": "apps/circuit/and.circ"}


@pytest.mark.skipif(not __import__("os").path.isdir(REFERENCE), reason="the reference tree only exists in the build container")
@pytest.mark.parametrize("name", sorted(REF_FILES))
def test_cpp_parsers_read_the_reference_files(name):
    """The C++ front end on the reference's own circuit files, byte for byte as shipped (Bristol text and the MPCLC
    binary format of circuit/parser.go:71-211 / Circuit.Marshal): gates, wires and I/O sizes equal the golden fixture
    (whose sha256xor reading reproduces the reference's transcript hashes, tests/test_reference_transcript.py)."""
    import os
    path = os.path.join(REFERENCE, REF_FILES[name])
    data = open(path, "rb").read()
    h = c_parse(data, 1 if path.endswith(".mpclc") else 0)
    try:
        info, gates, ins, outs = c_view(h)
        want = load_circuit(name)
        assert (info.num_gates, info.num_wires) == (want.num_gates, want.num_wires)
        assert ins == list(want.inputs) and outs == list(want.outputs)
        for f in ("in0", "in1", "out", "op"):
            assert np.array_equal(gates[f], want.gates[f]), f
    finally:
        _lib.lib().gcb_circuit_destroy(h)


def test_parsers_survive_mutated_files():
    """Circuit files can come from a peer: thousands of mutations (flipped bytes, truncations, insertions, huge 32-bit
    fields) of a valid Bristol file and a valid MPCLC file must each be either refused with an error or parsed into a
    circuit that also plans (or is refused there) -- never a crash, a hang or an allocation sized by the peer."""
    circ = mixed_circuit(11, 120, 12, 5)
    good = {0: circ.to_bristol().encode(), 1: to_mpclc(circ)}
    rng = np.random.default_rng(1)
    parsed = 0
    for fmt, data in good.items():
        for _ in range(1500):
            b = bytearray(data)
            kind = int(rng.integers(0, 4))
            if kind == 0:
                for _ in range(int(rng.integers(1, 6))):
                    b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            elif kind == 1:
                b = b[: int(rng.integers(0, len(b)))]
            elif kind == 2:
                i = int(rng.integers(0, len(b)))
                b[i:i] = bytes(rng.integers(0, 256, int(rng.integers(1, 9)), dtype=np.uint8))
            else:
                i = int(rng.integers(0, max(1, len(b) - 4)))
                b[i:i + 4] = struct.pack(">I", int(rng.choice([0xffffffff, 0x7fffffff, 0x80000000, 1 << 20])))
            h = C.c_void_p()
            buf = np.frombuffer(bytes(b), dtype=np.uint8) if len(b) else np.zeros(0, np.uint8)
            if _lib.lib().gcb_circuit_parse(ptr(buf) if len(buf) else None, len(buf), fmt, C.byref(h)) != 0:
                continue
            parsed += 1
            ph = C.c_void_p()
            if _lib.lib().gcb_circuit_plan(h, C.byref(ph)) == 0:
                _lib.lib().gcb_plan_destroy(ph)
            _lib.lib().gcb_circuit_destroy(h)
    assert parsed > 0
