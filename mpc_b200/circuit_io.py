"""Circuit containers and parsers (host side, no device code).

Mirrors the reference's circuit description types so that the same files and
the same gate arrays drive the oracle, the C-ABI library and the tests:

* ``GATE_DTYPE`` is the 20-byte in-memory layout of ``circuit.Gate``
  (reference circuit/circuit.go:260-266, size pinned by
  circuit/circuit_test.go:14-19): ``u32 Input0, u32 Input1, u32 Output,
  u8 Op, 3 pad bytes, u32 Level``.
* ``parse_bristol`` follows circuit/parser.go:265-494 (text format: gate
  count/wire count line, input line, output line, one gate per line).
* ``parse_mpclc`` follows circuit/parser.go:71-211 (big-endian binary:
  magic, ngates, nwires, ninputs, noutputs, IOArgs, then gates).
* ``Circuit.compute`` follows circuit/computer.go:15-91 (plaintext
  evaluation, used by the tests to check decoded garbled outputs).
* ``Circuit.assign_levels`` follows circuit/circuit.go:206-254 (TargetYao).

Operation codes are the reference's (circuit/circuit.go:24-34).
"""
from __future__ import annotations

import io
import struct
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

XOR, XNOR, AND, OR, INV = 0, 1, 2, 3, 4
OP_NAMES = {"XOR": XOR, "XNOR": XNOR, "AND": AND, "OR": OR, "INV": INV}
OP_ROWS = {XOR: 0, XNOR: 0, AND: 2, OR: 3, INV: 1}      # garbled rows per gate
OP_TWEAKS = {XOR: 0, XNOR: 0, AND: 2, OR: 1, INV: 1}    # tweak ids consumed

GATE_DTYPE = np.dtype(
    {
        "names": ["in0", "in1", "out", "op", "level"],
        "formats": ["<u4", "<u4", "<u4", "u1", "<u4"],
        "offsets": [0, 4, 8, 12, 16],
        "itemsize": 20,
    }
)

LABEL_DTYPE = np.dtype([("d0", "<u8"), ("d1", "<u8")])              # ot.Label
WIRE_DTYPE = np.dtype([("l0", LABEL_DTYPE), ("l1", LABEL_DTYPE)])   # ot.Wire


class CircuitError(ValueError):
    pass


@dataclass
class Circuit:
    num_gates: int
    num_wires: int
    inputs: List[int]            # bits per input argument (IO.Size() == sum)
    outputs: List[int]           # bits per output argument
    gates: np.ndarray            # GATE_DTYPE[num_gates]
    name: str = ""
    stats: dict = field(default_factory=dict)

    # -- sizes ------------------------------------------------------------
    @property
    def num_inputs(self) -> int:
        return int(sum(self.inputs))

    @property
    def num_outputs(self) -> int:
        return int(sum(self.outputs))

    def count(self, op: int) -> int:
        return int(np.count_nonzero(self.gates["op"] == op))

    @property
    def num_rows(self) -> int:
        """Garbled-table labels per instance (slab size, garble.go:195-207)."""
        ops = self.gates["op"]
        return int(2 * np.count_nonzero(ops == AND) + 3 * np.count_nonzero(ops == OR)
                   + np.count_nonzero(ops == INV))

    def row_offsets(self) -> np.ndarray:
        """Slab offset of each gate's first row (exclusive prefix sum)."""
        rows = np.array([0, 0, 2, 3, 1], dtype=np.int64)[self.gates["op"]]
        off = np.zeros(self.num_gates + 1, dtype=np.int64)
        np.cumsum(rows, out=off[1:])
        return off

    # -- reference algorithms ---------------------------------------------
    def assign_levels(self) -> None:
        """TargetYao levels (circuit/circuit.go:206-254)."""
        levels = np.zeros(self.num_wires, dtype=np.uint32)
        g = self.gates
        in0, in1, out, op = g["in0"], g["in1"], g["out"], g["op"]
        lv = np.zeros(self.num_gates, dtype=np.uint32)
        for i in range(self.num_gates):
            level = levels[in0[i]]
            if op[i] != INV:
                l1 = levels[in1[i]]
                if l1 > level:
                    level = l1
            lv[i] = level
            levels[out[i]] = level + 1
        self.gates["level"] = lv
        self.stats["levels"] = int(levels.max()) if self.num_wires else 0
        self.stats["width"] = int(np.bincount(lv).max()) if self.num_gates else 0

    def compute_bits(self, in_bits: Sequence[int]) -> np.ndarray:
        """Plaintext evaluation on wire bits; returns the output wire bits."""
        if len(in_bits) != self.num_inputs:
            raise CircuitError(f"invalid inputs: got {len(in_bits)}, expected {self.num_inputs}")
        w = np.zeros(self.num_wires, dtype=np.uint8)
        w[: self.num_inputs] = np.asarray(in_bits, dtype=np.uint8) & 1
        wl = w.tolist()
        g = self.gates
        for a, b, c, op in zip(g["in0"].tolist(), g["in1"].tolist(), g["out"].tolist(), g["op"].tolist()):
            if op == XOR:
                wl[c] = wl[a] ^ wl[b]
            elif op == XNOR:
                wl[c] = 1 ^ wl[a] ^ wl[b]
            elif op == AND:
                wl[c] = wl[a] & wl[b]
            elif op == OR:
                wl[c] = wl[a] | wl[b]
            elif op == INV:
                wl[c] = 1 ^ wl[a]
            else:
                raise CircuitError(f"invalid gate {op}")
        return np.array(wl[self.num_wires - self.num_outputs:], dtype=np.uint8)

    def compute(self, values: Sequence[int]) -> List[int]:
        """circuit.Compute: one big integer per input argument, LSB = first wire."""
        if len(values) != len(self.inputs):
            raise CircuitError(f"invalid inputs: got {len(values)}, expected {len(self.inputs)}")
        bits: List[int] = []
        for v, n in zip(values, self.inputs):
            bits.extend((v >> i) & 1 for i in range(n))
        ob = self.compute_bits(bits)
        res, pos = [], 0
        for n in self.outputs:
            res.append(sum(int(ob[pos + i]) << i for i in range(n)))
            pos += n
        return res

    # -- fixture (de)serialisation -----------------------------------------
    def save_npz(self, path: str) -> None:
        g = self.gates
        np.savez_compressed(
            path,
            meta=np.array([self.num_gates, self.num_wires], dtype=np.int64),
            inputs=np.array(self.inputs, dtype=np.int64),
            outputs=np.array(self.outputs, dtype=np.int64),
            op=g["op"].copy(),
            # deltas against the gate index compress far better than raw ids
            in0=(g["in0"].astype(np.int64) - np.arange(self.num_gates)).astype(np.int32),
            in1=(g["in1"].astype(np.int64) - np.arange(self.num_gates)).astype(np.int32),
            out=(g["out"].astype(np.int64) - np.arange(self.num_gates)).astype(np.int32),
        )

    @staticmethod
    def load_npz(path: str, name: str = "") -> "Circuit":
        z = np.load(path)
        ng, nw = (int(x) for x in z["meta"])
        gates = np.zeros(ng, dtype=GATE_DTYPE)
        idx = np.arange(ng, dtype=np.int64)
        gates["op"] = z["op"]
        gates["in0"] = (z["in0"].astype(np.int64) + idx).astype(np.uint32)
        gates["in1"] = (z["in1"].astype(np.int64) + idx).astype(np.uint32)
        gates["out"] = (z["out"].astype(np.int64) + idx).astype(np.uint32)
        c = Circuit(ng, nw, [int(x) for x in z["inputs"]], [int(x) for x in z["outputs"]], gates, name)
        c.validate()
        return c

    def validate(self) -> None:
        """The parser's "wire seen" checks (parser.go:97-104,140-160,199-204)."""
        seen = np.zeros(self.num_wires, dtype=bool)
        if self.num_inputs > self.num_wires:
            raise CircuitError("more input wires than wires")
        seen[: self.num_inputs] = True
        g = self.gates
        for i, (a, b, c, op) in enumerate(zip(g["in0"].tolist(), g["in1"].tolist(),
                                              g["out"].tolist(), g["op"].tolist())):
            if op > INV:
                raise CircuitError(f"unsupported gate type {op}")
            if a >= self.num_wires or not seen[a]:
                raise CircuitError(f"input {a} of gate {i} not set")
            if op != INV and (b >= self.num_wires or not seen[b]):
                raise CircuitError(f"input {b} of gate {i} not set")
            if c >= self.num_wires:
                raise CircuitError(f"wire {c} out of range")
            seen[c] = True
        if not seen.all():
            raise CircuitError(f"wire {int(np.argmin(seen))} not assigned")

    def to_bristol(self) -> str:
        out = io.StringIO()
        out.write(f"{self.num_gates} {self.num_wires}\n")
        out.write(f"{len(self.inputs)} " + " ".join(str(b) for b in self.inputs) + "\n")
        out.write(f"{len(self.outputs)} " + " ".join(str(b) for b in self.outputs) + "\n\n")
        names = {v: k for k, v in OP_NAMES.items()}
        g = self.gates
        for a, b, c, op in zip(g["in0"].tolist(), g["in1"].tolist(), g["out"].tolist(), g["op"].tolist()):
            if op == INV:
                out.write(f"1 1 {a} {c} INV\n")
            else:
                out.write(f"2 1 {a} {b} {c} {names[op]}\n")
        return out.getvalue()


def _mk(gl: list, nw: int, inputs: List[int], outputs: List[int], name: str) -> Circuit:
    gates = np.zeros(len(gl), dtype=GATE_DTYPE)
    if gl:
        arr = np.array(gl, dtype=np.int64)
        gates["in0"], gates["in1"], gates["out"], gates["op"] = arr[:, 0], arr[:, 1], arr[:, 2], arr[:, 3]
    c = Circuit(len(gl), nw, inputs, outputs, gates, name)
    c.validate()
    return c


def parse_bristol(text: str, name: str = "") -> Circuit:
    """Bristol text format, as circuit/parser.go:265-494 reads it."""
    lines = [ln.split() for ln in text.splitlines()]
    lines = [ln for ln in lines if ln]                 # readLine skips blank lines
    if len(lines) < 3 or len(lines[0]) != 2:
        raise CircuitError("invalid 1st line")
    ng, nw = int(lines[0][0]), int(lines[0][1])
    niv = int(lines[1][0])
    if 1 + niv != len(lines[1]):
        raise CircuitError("invalid inputs line")
    inputs = [int(x) for x in lines[1][1:]]
    if sum(inputs) == 0:
        raise CircuitError("no inputs defined")
    nov = int(lines[2][0])
    if 1 + nov != len(lines[2]):
        raise CircuitError("invalid outputs line")
    outputs = [int(x) for x in lines[2][1:]]
    gl = []
    for ln in lines[3:]:
        if len(gl) >= ng:
            raise CircuitError("too many gates")
        if len(ln) < 3:
            raise CircuitError(f"invalid gate: {ln}")
        n1, n2 = int(ln[0]), int(ln[1])
        if 2 + n1 + n2 + 1 != len(ln):
            raise CircuitError(f"invalid gate: {ln}")
        opname = ln[-1]
        if opname not in OP_NAMES:
            raise CircuitError(f"invalid operation '{opname}'")
        op = OP_NAMES[opname]
        want = 1 if op == INV else 2
        if n1 != want:
            raise CircuitError(f"invalid number of inputs {n1} for {opname}")
        if n2 != 1:
            raise CircuitError(f"invalid number of outputs {n2} for {opname}")
        ins = [int(x) for x in ln[2:2 + n1]]
        gl.append((ins[0], ins[1] if n1 > 1 else 0, int(ln[2 + n1]), op))
    if len(gl) != ng:
        raise CircuitError(f"not enough gates: got {len(gl)}, expected {ng}")
    return _mk(gl, nw, inputs, outputs, name)


def _read_ioarg(buf: memoryview, pos: int):
    def rd_str(p):
        (n,) = struct.unpack_from(">I", buf, p)
        p += 4
        return bytes(buf[p:p + n]).decode(), p + n

    nm, pos = rd_str(pos)
    ty, pos = rd_str(pos)
    bits, ncomp = struct.unpack_from(">II", buf, pos)
    pos += 8
    for _ in range(ncomp):
        _, pos = _read_ioarg(buf, pos)
    return (nm, ty, bits), pos


def parse_mpclc(data: bytes, name: str = "") -> Circuit:
    """MPCLC binary format, as circuit/parser.go:71-211 reads it."""
    buf = memoryview(data)
    _magic, ng, nw, ni, no = struct.unpack_from(">5I", buf, 0)
    pos = 20
    inputs, outputs = [], []
    for _ in range(ni):
        a, pos = _read_ioarg(buf, pos)
        inputs.append(a[2])
    for _ in range(no):
        a, pos = _read_ioarg(buf, pos)
        outputs.append(a[2])
    gl = []
    n = len(data)
    while pos < n:
        op = buf[pos]
        pos += 1
        if op in (XOR, XNOR, AND, OR):
            a, b, c = struct.unpack_from(">3I", buf, pos)
            pos += 12
            gl.append((a, b, c, op))
        elif op == INV:
            a, c = struct.unpack_from(">2I", buf, pos)
            pos += 8
            gl.append((a, 0, c, op))
        else:
            raise CircuitError(f"unsupported gate type {op}")
    if len(gl) != ng:
        raise CircuitError(f"not enough gates: got {len(gl)}, expected {ng}")
    return _mk(gl, nw, inputs, outputs, name)


def parse_file(path: str) -> Circuit:
    """circuit.Parse (parser.go:53-68): pick the format from the suffix."""
    nm = path.rsplit("/", 1)[-1]
    if path.endswith(".circ") or path.endswith(".bristol"):
        with open(path, "r") as f:
            return parse_bristol(f.read(), nm)
    if path.endswith(".mpclc"):
        with open(path, "rb") as f:
            return parse_mpclc(f.read(), nm)
    if path.endswith(".npz"):
        return Circuit.load_npz(path, nm)
    raise CircuitError("unsupported circuit format")
