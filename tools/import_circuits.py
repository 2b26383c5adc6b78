#!/usr/bin/env python
"""Convert the circuit description files shipped with the reference into the
compact ``.npz`` fixtures under ``tests/golden/circuits/``.

Run in the build container only (``/root/reference`` does not exist on the GPU
box).  The fixtures are data, not code: gate lists of the public Bristol
circuits and of the MPCLC circuits the reference embeds, re-encoded
column-wise (op, in0-i, in1-i, out-i) and deflated.

    python tools/import_circuits.py [/root/reference]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from mpc_b200.circuit_io import parse_file  # noqa: E402

SOURCES = {
    "aes_128": "pkg/crypto/aes/aes_128.circ",
    "aes_256": "pkg/crypto/aes/aes_256.circ",
    "sha256": "pkg/crypto/sha256/sha256.circ",
    "sha512": "pkg/crypto/sha512/sha512.circ",
    "add64": "pkg/math/add64.circ",
    "sub64": "pkg/math/sub64.circ",
    "mul64": "pkg/math/mul64.circ",
    "div64": "pkg/math/div64.circ",
    "sha256xor": "sha2pc/sha256xor.mpclc",
    "chacha20block": "pkg/crypto/chacha20/chacha20block.mpclc",
    "and": "apps/circuit/and.circ",
    "not": "apps/circuit/not.circ",
}


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    dst = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "circuits")
    os.makedirs(dst, exist_ok=True)
    for name, rel in SOURCES.items():
        c = parse_file(os.path.join(ref, rel))
        out = os.path.join(dst, name + ".npz")
        c.save_npz(out)
        print(f"{name:14s} gates={c.num_gates:7d} wires={c.num_wires:7d} in={c.inputs} out={c.outputs} "
              f"AND={c.count(2)} OR={c.count(3)} INV={c.count(4)} rows={c.num_rows} -> {os.path.getsize(out)} B")


if __name__ == "__main__":
    main()
