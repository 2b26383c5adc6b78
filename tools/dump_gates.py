"""Write the golden circuits as raw gate files for the host-only plan model (tools/plan_model.cpp):
u32 num_gates, num_wires, num_inputs, num_outputs, then num_gates 20-byte circuit.Gate records.

  python tools/dump_gates.py [circuit ...]          -> tools/_build/<circuit>.gates
  g++ -O2 -std=c++17 -pthread -I include -o tools/_build/plan_model tools/plan_model.cpp mpc_b200/csrc/plan.cpp
  tools/_build/plan_model tools/_build/aes_128.gates 96
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_circuit  # noqa: E402

out = os.path.join(ROOT, "tools", "_build")
os.makedirs(out, exist_ok=True)
for name in sys.argv[1:] or ["aes_128", "sha256", "sha512", "mul64", "add64", "aes_256"]:
    c = load_circuit(name)
    with open(os.path.join(out, name + ".gates"), "wb") as f:
        f.write(np.array([c.num_gates, c.num_wires, c.num_inputs, c.num_outputs], dtype=np.uint32).tobytes())
        f.write(np.ascontiguousarray(c.gates).tobytes())
    print(name, c.num_gates, "gates")
