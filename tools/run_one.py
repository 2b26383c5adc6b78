"""One garble (+eval) launch of a circuit, for ncu: python tools/run_one.py sha256 592 [reps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_circuit
from mpc_b200.circuit import GarbleEngine
name, batch = sys.argv[1], int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
circ = load_circuit(name); eng = GarbleEngine(circ); dev = torch.device("cuda:0")
nin, nout, rows = circ.num_inputs, circ.num_outputs, circ.num_rows
rnd = lambda *s: torch.randint(0, 256, s, dtype=torch.uint8, device=dev)
key, r, l0 = rnd(16), rnd(batch, 16), rnd(batch, nin, 16)
tab = torch.empty((batch, rows, 16), dtype=torch.uint8, device=dev); io = torch.empty((batch, nin + nout, 32), dtype=torch.uint8, device=dev)
out = torch.empty((batch, nout, 16), dtype=torch.uint8, device=dev)
for _ in range(reps):
    eng.garble_dev(key, 16, 0, batch, r, l0, tab, io)
    eng.eval_dev(key, 16, 0, batch, tab, l0, out)
torch.cuda.synchronize()
