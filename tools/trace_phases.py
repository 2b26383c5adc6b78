"""Where one team's time goes: per-phase clock64 timestamps of block 0 / team 0 of the
garble kernel (XOR run, barrier, cipher level).  python tools/trace_phases.py [circuit] [batch]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_circuit  # noqa: E402
from mpc_b200 import _lib  # noqa: E402
from mpc_b200.circuit import GarbleEngine  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "aes_128"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
circ = load_circuit(name)
eng = GarbleEngine(circ)
dev = torch.device("cuda:0")
nin, nout, rows = circ.num_inputs, circ.num_outputs, circ.num_rows
rnd = lambda *s: torch.randint(0, 256, s, dtype=torch.uint8, device=dev)
key, r, l0 = rnd(16), rnd(batch, 16), rnd(batch, nin, 16)
tab = torch.empty((batch, rows, 16), dtype=torch.uint8, device=dev)
io = torch.empty((batch, nin + nout, 32), dtype=torch.uint8, device=dev)
L = _lib.lib()
L.gcb_debug_set_trace.argtypes = [C.c_void_p]
nph = 100000
trace = torch.zeros(4 * nph, dtype=torch.int64, device=dev)
for it in range(2):
    eng.garble_dev(key, 16, 0, batch, r, l0, tab, io)
torch.cuda.synchronize()
L.gcb_debug_set_trace(trace.data_ptr())
eng.garble_dev(key, 16, 0, batch, r, l0, tab, io)
torch.cuda.synchronize()
L.gcb_debug_set_trace(None)
t = trace.cpu().numpy().reshape(-1, 4)
n = int(np.count_nonzero(t[:, 0]))
t = t[:n]
print(f"{name}: {n} phases traced; teams={eng.info.teams_per_sm} x {eng.info.team_threads}")
xor = np.where(t[:, 1] > 0, t[:, 1] - t[:, 0], 0)
bar1 = np.where(t[:, 1] > 0, t[:, 2] - t[:, 1], 0)
ciph = np.where(t[:, 3] > 0, t[:, 3] - t[:, 2], 0)
nxt = np.zeros(n, dtype=np.int64)
nxt[:-1] = t[1:, 0] - np.where(t[:-1, 3] > 0, t[:-1, 3], t[:-1, 2])
total = t[-1, 3 if t[-1, 3] > 0 else 2] - t[0, 0]
print(f"total {total} cycles; XOR runs {xor.sum()} ({100 * xor.sum() / total:.1f}%), barrier after XOR {bar1.sum()} "
      f"({100 * bar1.sum() / total:.1f}%), cipher levels {ciph.sum()} ({100 * ciph.sum() / total:.1f}%), "
      f"barrier after cipher + phase top {nxt.sum()} ({100 * nxt.sum() / total:.1f}%)")
print("per phase mean: xor %.0f  bar %.0f  cipher %.0f  next %.0f" % (xor.mean(), bar1.mean(), ciph.mean(), nxt.mean()))
for i in list(range(min(n, 12))):
    print(i, int(xor[i]), int(bar1[i]), int(ciph[i]), int(nxt[i]))
