"""Device time of garble + eval of one circuit with the geometry the library picks (GCB_* environment overrides apply:
GCB_HOT_TEAMS, GCB_TEAMS, GCB_NT, GCB_ILP, GCB_TEAM_THREADS, GCB_STAGGER), best of 4, plus a digest of the bytes so that
runs with different geometry can be compared.   python tools/time_circuit.py sha256 1184 [keylen]"""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_circuit  # noqa: E402
from mpc_b200.circuit import GarbleEngine, select_labels_dev  # noqa: E402

name, batch = sys.argv[1], int(sys.argv[2])
klen = int(sys.argv[3]) if len(sys.argv) > 3 else 16
circ = load_circuit(name)
eng = GarbleEngine(circ)
nin, nout, rows = circ.num_inputs, circ.num_outputs, circ.num_rows
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(1)
rnd = lambda *s: torch.randint(0, 256, s, dtype=torch.uint8, generator=g).to(dev)
d_key, d_r, d_l0 = rnd(klen), rnd(batch, 16), rnd(batch, nin, 16)
d_bits = torch.randint(0, 2, (batch, nin), dtype=torch.uint8, generator=g).to(dev)
d_tab = torch.zeros((batch, rows, 16), dtype=torch.uint8, device=dev)
d_io = torch.zeros((batch, nin + nout, 32), dtype=torch.uint8, device=dev)
d_in = torch.zeros((batch, nin, 16), dtype=torch.uint8, device=dev)
d_out = torch.zeros((batch, nout, 16), dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
bg = be = 1e9
for it in range(5):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record()
    eng.garble_dev(d_key, klen, 0, batch, d_r, d_l0, d_tab, d_io, stream=s)
    e[1].record()
    select_labels_dev(d_io, nin + nout, d_bits, d_in, batch, nin, stream=s)
    e[2].record()
    eng.eval_dev(d_key, klen, 0, batch, d_tab, d_in, d_out, stream=s)
    e[3].record()
    torch.cuda.synchronize()
    if it:
        bg, be = min(bg, e[0].elapsed_time(e[1])), min(be, e[2].elapsed_time(e[3]))
i = eng.info
dig = hashlib.sha256(d_tab.cpu().numpy().tobytes() + d_out.cpu().numpy().tobytes()).hexdigest()[:16]
env = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("GCB_"))
print(f"{name} x {batch} [{env}]: {i.teams_per_sm} teams x {i.team_threads}, slots {i.num_slots} (hot {i.num_hot_slots}), passes {i.garble_passes}/{i.eval_passes}: "
      f"garble {bg:.3f} ms eval {be:.3f} ms -> {circ.count(2) * batch / (bg + be) / 1e3:.1f} M AND/s  digest {dig}", flush=True)
