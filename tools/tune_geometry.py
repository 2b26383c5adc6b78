"""Sweep the team geometry (GCB_ILP x GCB_TEAM_THREADS) of the gate kernels on one GPU.

  python tools/tune_geometry.py [circuit] [batch] [keylen] [ilp,team_threads,stagger[,tables[,teams]] ...]

Prints device time of garble and eval per configuration; every configuration's
tables and output labels are compared with the first one's (they must be
identical: the geometry is not allowed to change a single bit)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_circuit  # noqa: E402
from mpc_b200.circuit import GarbleEngine, select_labels_dev  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "aes_128"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    klen = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    circ = load_circuit(name)
    nin, nout, rows = circ.num_inputs, circ.num_outputs, circ.num_rows
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(1)
    d_key = torch.randint(0, 256, (klen,), dtype=torch.uint8, generator=g).to(dev)
    d_r = torch.randint(0, 256, (batch, 16), dtype=torch.uint8, generator=g).to(dev)
    d_l0 = torch.randint(0, 256, (batch, nin, 16), dtype=torch.uint8, generator=g).to(dev)
    d_bits = torch.randint(0, 2, (batch, nin), dtype=torch.uint8, generator=g).to(dev)
    d_tab = torch.zeros((batch, rows, 16), dtype=torch.uint8, device=dev)
    d_io = torch.zeros((batch, nin + nout, 32), dtype=torch.uint8, device=dev)
    d_in = torch.zeros((batch, nin, 16), dtype=torch.uint8, device=dev)
    d_out = torch.zeros((batch, nout, 16), dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    ref = None
    configs = [(None, None, 0)] + [(i, t, sg) for i in (1, 2, 4) for t in (32, 64, 96, 128, 160, 256)
                                   for sg in (0, 100000)]
    if len(sys.argv) > 4:
        configs = [tuple(int(x) for x in c.split(",")) for c in sys.argv[4:]]
    print(f"{name} batch={batch} keylen={klen}: slots/rows see plan; times in ms")
    for cfg in configs:
        ilp, tt, sg = cfg[:3]
        nt = cfg[3] if len(cfg) > 3 else None
        teams = cfg[4] if len(cfg) > 4 else None
        for k in ("GCB_ILP", "GCB_TEAM_THREADS", "GCB_STAGGER", "GCB_NT", "GCB_TEAMS"):
            os.environ.pop(k, None)
        if nt:
            os.environ["GCB_NT"] = str(nt)
        if teams:
            os.environ["GCB_TEAMS"] = str(teams)
        if ilp:
            os.environ["GCB_ILP"] = str(ilp)
            os.environ["GCB_TEAM_THREADS"] = str(tt)
            os.environ["GCB_STAGGER"] = str(sg)
        eng = GarbleEngine(circ)
        info = eng.info
        if ilp and info.team_threads != tt:
            continue                                  # geometry not realisable
        d_tab.zero_(); d_out.zero_()
        best_g, best_e = 1e9, 1e9
        for it in range(4):
            e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
            e0.record()
            eng.garble_dev(d_key, klen, 0, batch, d_r, d_l0, d_tab, d_io, stream=s)
            e1.record()
            select_labels_dev(d_io, nin + nout, d_bits, d_in, batch, nin, stream=s)
            e2.record()
            eng.eval_dev(d_key, klen, 0, batch, d_tab, d_in, d_out, stream=s)
            e3.record()
            torch.cuda.synchronize()
            if it:
                best_g, best_e = min(best_g, e0.elapsed_time(e1)), min(best_e, e2.elapsed_time(e3))
        eng.garble_dev(d_key, klen, 0, batch, d_r, d_l0, d_tab, d_io, stream=s)
        eng.eval_dev(d_key, klen, 0, batch, d_tab, d_in, d_out, stream=s)
        torch.cuda.synchronize()
        sig = (hash(d_tab.cpu().numpy().tobytes()), hash(d_out.cpu().numpy().tobytes()))
        if ref is None:
            ref = sig
        ok = "same" if sig == ref else "DIFFERENT"
        n_and = circ.count(2)
        print(f"ilp={ilp} tt={tt} stagger={sg} nt={nt}: teams={info.teams_per_sm} x {info.team_threads} slots={info.num_slots} "
              f"garble={best_g:.3f} eval={best_e:.3f} total={best_g + best_e:.3f} "
              f"-> {n_and * batch / (best_g + best_e) / 1e3:.1f} M AND/s  [{ok}]", flush=True)


if __name__ == "__main__":
    main()
