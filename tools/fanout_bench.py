"""Multi-GPU inside the library: ONE process, one host thread, the batch fanned out over
the devices of gcb_set_devices.  Times, for 1 .. all devices of the box,

  host   gcb_garble_begin / gcb_eval_begin / gcb_job_wait on page-locked host buffers (each device DMAs its block)
  dev    gcb_garble_dev / gcb_eval_dev with GCB_FLAG_FANOUT: operands on device 0, inputs scattered and tables
         gathered with peer copies over NVLink
  iknp   gcb_iknp_receiver_expand + gcb_iknp_sender_expand on host buffers, 2^24 OTs split by chunk ranges

and checks every result against the one-device bytes.  One JSON line per configuration.

  python tools/fanout_bench.py [--batch 4096] [--circuit aes_128]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from conftest import load_circuit  # noqa: E402
from mpc_b200 import _lib  # noqa: E402
from mpc_b200 import circuit as gc  # noqa: E402
from mpc_b200.circuit import FLAG_FANOUT, GarbleEngine, host_alloc  # noqa: E402
from mpc_b200.circuit_io import LABEL_DTYPE, WIRE_DTYPE  # noqa: E402
from mpc_b200.ot import IKNPReceiver, IKNPSender, u_size  # noqa: E402

KEY = b"0123456789abcdef"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--circuit", default="aes_128")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    ndev = torch.cuda.device_count()
    circ = load_circuit(a.circuit)
    eng = GarbleEngine(circ)
    nin, nout, rows, n_and = circ.num_inputs, circ.num_outputs, circ.num_rows, circ.count(2)
    batch = a.batch
    rng = np.random.default_rng(5)
    h_r, h_l0 = host_alloc((batch,), LABEL_DTYPE), host_alloc((batch, nin), LABEL_DTYPE)
    h_tab, h_io = host_alloc((batch, rows), LABEL_DTYPE), host_alloc((batch, nin + nout), WIRE_DTYPE)
    h_in, h_out = host_alloc((batch, nin), LABEL_DTYPE), host_alloc((batch, nout), LABEL_DTYPE)
    h_r.view(np.uint8)[:] = rng.integers(0, 256, h_r.nbytes, dtype=np.uint8)
    h_l0.view(np.uint8).reshape(-1)[:] = rng.integers(0, 256, h_l0.nbytes, dtype=np.uint8)
    counts = [n for n in (1, 2, 4, 8) if n <= ndev]
    ref = {}
    for n in counts:
        gc.set_devices(list(range(n)))
        parts = [slice(k * batch // 16, (k + 1) * batch // 16) for k in range(16)]

        def step():
            gj = [eng.garble_begin(KEY, h_r[sl], h_l0[sl], h_tab[sl], h_io[sl]) for sl in parts]
            ej = []
            for j, sl in zip(gj, parts):
                j.wait()
                ej.append(eng.eval_begin(KEY, h_tab[sl], h_in[sl], h_out[sl]))
            for j in ej:
                j.wait()

        step()
        h_in[:] = h_io["l0"][:, :nin]
        step()
        t0 = time.perf_counter()
        for _ in range(a.reps):
            step()
        sec = (time.perf_counter() - t0) / a.reps
        sig = (hash(h_tab.tobytes()), hash(h_out.tobytes()))
        ref.setdefault("host", sig)
        print(json.dumps({"path": "host fan-out", "devices": n, "ms_per_step": sec * 1e3, "m_and_per_s": n_and * batch / sec / 1e6,
                          "same_bytes_as_one_device": sig == ref["host"],
                          "how": "one process, one host thread, 16 parts in flight, each part split over the devices"}), flush=True)
    # device-resident fan-out: operands on device 0
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    _lib.check(_lib.lib().gcb_set_device(0))

    def to_dev(x):
        return torch.from_numpy(np.ascontiguousarray(x).view(np.uint8).reshape(-1)).to(dev)

    d_key, d_r, d_l0 = to_dev(np.frombuffer(KEY, dtype=np.uint8)), to_dev(h_r), to_dev(h_l0)
    d_tab = torch.zeros(batch * rows * 16, dtype=torch.uint8, device=dev)
    d_io = torch.zeros(batch * (nin + nout) * 32, dtype=torch.uint8, device=dev)
    d_in, d_out = to_dev(h_in), torch.zeros(batch * nout * 16, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream(dev).cuda_stream
    for n in counts:
        gc.set_devices(list(range(n)))
        flags = FLAG_FANOUT if n > 1 else 0

        def step():
            eng.garble_dev(d_key, 16, 0, batch, d_r, d_l0, d_tab, d_io, stream=s, flags=flags)
            eng.eval_dev(d_key, 16, 0, batch, d_tab, d_in, d_out, stream=s, flags=flags)

        for _ in range(3):
            step()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            step()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / a.reps
        sig = (hash(d_tab.cpu().numpy().tobytes()), hash(d_out.cpu().numpy().tobytes()))
        ref.setdefault("dev", sig)
        print(json.dumps({"path": "device fan-out (GCB_FLAG_FANOUT)", "devices": n, "ms_per_step": ms, "m_and_per_s": n_and * batch / ms / 1e3,
                          "same_bytes_as_one_device": sig == ref["dev"], "gathered_gb_per_step": (n - 1) / n * batch * rows * 16 / 1e9,
                          "how": "operands on device 0; inputs scattered, tables gathered by cudaMemcpyPeerAsync behind the kernels"}), flush=True)
    _lib.check(_lib.lib().gcb_set_device(-1))
    # IKNP on host buffers
    n_ot = 1 << 24
    k0 = rng.integers(0, 2**63, (128, 2), dtype=np.uint64).view(LABEL_DTYPE).reshape(128)
    k1 = rng.integers(0, 2**63, (128, 2), dtype=np.uint64).view(LABEL_DTYPE).reshape(128)
    delta = rng.integers(0, 2**63, (1, 2), dtype=np.uint64).view(LABEL_DTYPE).reshape(1)
    choice = rng.integers(0, 2, n_ot, dtype=np.uint8)
    for n in counts:
        gc.set_devices(list(range(n)))
        rcv, snd = IKNPReceiver(k0, k1), IKNPSender(k0, delta)
        u, t = rcv.receive(choice)
        q = snd.send(u, n_ot)                        # warm both sides: staging arenas of every device, result arrays
        rcv.pos = snd.pos = 0
        t0 = time.perf_counter()
        u, t = rcv.receive(choice)
        q = snd.send(u, n_ot)
        sec = time.perf_counter() - t0
        sig = (hash(u.tobytes()), hash(t.tobytes()), hash(q.tobytes()))
        ref.setdefault("iknp", sig)
        print(json.dumps({"path": "IKNP 2^24 host fan-out (pageable buffers)", "devices": n, "ms": sec * 1e3,
                          "same_bytes_as_one_device": sig == ref["iknp"]}), flush=True)
    gc.set_devices([])


if __name__ == "__main__":
    main()
