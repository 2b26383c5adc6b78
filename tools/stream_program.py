"""BASELINE config 5 stand-in: a long streaming program garbled step by step for a batch of
independent instances, sharded over the GPUs of one box (one rank per GPU, no data-path
collective).  The real workload (ed25519 sign.mpcl, 8.45e8 gates) needs the MPCL compiler (Go);
this program has the same shape: large sub-circuits (sha512.circ, 349,617 gates) chained through
permanent wires, with a mul64.circ step every fourth step.

  python tools/stream_program.py [--steps 40] [--batch 148] [--check 2]
  torchrun --nproc-per-node N tools/stream_program.py ...

Every step: gcb_stream_garble (device garble + record serialisation + D2H of the streams), then
gcb_seval_circuit on the same bytes (H2D + device eval).  Checks: instance 0's first streams
equal the oracle's byte for byte; the final state of `--check` instances decodes to the plaintext
evaluation of the same program."""
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
from conftest import load_circuit  # noqa: E402
from mpc_b200 import _lib  # noqa: E402
from mpc_b200.circuit import GarbleEngine, StreamEval, Streaming  # noqa: E402
from mpc_b200.circuit_io import LABEL_DTYPE  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--batch", type=int, default=148)
    ap.add_argument("--check", type=int, default=2)
    ap.add_argument("--no-eval", action="store_true")
    ap.add_argument("--warm", type=int, default=4, help="leading steps outside the timed region (both sub-circuits "
                    "occur in them: plans are compiled and recovered, staging buffers allocated -- one-time work)")
    ap.add_argument("--serial", action="store_true", help="evaluate step k before garbling step k+1 (default: "
                    "the evaluator runs on its own host thread one step behind the garbler, as a second party would)")
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.check(_lib.lib().gcb_set_device(local))
    big, small = load_circuit("sha512"), load_circuit("mul64")
    eb, es = GarbleEngine(big), GarbleEngine(small)
    rng = np.random.default_rng(1000 + rank)
    key = rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
    batch = a.batch
    # permanent wires: state 0..511, then one 1024-bit block per big step
    n_big = sum(1 for k in range(a.steps) if k % 4 != 3)
    nin = 512 + 1024 * n_big
    ids = list(range(nin))
    r = rng.integers(0, 2**63, (batch, 2), dtype=np.uint64).view(LABEL_DTYPE).reshape(batch)
    l0 = rng.integers(0, 2**63, (batch, nin, 2), dtype=np.uint64).view(LABEL_DTYPE).reshape(batch, nin)
    st = Streaming(key, r, ids, l0)
    st.stream_buffers = 1 if a.serial else 4        # written by DMA / just completed / queued / being evaluated
    sev = None if a.no_eval else StreamEval(key, batch)
    bits = rng.integers(0, 2, (batch, nin)).astype(bool)
    if sev:
        w = st.get_inputs(ids)
        sev.set(ids, np.where(bits, w["l1"], w["l0"]).astype(LABEL_DTYPE))
        del w
    state, next_id, blk = list(range(512)), nin, 0
    plain = [bits[i].astype(np.uint8).tolist() for i in range(min(a.check, batch))]      # plaintext wire values by id
    plain = [dict(enumerate(p)) for p in plain]
    gates = stream_bytes = 0
    t_g = t_e = t_init = t_dev = 0.0
    first_streams = []
    import queue
    import threading
    work, eval_time, eval_err = queue.Queue(maxsize=1), [0.0], []

    def evaluator():
        _lib.check(_lib.lib().gcb_set_device(local))
        while True:
            item = work.get()
            if item is None:
                work.task_done()
                return
            t1 = time.perf_counter()
            try:
                sev.circuit(*item)
            except Exception as e:                                  # surfaced after the loop
                eval_err.append(e)
            eval_time[0] += time.perf_counter() - t1
            work.task_done()

    worker, pending = None, None
    if sev and not a.serial:
        worker = threading.Thread(target=evaluator, daemon=True)
        worker.start()
    # set-up outside the timed region: page-locking the output buffers (a few GB) is a one-time cost that a
    # program of thousands of steps does not see
    for _ in range(st.stream_buffers):
        st._stream_buffer(13 * big.num_gates + 16 * big.num_rows)
    st._turn = 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(a.steps):
        if k == a.warm and k:                             # the timed region starts here
            if not a.serial:
                st.garble_wait(0)
            if worker:
                if pending is not None:
                    work.put(pending)
                    pending = None
                work.join()
            torch.cuda.synchronize()
            gates = stream_bytes = 0
            t_g = t_e = t_init = t_dev = 0.0
            eval_time[0] = 0.0
            t0 = time.perf_counter()
        if k % 4 == 3:
            circ, eng = small, es
            ins = state[:128]
            outs = list(range(next_id, next_id + 64)); next_id += 64
        else:
            circ, eng = big, eb
            ins = list(range(512 + 1024 * blk, 512 + 1024 * (blk + 1))) + state
            blk += 1
            outs = list(range(next_id, next_id + 512)); next_id += 512
        ta = time.perf_counter()
        if a.serial:
            buf, ns_i, ns_g = st.garble(eng, ins, outs)
            t_init += ns_i / 1e9; t_dev += ns_g / 1e9
        else:
            # one step in flight: the kernel of step k runs while the bytes of step k-1 cross PCIe
            buf = st.garble_begin(eng, ins, outs)
            st.garble_wait(0 if k < 2 and rank == 0 else 1)       # the first streams are compared with the oracle
        tb = time.perf_counter()
        t_g += tb - ta
        if sev and a.serial:
            sev.circuit(buf, circ.num_gates, circ.num_wires, next_id)
            t_e += time.perf_counter() - tb
        elif sev:
            if pending is not None:
                work.put(pending)                                   # blocks while a step is queued
            pending = (buf, circ.num_gates, circ.num_wires, next_id)
        gates += circ.num_gates * batch
        stream_bytes += buf.shape[1] * batch
        if k < 2 and rank == 0:
            first_streams.append((circ, list(ins), list(outs), buf[0].tobytes()))
        for p in plain:
            ob = circ.compute_bits([p[i] for i in ins])
            for i, o in zip(outs, ob.tolist()):
                p[i] = o
        if circ is big:
            state = outs
        del buf
    if not a.serial:
        st.garble_wait(0)
    if worker:
        if pending is not None:
            work.put(pending)
        work.put(None)
        worker.join()
        t_e = eval_time[0]
        if eval_err:
            raise eval_err[0]
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ok = True
    if sev and plain:
        ow = st.get_inputs(state)
        got = sev.get(state)
        for i, p in enumerate(plain):
            dec = np.where(got[i] == ow[i]["l1"], 1, np.where(got[i] == ow[i]["l0"], 0, 2))
            ok &= bool(np.array_equal(dec, np.array([p[j] for j in state])))
    if rank == 0 and first_streams:
        from oracle import pyoracle as O          # checker only
        rand0 = np.concatenate([r[:1].view(np.uint64), l0[0].view(np.uint64).reshape(-1)]).astype(">u8").tobytes()
        ost = O.Streaming(key, rand0, ids)
        for circ, ins, outs, bytes0 in first_streams:
            ok &= ost.garble(circ, ins, outs) == bytes0
    tt = torch.tensor([wall, t_g, t_e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        wall, t_g, t_e = tt.tolist()
        print(json.dumps({"path": "stream_program", "n_gpus": world, "steps": a.steps, "timed_steps": a.steps - a.warm, "batch_per_gpu": batch,
                          "gates_per_instance": gates // batch, "total_gates": gates * world,
                          "stream_gb": stream_bytes * world / 1e9, "wall_s": wall, "garble_s": t_g, "garble_init_s": t_init, "garble_device_s": t_dev, "eval_s": t_e,
                          "m_gates_per_s": gates * world / wall / 1e6, "m_gates_per_s_garble_only": gates * world / t_g / 1e6,
                          "evaluator": "serial" if a.serial else "own host thread, one step behind", "checks_ok": ok}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
