"""BASELINE config 5 stand-in: a long streaming program garbled step by step for a batch of
independent instances, sharded over the GPUs of one box (one rank per GPU, no data-path
collective).  The real workload (ed25519 sign.mpcl, 8.45e8 gates, benchmarks.md:677-703) needs the
MPCL compiler (Go); this program has the same shape: large sub-circuits
(sha512.circ, 349,617 gates) chained through permanent wires, with a mul64.circ step every fourth
step.  302 steps exceed 10^8 gates per instance.

  python tools/stream_program.py [--steps 304] [--batch 148] [--check 2]
  torchrun --nproc-per-node N tools/stream_program.py ...

Every step: gcb_stream_garble_begin (device garble + record serialisation + D2H of the streams), then
gcb_seval_circuit on the same bytes (H2D + device eval) on the evaluator's own host thread.  Checks:
the first streams of instance 0 equal the oracle's byte for byte; the final state of `check`
instances decodes to the plaintext evaluation of the same program (gcb_circuit_compute)."""
import argparse
import json
import os
import queue
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

BIG_GATES, SMALL_GATES = 349617, 13675          # sha512.circ, mul64.circ


def steps_for_gates(gates: float) -> int:
    """Smallest step count whose program has at least `gates` gates per instance."""
    n = g = 0
    while g < gates:
        g += SMALL_GATES if n % 4 == 3 else BIG_GATES
        n += 1
    return n


def run_program(steps: int, batch: int, check: int = 2, warm: int = 4, serial: bool = False, no_eval: bool = False,
                rank: int = 0, oracle_steps: int = 2, sync=None):
    """Runs the program on the calling thread's device.  Returns a dict of figures for THIS rank (the caller takes
    the max over ranks).  `sync`: optional callable run at the start of the timed region (a barrier)."""
    import torch
    from conftest import load_circuit
    from mpc_b200.circuit import GarbleEngine, HostCircuit, StreamEval, Streaming
    from mpc_b200.circuit_io import LABEL_DTYPE

    big, small = load_circuit("sha512"), load_circuit("mul64")
    eb, es = GarbleEngine(big), GarbleEngine(small)
    hb, hs = HostCircuit(big), HostCircuit(small)
    rng = np.random.default_rng(1000 + rank)
    key = rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
    # permanent wires: state 0..511, then one 1024-bit block per big step
    n_big = sum(1 for k in range(steps) if k % 4 != 3)
    nin = 512 + 1024 * n_big
    ids = list(range(nin))
    r = rng.integers(0, 2**63, (batch, 2), dtype=np.uint64).view(LABEL_DTYPE).reshape(batch)
    l0 = rng.integers(0, 2**63, (batch, nin, 2), dtype=np.uint64).view(LABEL_DTYPE).reshape(batch, nin)
    st = Streaming(key, r, ids, l0)
    st.stream_buffers = 1 if serial else 4        # written by DMA / just completed / queued / being evaluated
    sev = None if no_eval else StreamEval(key, batch)
    bits = rng.integers(0, 2, (batch, nin)).astype(bool)
    if sev:
        w = st.get_inputs(ids)
        sev.set(ids, np.where(bits, w["l1"], w["l0"]).astype(LABEL_DTYPE))
        del w
    n_check = min(check, batch)
    l0_first = l0[0].copy() if rank == 0 and oracle_steps else None
    del l0
    state, next_id, blk = list(range(512)), nin, 0
    # plaintext wire values by permanent id, one row per checked instance
    plain = np.zeros((n_check, nin + 512 * n_big + 64 * (steps - n_big) + 64), dtype=np.uint8)
    plain[:, :nin] = bits[:n_check]
    gates = stream_bytes = 0
    t_g = t_e = 0.0
    first_streams = []
    work, eval_time, eval_err = queue.Queue(maxsize=1), [0.0], []
    dev_index = torch.cuda.current_device()

    def evaluator():
        from mpc_b200 import _lib
        _lib.check(_lib.lib().gcb_set_device(dev_index))
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
    if sev and not serial:
        worker = threading.Thread(target=evaluator, daemon=True)
        worker.start()
    # set-up outside the timed region: page-locking the output buffers (a few GB) is a one-time cost that a
    # program of thousands of steps does not see
    for _ in range(st.stream_buffers):
        st._stream_buffer(13 * big.num_gates + 16 * big.num_rows)
    st._turn = 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    warm = min(warm, max(steps - 1, 0))
    for k in range(steps):
        if k == warm and k:                             # the timed region starts here
            if not serial:
                st.garble_wait(0)
            if worker:
                if pending is not None:
                    work.put(pending)
                    pending = None
                work.join()
            torch.cuda.synchronize()
            if sync:
                sync()
            gates = stream_bytes = 0
            t_g = t_e = 0.0
            eval_time[0] = 0.0
            t0 = time.perf_counter()
        if k % 4 == 3:
            circ, eng, hc = small, es, hs
            ins = state[:128]
            outs = list(range(next_id, next_id + 64)); next_id += 64
        else:
            circ, eng, hc = big, eb, hb
            ins = list(range(512 + 1024 * blk, 512 + 1024 * (blk + 1))) + state
            blk += 1
            outs = list(range(next_id, next_id + 512)); next_id += 512
        ta = time.perf_counter()
        check_bytes = k < oracle_steps and rank == 0
        if serial:
            buf, _, _ = st.garble(eng, ins, outs)
        else:
            # one step in flight: the kernel of step k runs while the bytes of step k-1 cross PCIe
            buf = st.garble_begin(eng, ins, outs)
            st.garble_wait(0 if check_bytes else 1)       # the first streams are compared with the oracle
        tb = time.perf_counter()
        t_g += tb - ta
        if sev and serial:
            sev.circuit(buf, circ.num_gates, circ.num_wires, next_id)
            t_e += time.perf_counter() - tb
        elif sev:
            if pending is not None:
                work.put(pending)                                   # blocks while a step is queued
            pending = (buf, circ.num_gates, circ.num_wires, next_id)
        gates += circ.num_gates * batch
        stream_bytes += buf.shape[1] * batch
        if check_bytes:
            first_streams.append((circ, list(ins), list(outs), buf[0].tobytes()))
        if n_check:
            plain[:, outs] = hc.compute_bits(plain[:, ins])
        if circ is big:
            state = outs
        del buf
    if not serial:
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
    if sev and n_check:
        ow = st.get_inputs(state)
        got = sev.get(state)
        for i in range(n_check):
            dec = np.where(got[i] == ow[i]["l1"], 1, np.where(got[i] == ow[i]["l0"], 0, 2))
            ok &= bool(np.array_equal(dec, plain[i, state]))
    if rank == 0 and first_streams:
        from oracle import pyoracle as O          # checker only
        rand0 = np.concatenate([r[:1].view(np.uint64), l0_first.view(np.uint64).reshape(-1)]).astype(">u8").tobytes()
        ost = O.Streaming(key, rand0, ids)
        for circ, ins, outs, bytes0 in first_streams:
            ok &= ost.garble(circ, ins, outs) == bytes0
    total_gates = sum(SMALL_GATES if k % 4 == 3 else BIG_GATES for k in range(steps))
    return {"steps": steps, "timed_steps": steps - warm, "batch_per_gpu": batch, "gates_per_instance": total_gates,
            "timed_gates_per_instance": gates // batch, "timed_gates": gates, "stream_bytes": stream_bytes,
            "wall_s": wall, "garble_s": t_g, "eval_s": t_e,
            "evaluator": "none" if no_eval else "serial" if serial else "own host thread, one step behind",
            "checks_ok": bool(ok)}


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=steps_for_gates(1e8))
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
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mpc_b200 import _lib
    _lib.check(_lib.lib().gcb_set_device(local))
    res = run_program(a.steps, a.batch, a.check, a.warm, a.serial, a.no_eval, rank,
                      sync=(dist.barrier if dist else None))
    tt = torch.tensor([res["wall_s"], res["garble_s"], res["eval_s"], 0.0 if res["checks_ok"] else 1.0],
                      dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        wall, t_g, t_e, bad = tt.tolist()
        g = res["timed_gates"] * world
        print(json.dumps({"path": "stream_program", "n_gpus": world, **{k: res[k] for k in (
            "steps", "timed_steps", "batch_per_gpu", "gates_per_instance", "timed_gates_per_instance", "evaluator")},
            "total_gates": g, "stream_gb": res["stream_bytes"] * world / 1e9, "wall_s": wall, "garble_s": t_g, "eval_s": t_e,
            "m_gates_per_s": g / wall / 1e6, "m_gates_per_s_garble_only": g / max(t_g, 1e-9) / 1e6,
            "checks_ok": bad == 0.0,
            "note": "synthetic stand-in for ed25519 sign.mpcl (needs the Go MPCL compiler): sha512.circ steps chained "
                    "through permanent wires, a mul64.circ step every fourth"}), flush=True)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
