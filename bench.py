#!/usr/bin/env python
"""Headline benchmark: million AND-gates/s garbled + evaluated on the AES-128
Bristol circuit, batch 4096 per GPU (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl gcb|reference]

One step = garble the whole batch, select the evaluator's input labels,
evaluate the whole batch, decode the outputs.  `value` is measured with every
input resident in HBM; `e2e` goes through the host-pointer C ABI (gcb_garble /
gcb_eval on pinned host buffers, copies inside the timed region).  The
reference arm times the CPU oracle (the C restatement of the reference's Go
loops, AES-NI) on all host cores.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "million AND-gates/sec garble+eval (AES-128 circuit)"
UNIT = "M AND-gates/s"
E2E_PARTS = 16
E2E_WORKERS = 2
KEY = b"0123456789abcdef"            # circuit/garble_bench_test.go:34
CIRCUIT_DIR = os.path.join(ROOT, "tests", "golden", "circuits")
# --circuit: the headline workload (aes_128, BASELINE.json configs[1]) or the second circuit BASELINE's
# target names (sha256: 22,573 AND; batch = 8 resident instances x 148 SMs)
WORKLOADS = {"aes_128": ("AES-128 circuit", 4096), "sha256": ("SHA-256 circuit", 1184)}


def load_circuit(name: str = "aes_128"):
    from mpc_b200.circuit_io import Circuit
    return Circuit.load_npz(os.path.join(CIRCUIT_DIR, name + ".npz"), name)


def synthetic_inputs(circ, batch: int, rank: int):
    """R and L0 draws per instance from DRBG("aes128/<global instance>") in the
    reference's reader order; evaluator plaintext: key 000102..0f, block = instance index."""
    from mpc_b200.drbg import DRBG
    from util import rand_to_labels
    nin = circ.num_inputs
    rand = np.empty((batch, 16 * (1 + nin)), dtype=np.uint8)
    tag = "aes128" if circ.name == "aes_128" else circ.name
    for i in range(batch):
        rand[i] = DRBG(f"{tag}/{rank * batch + i}").array(16 * (1 + nin))
    r, l0 = rand_to_labels(rand, nin)
    if circ.name != "aes_128":                       # other circuits: seeded random plaintext inputs
        bits = np.random.default_rng(1234 + rank).integers(0, 2, (batch, nin)).astype(np.uint8)
        return rand, r, l0, bits
    pt_key = int.from_bytes(bytes(range(16)), "big")
    bits = np.zeros((batch, nin), dtype=np.uint8)
    kb = np.array([(pt_key >> b) & 1 for b in range(128)], dtype=np.uint8)
    idx = np.arange(rank * batch, (rank + 1) * batch, dtype=np.uint64)
    bits[:, :128] = kb
    for b in range(64):
        bits[:, 128 + b] = (idx >> np.uint64(b)) & np.uint64(1)
    return rand, r, l0, bits


def algorithmic_bytes(circ):
    """per-instance bytes that must cross HBM."""
    nin, nout, rows = circ.num_inputs, circ.num_outputs, circ.num_rows
    garble = 16 * (1 + nin) + 16 * rows + 32 * (nin + nout)
    evalb = 16 * rows + 16 * nin + 16 * nout
    return garble, evalb


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML during the timed region."""

    def __init__(self, index: int):
        self.index, self.sm, self.reasons, self.mx = index, [], set(), None
        self._stop = threading.Event()
        self.t = None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                     "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                     "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}

            def loop():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                        for k, bit in names.items():
                            if r & bit:
                                self.reasons.add(k)
                    except Exception:
                        pass
                    time.sleep(0.005)

            self.t = threading.Thread(target=loop, daemon=True)
            self.t.start()
        except Exception as e:               # NVML missing: say so instead of inventing a clock
            self.reasons.add(f"nvml unavailable: {e}")

    def stop(self):
        self._stop.set()
        if self.t:
            self.t.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                "samples": len(self.sm), "reasons": sorted(self.reasons)}


def cpu_arm(circ, threads: int, sample: int, reps: int):
    """Garble + eval of `sample` instances, `reps` times, on `threads` host threads (the oracle)."""
    from mpc_b200.drbg import garble_inputs
    from oracle import pyoracle as O
    from util import select
    _, rand = garble_inputs("cpu", sample, circ.num_inputs, 0)
    bits = np.random.default_rng(0).integers(0, 2, (sample, circ.num_inputs), dtype=np.uint8)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        _, tables, io = O.garble_batch(circ, KEY, rand, threads=threads)
        t1 = time.perf_counter()
        inl = select(io[:, : circ.num_inputs], bits)
        t2 = time.perf_counter()
        O.eval_batch(circ, KEY, tables, inl, threads=threads)
        t3 = time.perf_counter()
        times.append((t1 - t0) + (t3 - t2))
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    circ = load_circuit(args.circuit)
    n_and = circ.count(2)
    batch = args.batch or WORKLOADS[args.circuit][1]
    threads = os.cpu_count() or 1
    sample = max(threads * 8, 64)
    # size one step to roughly 1-2 s of wall time
    t = cpu_arm(circ, threads, sample, 1)[0]
    sample = int(min(batch, max(sample, sample * 1.0 / max(t, 1e-3))))
    for _ in range(args.warmup):
        cpu_arm(circ, threads, sample, 1)
    times = cpu_arm(circ, threads, sample, args.steps)
    total = sum(times)
    value = n_and * sample * args.steps / total / 1e6
    line = {
        "impl": "reference", "metric": METRIC.replace("AES-128 circuit", WORKLOADS[args.circuit][0]), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64/u8 (AES-NI)",
        "data": "synthetic",
        "config": {"workload": f"{circ.name}.circ ({n_and} AND, {circ.count(4)} INV, {circ.count(0) + circ.count(1)} XOR) "
                               f"garble+eval, batch {batch} per GPU",
                   "batch_per_gpu": batch, "key": "shared 16-byte (AES-128)",
                   "sample_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} instances per step x {args.steps} steps, C oracle (AES-NI), "
                                   f"{threads} pthreads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(index: int):
    """One process per GPU: run this rank's host threads, and so first-touch its pinned staging buffers, on
    the NUMA node the GPU hangs off (sysfs numa_node / local_cpulist of the PCI function).  Without it the
    host-pointer path of an 8-rank run crosses the socket interconnect for half of its DMA traffic.
    Returns a short description for the JSON line; silently does nothing where sysfs has no answer."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()
        if len(bdf.split(":")[0]) == 8:                       # NVML prints an 8-digit domain, sysfs a 4-digit one
            bdf = bdf[4:]
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(base + "/numa_node").read())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if node < 0 or not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception:
        return None


def run_gcb(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        # convenience: `python bench.py --gpus N` re-launches itself under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from mpc_b200 import _lib
    from mpc_b200.circuit import GarbleEngine, decode_bits_dev, select_labels_dev
    from mpc_b200.circuit_io import LABEL_DTYPE, WIRE_DTYPE

    _lib.check(_lib.lib().gcb_set_device(local))
    circ = load_circuit(args.circuit)
    eng = GarbleEngine(circ)
    nin, nout, rows, n_and = circ.num_inputs, circ.num_outputs, circ.num_rows, circ.count(2)
    batch = args.batch or WORKLOADS[args.circuit][1]
    rand, r, l0, bits = synthetic_inputs(circ, batch, rank)

    def to_dev(a):
        return torch.from_numpy(a.view(np.uint8).reshape(a.shape + (-1,)) if a.dtype.fields else a).to(dev)

    d_key = torch.frombuffer(bytearray(KEY), dtype=torch.uint8).to(dev)
    d_r, d_l0, d_bits = to_dev(r), to_dev(l0), to_dev(bits)
    d_tab = torch.empty((batch, rows, 16), dtype=torch.uint8, device=dev)
    d_io = torch.empty((batch, nin + nout, 32), dtype=torch.uint8, device=dev)
    d_in = torch.empty((batch, nin, 16), dtype=torch.uint8, device=dev)
    d_out = torch.empty((batch, nout, 16), dtype=torch.uint8, device=dev)
    d_obits = torch.empty((batch, nout), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    s = stream.cuda_stream

    ev = lambda: torch.cuda.Event(enable_timing=True)
    kern = {"garble": [], "eval": []}

    def step(timed: bool):
        e0, e1, e2, e3 = (ev(), ev(), ev(), ev()) if timed else (None,) * 4
        if timed: e0.record(stream)
        eng.garble_dev(d_key, 16, 0, batch, d_r, d_l0, d_tab, d_io, stream=s)
        if timed: e1.record(stream)
        select_labels_dev(d_io, nin + nout, d_bits, d_in, batch, nin, stream=s)
        if timed: e2.record(stream)
        eng.eval_dev(d_key, 16, 0, batch, d_tab, d_in, d_out, stream=s)
        if timed: e3.record(stream)
        if timed:
            kern["garble"].append((e0, e1)); kern["eval"].append((e2, e3))

    # output wires are a strided view of io_wires: decode takes the wire stride
    def decode():
        base = d_io.data_ptr() + nin * 32
        decode_bits_dev(base, nin + nout, d_out, d_obits, batch, nout, stream=s)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(False); decode()
    barrier()
    # correctness of what is being timed: decoded outputs equal OpenSSL AES of the instance index
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
    ob = d_obits.cpu().numpy()
    enc = Cipher(algorithms.AES(bytes(range(16))), modes.ECB()).encryptor()
    for i in (0, 1, batch // 2, batch - 1):
        if args.circuit == "aes_128":
            want = int.from_bytes(enc.update((rank * batch + i).to_bytes(16, "big")), "big")
            got = sum(int(b) << k for k, b in enumerate(ob[i]))
            assert got == want, f"instance {i}: decoded output is not AES(key, index)"
        else:                                        # plaintext evaluation of the same circuit file
            assert np.array_equal(ob[i], circ.compute_bits(bits[i].tolist())), f"instance {i}: decoded output is wrong"

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t_start, t_end = ev(), ev()
    t_start.record(stream)
    for _ in range(args.steps):
        step(True); decode()
    t_end.record(stream)
    barrier()
    ms = t_start.elapsed_time(t_end)
    clocks = sampler.stop() if rank == 0 else None
    g_ms = float(np.mean([a.elapsed_time(b) for a, b in kern["garble"]]))
    e_ms = float(np.mean([a.elapsed_time(b) for a, b in kern["eval"]]))

    e2e_s = float("nan")
    if not args.no_e2e:
        # ---- e2e: host-pointer C ABI, pinned host buffers, copies inside the timed region
        L = _lib.lib()

        def pinned(shape, dtype):
            n = int(np.prod(shape)) * np.dtype(dtype).itemsize
            p = L.gcb_host_alloc(max(n, 1))
            assert p, "gcb_host_alloc failed"
            import ctypes as C
            arr = np.frombuffer((C.c_uint8 * n).from_address(p), dtype=dtype).reshape(shape)
            return arr, p

        h_r, p1 = pinned((batch,), LABEL_DTYPE); h_r[:] = r
        h_l0, p2 = pinned((batch, nin), LABEL_DTYPE); h_l0[:] = l0
        h_tab, p3 = pinned((batch, rows), LABEL_DTYPE)
        h_io, p4 = pinned((batch, nin + nout), WIRE_DTYPE)
        h_in, p5 = pinned((batch, nin), LABEL_DTYPE)
        h_out, p6 = pinned((batch, nout), LABEL_DTYPE)

        # The call a user makes: gcb_garble / gcb_eval on host buffers.  The batch goes through in
        # E2E_PARTS sub-batches on two small thread pools -- the garbler's tables of part i stream
        # back (D2H) while the evaluator's tables of earlier parts stream in (H2D), as two parties
        # would; two workers per side hide the start-up latency of each blocking call.
        from concurrent.futures import ThreadPoolExecutor
        # Sub-batches: as many as keep the PCIe time of one part well above the latency of garbling one
        # instance (a part's kernel cannot finish sooner than that): ~1.2 us per dependency step of the plan.
        lat_ms = eng.info.num_steps * 1.2e-3
        pcie_ms = batch * (16 * rows + 32 * (nin + nout)) / 50e6
        n_parts = int(max(2, min(E2E_PARTS, pcie_ms / (2 * lat_ms))))
        parts = [slice(k * batch // n_parts, (k + 1) * batch // n_parts) for k in range(n_parts)]

        def garble_part(sl):
            _lib.check(L.gcb_set_device(local))
            eng.garble_batch(KEY, h_r[sl], h_l0[sl], tables=h_tab[sl], io_wires=h_io[sl])

        def eval_part(sl):
            _lib.check(L.gcb_set_device(local))
            eng.eval_batch(KEY, h_tab[sl], h_in[sl], out_labels=h_out[sl])

        gpool, epool = ThreadPoolExecutor(E2E_WORKERS), ThreadPoolExecutor(E2E_WORKERS)

        def e2e_step():
            gfs = [gpool.submit(garble_part, sl) for sl in parts]
            efs = []
            for f, sl in zip(gfs, parts):
                f.result()                      # re-raises worker exceptions
                efs.append(epool.submit(eval_part, sl))
            for f in efs:
                f.result()

        e2e_step()
        h_in[:] = np.where(bits.astype(bool), h_io["l1"][:, :nin], h_io["l0"][:, :nin])
        e2e_steps = max(2, min(args.steps, 10))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        ok_e2e = h_out.tobytes() == d_out.cpu().numpy().tobytes() and h_tab.tobytes() == d_tab.cpu().numpy().tobytes()
        assert ok_e2e, "host-pointer path and device-resident path disagree"
        gpool.shutdown(); epool.shutdown()
        for p in (p1, p2, p3, p4, p5, p6):
            L.gcb_host_free(p)

    # max over ranks
    tt = torch.tensor([ms, e2e_s * 1e3, g_ms, e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms, e2e_ms, g_ms, e_ms = tt.tolist()

    if rank == 0:
        gb, eb = algorithmic_bytes(circ)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = gb * batch / (g_ms * 1e-3) / 1e9
        traffic = None                              # DRAM bytes per launch from the committed ncu capture
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["garble_kernel"]["dram_bytes"]
        except Exception:
            pass
        total_and = n_and * batch * world
        value = total_and * args.steps / (ms * 1e-3) / 1e6
        cores = os.cpu_count() or 1
        # CPU baseline: bounded sample on the host cores (rank 0, N=1 only)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sample = max(cores * 8, 64)
            t1 = cpu_arm(circ, cores, sample, 1)[0]
            sample = int(min(batch, max(sample, sample * 1.5 / max(t1, 1e-3))))
            ts = cpu_arm(circ, cores, sample, 3)
            cpu = {"value": n_and * sample / min(ts) / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{sample} instances garble+eval, best of 3, C oracle (AES-NI), {cores} pthreads"}
        line = {
            "metric": METRIC.replace("AES-128 circuit", WORKLOADS[args.circuit][0]), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32 (AES T-tables, 128-bit label XOR)",
            "data": "synthetic",
            "config": {"workload": f"{circ.name}.circ ({n_and} AND, {circ.count(4)} INV, {circ.count(0) + circ.count(1)} XOR) "
                                   f"garble+eval, batch {batch} per GPU",
                       "batch_per_gpu": batch, "key": "shared 16-byte (AES-128)",
                       "l2": f"tables are {batch * rows * 16 // 1000000} MB per step, larger than L2; no flush needed",
                       "kernel_ms": {"garble": g_ms, "eval": e_ms}},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes": gb * batch,
                         "kernel": f"garble_kernel<10,PLAIN> ({eng.info.teams_per_sm} teams x {eng.info.team_threads} threads per SM)",
                         "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback",
                         "note": "bound by the shared-memory pipe (AES T-table lookups, 83% busy in the ncu capture), "
                                 "not HBM: there is no AES instruction on the GPU; see DESIGN.md and profiles/"},
            "cpu_baseline": cpu,
            "e2e": None if args.no_e2e else {"value": total_and / (e2e_ms * 1e-3) / 1e6, "unit": UNIT,
                    "h2d_bytes_per_step": int(world * batch * (16 * (1 + nin) + 16 * rows + 16 * nin)),
                    "d2h_bytes_per_step": int(world * batch * (16 * rows + 32 * (nin + nout) + 16 * nout)),
                    "ms_per_step": e2e_ms,
                    "how": f"gcb_garble + gcb_eval on pinned host buffers, {n_parts} sub-batches, "
                           f"{E2E_WORKERS} garbler + {E2E_WORKERS} evaluator host threads"
                           + (f", ranks bound to their GPU's NUMA node ({numa['cpus']} cpus)" if numa else "")},
            "gpu_launches": 4 * args.steps,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gcb", choices=["gcb", "reference"])
    ap.add_argument("--circuit", default="aes_128", choices=sorted(WORKLOADS), help="aes_128 = the headline workload")
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU (default: 4096 for aes_128, 1184 for sha256)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: device-resident loop only")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gcb(args)


if __name__ == "__main__":
    main()
