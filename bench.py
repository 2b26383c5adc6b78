#!/usr/bin/env python
"""Headline benchmark: million AND-gates/s garbled + evaluated on the AES-128
Bristol circuit, batch 4096 per GPU (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl gcb|reference] [--no-extra]

One step = garble the whole batch, select the evaluator's input labels,
evaluate the whole batch, decode the outputs.  `value` is measured with every
input resident in HBM; `e2e` goes through the host-pointer C ABI
(gcb_garble_begin / gcb_eval_begin / gcb_job_wait on page-locked host buffers,
one host thread, copies inside the timed region).  The reference arm times the
CPU oracle (the C restatement of the reference's Go loops, AES-NI) on all host
cores.  Prints ONE JSON line on rank 0.

`extra` in the same line carries the other BASELINE.json configurations, each
timed at this N with its own algorithmic bytes and HBM fraction: aes_128 with
per-instance 32-byte keys (the production key shape), sha256.circ x 2368,
IKNP 2^24 (sender + receiver), one streaming sha256 step, the >= 10^8-gate
streaming program (config 5 stand-in), a copy-only PCIe probe (the floor of the
e2e figure on this box with N ranks copying at once) and, at N = 1, the latency
of the unchanged batch = 1 call next to one CPU thread.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

# more hardware work queues than the default 8, before the CUDA context exists: the job API keeps 8 library
# streams busy beside the caller's own
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

METRIC = "million AND-gates/sec garble+eval (AES-128 circuit)"
UNIT = "M AND-gates/s"
E2E_PARTS = 8
KEY = b"0123456789abcdef"            # circuit/garble_bench_test.go:34
CIRCUIT_DIR = os.path.join(ROOT, "tests", "golden", "circuits")
# --circuit: the headline workload (aes_128, BASELINE.json configs[1]) or the second circuit BASELINE's
# target names (sha256: 22,573 AND; batch = 16 resident instances x 148 SMs)
WORKLOADS = {"aes_128": ("AES-128 circuit", 4096), "sha256": ("SHA-256 circuit", 2368)}


def load_circuit(name: str = "aes_128"):
    from mpc_b200.circuit_io import Circuit
    return Circuit.load_npz(os.path.join(CIRCUIT_DIR, name + ".npz"), name)


def workload_name(circ, batch: int) -> str:
    return (f"{circ.name}.circ ({circ.count(2)} AND, {circ.count(4)} INV, {circ.count(0) + circ.count(1)} XOR) "
            f"garble+eval, batch {batch} per GPU")


def config_of(circ, batch: int) -> dict:
    """The `config` object: identical keys and values in both arms."""
    return {"workload": workload_name(circ, batch), "batch_per_gpu": batch, "key": "shared 16-byte (AES-128)",
            "l2": f"tables are {batch * circ.num_rows * 16 // 1000000} MB per step, larger than L2; no flush needed"}


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


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json (measured)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(circuit: str, kernel: str):
    """DRAM bytes per launch of this circuit's kernel from the committed ncu capture, or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t[circuit][kernel]["dram_bytes"]
    except Exception:
        return None


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
        "impl": "reference", "metric": METRIC.replace("AES-128 circuit", WORKLOADS[args.circuit][0]), "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64/u8 (AES-NI)",
        "data": "synthetic",
        "config": config_of(circ, batch),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} instances per step x {args.steps} steps, C oracle (AES-NI), "
                                   f"{threads} pthreads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(index: int):
    """One process per GPU: run this rank's host threads, and so first-touch its pinned staging buffers, on
    the NUMA node the GPU hangs off (sysfs numa_node / local_cpulist of the PCI function).  Returns a short
    description for the JSON line; does nothing where sysfs has no answer (single-node VMs)."""
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


def host_topology():
    """What the box offers the host-pointer path: NUMA nodes and cores."""
    nodes = []
    try:
        for d in sorted(os.listdir("/sys/devices/system/node")):
            if d.startswith("node") and d[4:].isdigit():
                nodes.append(int(d[4:]))
    except Exception:
        pass
    return {"numa_nodes": len(nodes) or None, "cpus": os.cpu_count()}


# ------------------------------------------------------------------------------------------------
class Bench:
    """Per-rank state shared by the headline loop and the extras."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.numa = bind_to_gpu_numa_node(self.local) if self.world > 1 else None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        from mpc_b200 import _lib
        self._lib, self.L = _lib, _lib.lib()
        _lib.check(self.L.gcb_set_device(self.local))
        self.stream = torch.cuda.current_stream()
        self.s = self.stream.cuda_stream
        self.peak, self.peak_source = measured_peak()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor([float(v) for v in values], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)

    def to_dev(self, a):
        return self.torch.from_numpy(a.view(np.uint8).reshape(a.shape + (-1,)) if a.dtype.fields else a).to(self.dev)

    def launches(self) -> int:
        return int(self.L.gcb_launch_count())


def device_loop(b: Bench, circ, eng, batch: int, keys, steps: int, warmup: int, sampler=None):
    """Garble + select + eval + decode, device-resident, `steps` times.  keys: bytes (one shared key) or
    uint8[batch, keylen] (per-instance keys).  Returns timings (max over ranks) and the device tensors."""
    from mpc_b200.circuit import decode_bits_dev, select_labels_dev
    torch = b.torch
    nin, nout, rows = circ.num_inputs, circ.num_outputs, circ.num_rows
    rand, r, l0, bits = synthetic_inputs(circ, batch, b.rank)
    if isinstance(keys, bytes):
        d_key, klen, kstride = torch.frombuffer(bytearray(keys), dtype=torch.uint8).to(b.dev), len(keys), 0
    else:
        d_key, klen, kstride = b.to_dev(np.ascontiguousarray(keys)), keys.shape[1], keys.shape[1]
    d_r, d_l0, d_bits = b.to_dev(r), b.to_dev(l0), b.to_dev(bits)
    d_tab = torch.empty((batch, rows, 16), dtype=torch.uint8, device=b.dev)
    d_io = torch.empty((batch, nin + nout, 32), dtype=torch.uint8, device=b.dev)
    d_in = torch.empty((batch, nin, 16), dtype=torch.uint8, device=b.dev)
    d_out = torch.empty((batch, nout, 16), dtype=torch.uint8, device=b.dev)
    d_obits = torch.empty((batch, nout), dtype=torch.uint8, device=b.dev)
    kern = {"garble": [], "eval": []}

    def step(timed: bool):
        e = [b.ev() for _ in range(4)] if timed else None
        if timed: e[0].record(b.stream)
        eng.garble_dev(d_key, klen, kstride, batch, d_r, d_l0, d_tab, d_io, stream=b.s)
        if timed: e[1].record(b.stream)
        select_labels_dev(d_io, nin + nout, d_bits, d_in, batch, nin, stream=b.s)
        if timed: e[2].record(b.stream)
        eng.eval_dev(d_key, klen, kstride, batch, d_tab, d_in, d_out, stream=b.s)
        if timed: e[3].record(b.stream)
        if timed:
            kern["garble"].append((e[0], e[1])); kern["eval"].append((e[2], e[3]))
        # output wires are a strided view of io_wires: decode takes the wire stride
        decode_bits_dev(d_io.data_ptr() + nin * 32, nin + nout, d_out, d_obits, batch, nout, stream=b.s)

    for _ in range(max(warmup, 3)):
        step(False)
    b.barrier()
    # correctness of what is being timed: decoded outputs equal the plaintext function of the inputs
    ob = d_obits.cpu().numpy()
    if circ.name == "aes_128":
        from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
        enc = Cipher(algorithms.AES(bytes(range(16))), modes.ECB()).encryptor()
        for i in (0, 1, batch // 2, batch - 1):
            want = int.from_bytes(enc.update((b.rank * batch + i).to_bytes(16, "big")), "big")
            got = sum(int(x) << k for k, x in enumerate(ob[i]))
            assert got == want, f"instance {i}: decoded output is not AES(key, index)"
    else:
        for i in (0, batch - 1):
            assert np.array_equal(ob[i], circ.compute_bits(bits[i].tolist())), f"instance {i}: decoded output is wrong"
    if sampler:
        sampler.start()
    b.barrier()
    n0 = b.launches()
    t0, t1 = b.ev(), b.ev()
    t0.record(b.stream)
    for _ in range(steps):
        step(True)
    t1.record(b.stream)
    b.barrier()
    n1 = b.launches()
    ms = t0.elapsed_time(t1)
    g_ms = float(np.mean([x.elapsed_time(y) for x, y in kern["garble"]]))
    e_ms = float(np.mean([x.elapsed_time(y) for x, y in kern["eval"]]))
    ms, g_ms, e_ms = b.max_over_ranks([ms, g_ms, e_ms])
    return {"ms": ms, "garble_ms": g_ms, "eval_ms": e_ms, "launches": n1 - n0,
            "inputs": (r, l0, bits), "dev": (d_tab, d_out, d_io)}


def roofline_of(b: Bench, circ, eng, batch: int, g_ms: float, e_ms: float, nr: int):
    gb, eb = algorithmic_bytes(circ)
    ach_g, ach_e = gb * batch / (g_ms * 1e-3) / 1e9, eb * batch / (e_ms * 1e-3) / 1e9
    info = eng.info_for(batch)
    geo = f"{info.teams_per_sm} teams x {info.team_threads} threads per SM" + (
        f", split live ranges: {info.num_hot_slots} of {info.num_slots} labels in shared memory" if info.num_hot_slots < info.num_slots else "")
    return {"bound": "hbm", "achieved": ach_g, "peak": b.peak, "unit": "GB/s", "frac": ach_g / b.peak,
            "traffic": ncu_traffic(circ.name, "garble_kernel"), "algorithmic_bytes": gb * batch,
            "kernel": f"garble_kernel<NR={nr},PLAIN> ({geo})", "kernel_ms": g_ms,
            "eval": {"achieved": ach_e, "frac": ach_e / b.peak, "algorithmic_bytes": eb * batch, "kernel_ms": e_ms,
                     "traffic": ncu_traffic(circ.name, "eval_kernel"), "kernel": f"eval_kernel<NR={nr},PLAIN> ({geo})"},
            "peak_source": b.peak_source,
            "note": "bound by the shared-memory pipe (AES T-table lookups), not HBM: there is no AES instruction on the "
                    "GPU; see DESIGN.md section 4 and profiles/"}


def e2e_loop(b: Bench, circ, eng, batch: int, inputs, ref_dev, steps: int, pinned: bool):
    """The call a user makes, from ONE host thread, for a stream of batches: gcb_garble_begin on every part of the
    step's batch, then per part gcb_job_wait -> gcb_eval_begin; the eval jobs of a step are only waited for when
    their buffers come round again (three sets of host buffers) and the first parts of the next step's garbling are
    queued behind this step's last, so the garbler's tables of step k+1 stream back (D2H) while the evaluator's
    tables of step k stream in (H2D).  Every step copies its own inputs up and its own results down inside the
    timed region; the last step is fully drained before the clock stops."""
    from mpc_b200.circuit import host_alloc, host_free
    from mpc_b200.circuit_io import LABEL_DTYPE, WIRE_DTYPE
    nin, nout, rows = circ.num_inputs, circ.num_outputs, circ.num_rows
    r, l0, bits = inputs
    alloc = host_alloc if pinned else (lambda shape, dt: np.zeros(shape, dtype=dt))
    n_sets = 3
    sets = []
    for _ in range(n_sets):
        h = {"r": alloc((batch,), LABEL_DTYPE), "l0": alloc((batch, nin), LABEL_DTYPE), "tab": alloc((batch, rows), LABEL_DTYPE),
             "io": alloc((batch, nin + nout), WIRE_DTYPE), "in": alloc((batch, nin), LABEL_DTYPE), "out": alloc((batch, nout), LABEL_DTYPE)}
        h["r"][:] = r
        h["l0"][:] = l0
        sets.append(h)
    n_parts = max(1, min(int(os.environ.get("GCB_E2E_PARTS", E2E_PARTS)), batch // 64))
    parts = [slice(k * batch // n_parts, (k + 1) * batch // n_parts) for k in range(n_parts)]
    eval_jobs = [[] for _ in range(n_sets)]        # eval jobs still reading / writing each buffer set
    garble_jobs = [None] * n_sets

    prof = {"drain": 0.0, "issue_g": 0.0, "wait_g": 0.0, "issue_e": 0.0}
    ahead = min(int(os.environ.get("GCB_E2E_AHEAD", "3")), max(1, batch // 64))   # parts of the next step queued ahead

    def drain(k):
        t = time.perf_counter()
        for j in eval_jobs[k]:
            j.wait()
        eval_jobs[k] = []
        prof["drain"] += time.perf_counter() - t

    def garble_part(k, i):                          # queue part i of the garbler's step k
        h, sl = sets[k % n_sets], parts[i]
        if i == 0:
            drain(k % n_sets)                       # the set's previous step (k - n_sets) has left its buffers
            garble_jobs[k % n_sets] = []
        t = time.perf_counter()
        garble_jobs[k % n_sets].append(eng.garble_begin(KEY, h["r"][sl], h["l0"][sl], h["tab"][sl], h["io"][sl]))
        prof["issue_g"] += time.perf_counter() - t

    def run(n):
        # A step queues its garbling, then hands each part to the evaluator as soon as its tables are on the host.
        # The first `ahead` parts of the NEXT step's garbling are queued behind the last parts of this step, so that
        # their kernels have run when this step's last tables have crossed PCIe: the device->host engine never waits
        # for the host to queue a step or for its first kernel.  (Queuing the whole next step ahead measured slower:
        # 24.4 against 23.3 ms.)
        issued = 0                                  # parts of step k already queued ahead
        for k in range(n):
            for i in range(issued, n_parts):
                garble_part(k, i)
            issued = 0
            h = sets[k % n_sets]
            for i, sl in enumerate(parts):
                t = time.perf_counter()
                garble_jobs[k % n_sets][i].wait()   # this part's tables are on the host: its evaluation may start
                t1 = time.perf_counter()
                eval_jobs[k % n_sets].append(eng.eval_begin(KEY, h["tab"][sl], h["in"][sl], h["out"][sl]))
                prof["wait_g"] += t1 - t
                prof["issue_e"] += time.perf_counter() - t1
                if k + 1 < n and i >= n_parts - ahead:
                    garble_part(k + 1, issued)
                    issued += 1
        for k in range(n_sets):
            drain(k)

    for k in range(n_sets):                         # warm-up: also fills the evaluator's input labels of every set
        ahead, keep = 0, ahead
        run(1)
        ahead = keep
        sets.append(sets.pop(0))                    # run(1) uses set 0: rotate so that every set gets its turn
        h = sets[-1]
        h["in"][:] = np.where(bits.astype(bool), h["io"]["l1"][:, :nin], h["io"]["l0"][:, :nin])
    run(n_sets)
    b.barrier()
    for k in prof:
        prof[k] = 0.0
    t0 = time.perf_counter()
    run(steps)
    b.torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / steps
    if os.environ.get("GCB_E2E_TRACE") and b.rank == 0:
        sys.stderr.write(f"e2e trace (pinned={pinned}, ahead={ahead}, {steps} steps, {sec * 1e3:.2f} ms/step): "
                         + ", ".join(f"{k} {v * 1e3 / steps:.2f} ms/step" for k, v in prof.items()) + "\n")
    d_tab, d_out, _ = ref_dev
    for h in sets:
        ok = h["out"].tobytes() == d_out.cpu().numpy().tobytes() and h["tab"].tobytes() == d_tab.cpu().numpy().tobytes()
        assert ok, "host-pointer path and device-resident path disagree"
    if pinned:
        for h in sets:
            for a in h.values():
                host_free(a)
    return sec, n_parts


def pcie_probe(b: Bench, h2d_bytes: int, d2h_bytes: int, reps: int = 5, parts: int = E2E_PARTS, sets: int = 3):
    """Copy-only floors of the e2e step on this box, every rank at once, page-locked memory, two streams:
    `both_ms`     the step's bytes up and down as two single copies between fixed buffers (the best the links can do);
    `pattern_ms`  the same bytes in the e2e path's pattern: `parts` copies per direction per step, the device->host copy
                  of part i finished before the host->device copy of part i is queued FROM THE SAME HOST MEMORY (the
                  evaluator reads what the garbler wrote), rotating over `sets` sets of host buffers."""
    torch = b.torch
    h_in = [torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory() for _ in range(sets)]
    h_out = [torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory() for _ in range(sets)]
    d_in = torch.empty(h2d_bytes, dtype=torch.uint8, device=b.dev)
    d_out = torch.empty(d2h_bytes, dtype=torch.uint8, device=b.dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(up: bool, down: bool):
        b.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in[0], non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    h_out[0].copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def pattern():
        cu = [(k * h2d_bytes // parts, (k + 1) * h2d_bytes // parts) for k in range(parts)]
        cd = [(k * d2h_bytes // parts, (k + 1) * d2h_bytes // parts) for k in range(parts)]
        b.barrier()
        t0 = time.perf_counter()
        evs = []
        with torch.cuda.stream(s2):                       # step 0's results start down
            for lo, hi in cd:
                h_out[0][lo:hi].copy_(d_out[lo:hi], non_blocking=True)
                e = torch.cuda.Event(); e.record(s2); evs.append(e)
        for k in range(reps):
            nxt = []
            for i in range(parts):
                evs[i].synchronize()                      # part i is on the host ...
                with torch.cuda.stream(s1):               # ... and goes up again
                    lo, hi = cu[i]
                    d_in[lo:hi].copy_(h_out[k % sets][lo:hi], non_blocking=True)   # the bytes that just came down
                if k + 1 < reps:
                    with torch.cuda.stream(s2):           # the next step's part i comes down behind this step's
                        lo, hi = cd[i]
                        h_out[(k + 1) % sets][lo:hi].copy_(d_out[lo:hi], non_blocking=True)
                        e = torch.cuda.Event(); e.record(s2); nxt.append(e)
            evs = nxt
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    run(True, True)
    both, up, down = run(True, True), run(True, False), run(False, True)
    pattern()
    pat = pattern()
    both, up, down, pat = b.max_over_ranks([both, up, down, pat])
    return {"both_ms": both * 1e3, "pattern_ms": pat * 1e3,
            "h2d_alone_gbs_per_gpu": h2d_bytes / up / 1e9, "d2h_alone_gbs_per_gpu": d2h_bytes / down / 1e9,
            "both_gbs_per_gpu_each_way": (h2d_bytes + d2h_bytes) / 2 / both / 1e9,
            "box_total_gbs": (h2d_bytes + d2h_bytes) * b.world / both / 1e9,
            "how": f"{b.world} rank(s) at once, {h2d_bytes / 1e9:.2f} GB up + {d2h_bytes / 1e9:.2f} GB down per rank, pinned memory, "
                   f"two streams, mean of {reps}; pattern: {parts} copies per direction per step, up after down per part, {sets} buffer sets"}


def extra_iknp(b: Bench, n: int = 1 << 24):
    """BASELINE config 4: IKNP expansion of 2^24 OTs, receiver then sender on the receiver's U, labels in HBM."""
    torch, L, check, ptr = b.torch, b.L, b._lib.check, b._lib.ptr
    rnd = lambda *shape: torch.randint(0, 256, shape, dtype=torch.uint8, device=b.dev)
    k0, k1, delta = rnd(128, 16), rnd(128, 16), rnd(1, 16)
    choice = torch.randint(0, 2, (n,), dtype=torch.uint8, device=b.dev)
    ul = L.gcb_iknp_u_size(n)
    u = torch.empty(ul, dtype=torch.uint8, device=b.dev)
    t_lab = torch.empty((n, 16), dtype=torch.uint8, device=b.dev)
    q_lab = torch.empty((n, 16), dtype=torch.uint8, device=b.dev)
    out = {}
    for pos in (0, 17):
        recv = lambda: check(L.gcb_iknp_receiver_expand_dev(ptr(k0), ptr(k1), pos, ptr(choice), n, ptr(u), ptr(t_lab), b.s))
        send = lambda: check(L.gcb_iknp_sender_expand_dev(ptr(k0), ptr(delta), pos, ptr(u), ul, n, ptr(q_lab), b.s))
        for name, fn, blocks in (("receiver", recv, 2), ("sender", send, 1)):
            for _ in range(3):
                fn()
            b.barrier()
            e0, e1 = b.ev(), b.ev()
            e0.record(b.stream)
            for _ in range(5):
                fn()
            e1.record(b.stream)
            b.barrier()
            ms = b.max_over_ranks([e0.elapsed_time(e1) / 5])[0]
            gb = 32.0 * n / 1e9
            out[f"{name}_pos{pos}"] = {"ms": ms, "m_ot_per_s": n * b.world / ms / 1e3, "aes_blocks_per_ot": blocks,
                                       "algorithmic_bytes": int(32 * n), "achieved_gbs": gb / ms * 1e3, "hbm_frac": gb / ms * 1e3 / b.peak}
    out["workload"] = f"IKNP expansion of 2^24 OTs per GPU (transpose + AES-CTR column PRG), stream positions 0 and 17"
    return out


def extra_stream_step(b: Bench, batch: int = 256):
    """BASELINE config 3: sha256.circ as one Streaming.Garble step (gate kernel + record serialisation + D2H of the
    byte stream), then the streaming evaluator on those bytes."""
    from mpc_b200.circuit import GarbleEngine, StreamEval, Streaming
    from mpc_b200.circuit_io import LABEL_DTYPE
    circ = load_circuit("sha256")
    eng = GarbleEngine(circ)
    rng = np.random.default_rng(77 + b.rank)
    nin = circ.num_inputs
    r = rng.integers(0, 2**63, (batch, 2), dtype=np.uint64).view(LABEL_DTYPE).reshape(batch)
    l0 = rng.integers(0, 2**63, (batch, nin, 2), dtype=np.uint64).view(LABEL_DTYPE).reshape(batch, nin)
    ids = list(range(nin))
    st = Streaming(bytes(32), r, ids, l0)
    sev = StreamEval(bytes(32), batch)
    w = st.get_inputs(ids)
    sev.set(ids, w["l0"].astype(LABEL_DTYPE))
    outs = list(range(nin, nin + circ.num_outputs))
    nbytes = 0
    tg = te = 0.0
    for k in range(5):
        b.barrier()
        t0 = time.perf_counter()
        buf, _, _ = st.garble(eng, ids, outs)
        t1 = time.perf_counter()
        sev.circuit(buf, circ.num_gates, circ.num_wires, nin + circ.num_outputs)
        sev.get(outs[:1])                              # ordered behind the eval kernel
        t2 = time.perf_counter()
        if k >= 2:
            tg += (t1 - t0) / 3
            te += (t2 - t1) / 3
        nbytes = buf.shape[1]
    tg, te = b.max_over_ranks([tg, te])
    gates = circ.num_gates * batch * b.world
    return {"workload": f"sha256.circ as one streaming step, batch {batch} per GPU, 32-byte key; record stream to / from host memory",
            "stream_bytes_per_instance": int(nbytes), "garble_ms": tg * 1e3, "eval_ms": te * 1e3,
            "m_gates_per_s_garble": gates / tg / 1e6, "m_gates_per_s_eval": gates / te / 1e6,
            "pcie_gbs_garble": nbytes * batch / tg / 1e9, "pcie_gbs_eval": nbytes * batch / te / 1e9}


def extra_stream_program(b: Bench, batch: int):
    """BASELINE config 5 stand-in at >= 10^8 gates per instance (tools/stream_program.py)."""
    import stream_program as sp
    steps = sp.steps_for_gates(1e8)
    res = sp.run_program(steps, batch, check=1, warm=4, rank=b.rank, oracle_steps=1 if b.rank == 0 else 0,
                         sync=(b.dist.barrier if b.world > 1 else None))
    wall, tg, te, bad = b.max_over_ranks([res["wall_s"], res["garble_s"], res["eval_s"], 0.0 if res["checks_ok"] else 1.0])
    g = res["timed_gates"] * b.world
    return {"workload": f"{steps} steps of sha512.circ / mul64.circ chained through permanent wires (stand-in for ed25519 "
                        f"sign.mpcl, which needs the Go MPCL compiler), batch {batch} per GPU, garbler + evaluator",
            "gates_per_instance": res["gates_per_instance"], "timed_gates_per_instance": res["timed_gates_per_instance"],
            "total_gates_timed": g, "stream_gb": res["stream_bytes"] * b.world / 1e9, "wall_s": wall,
            "m_gates_per_s": g / wall / 1e6, "m_and_per_s": g / wall / 1e6 * (3 * 57947 + 4033) / (3 * 349617 + 13675),
            "pcie_gbs_each_way": res["stream_bytes"] * b.world / wall / 1e9, "checks_ok": bad == 0.0}


def extra_latency(b: Bench):
    """The unchanged drop-in call: batch = 1 gcb_garble (with the full Wires array) + gcb_eval from host buffers,
    wall clock, next to the oracle on ONE thread (the reference's execution model, circuit/garble_bench_test.go:38-64)."""
    from mpc_b200.circuit import GarbleEngine
    from mpc_b200.circuit_io import LABEL_DTYPE
    from mpc_b200.drbg import DRBG
    from oracle import pyoracle as O
    out = {}
    for name in ("sha256xor", "aes_128"):
        circ = load_circuit(name)
        eng = GarbleEngine(circ)
        key = bytes(range(32))
        rand = DRBG(f"lat/{name}").read(16 * (1 + circ.num_inputs))

        def gpu_once():
            g = eng.garble(rand, key)
            wires = np.zeros(circ.num_wires, dtype=LABEL_DTYPE)
            wires[: circ.num_inputs] = g.Wires["l0"][: circ.num_inputs]
            eng.eval(key, wires, g)

        def cpu_once():
            _, o_wires, o_slab, o_off = O.garble(circ, key, rand)
            O.eval_(circ, key, np.ascontiguousarray(o_wires["l0"][: circ.num_inputs]), o_slab, o_off)

        res = {}
        for tag, fn in (("gpu_ms", gpu_once), ("cpu_1thread_ms", cpu_once)):
            fn(); fn()
            ts = []
            for _ in range(5):
                t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
            res[tag] = 1e3 * float(np.median(ts))
        # the throughput form of the same call: how many instances one call needs before the GPU wins
        from util import garble_inputs, rand_to_labels
        for nb in (8, 64):
            keys, rr = garble_inputs(f"lat/{name}/{nb}", nb, circ.num_inputs, 32)
            r, l0 = rand_to_labels(rr, circ.num_inputs)
            eng.garble_batch(keys[0].tobytes(), r, l0)
            t0 = time.perf_counter()
            for _ in range(3):
                tables, io = eng.garble_batch(keys[0].tobytes(), r, l0)
            res[f"gpu_garble_batch{nb}_ms_per_instance"] = 1e3 * (time.perf_counter() - t0) / 3 / nb
        out[name] = res
    out["note"] = ("batch = 1, host buffers, wall clock incl. copies; gpu = gcb_garble(wires_full) + gcb_eval(wires_full) as the "
                   "Go binding's Garble / Eval call them; cpu = the C oracle on one thread")
    return out


def run_gcb(args):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: `python bench.py --gpus N` re-launches itself under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    b = Bench(args)
    rank, world = b.rank, b.world
    from mpc_b200.circuit import GarbleEngine

    circ = load_circuit(args.circuit)
    eng = GarbleEngine(circ)
    nin, nout, rows, n_and = circ.num_inputs, circ.num_outputs, circ.num_rows, circ.count(2)
    batch = args.batch or WORKLOADS[args.circuit][1]

    sampler = ClockSampler(b.local) if rank == 0 else None
    head = device_loop(b, circ, eng, batch, KEY, args.steps, args.warmup, sampler)
    clocks = sampler.stop() if sampler else None
    ms, g_ms, e_ms = head["ms"], head["garble_ms"], head["eval_ms"]

    h2d = batch * (16 * (1 + nin) + 16 * rows + 16 * nin)
    d2h = batch * (16 * rows + 32 * (nin + nout) + 16 * nout)
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(2, min(args.steps, 10))
        sec, n_parts = e2e_loop(b, circ, eng, batch, head["inputs"], head["dev"], e2e_steps, pinned=True)
        e2e_ms = b.max_over_ranks([sec * 1e3])[0]
        e2e = {"value": n_and * batch * world / (e2e_ms * 1e-3) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(world * h2d), "d2h_bytes_per_step": int(world * d2h), "ms_per_step": e2e_ms,
               "how": f"gcb_garble_begin / gcb_eval_begin / gcb_job_wait on page-locked host buffers (gcb_host_alloc), "
                      f"{n_parts} parts per step, three buffer sets, the first parts of the next step's garbling queued behind this step's last (the tables of step k+1 stream back while those of step k "
                      f"stream in), ONE host thread per GPU"
                      + (f", ranks bound to their GPU's NUMA node ({b.numa['cpus']} cpus)" if b.numa else "")}

    extra = {}
    if not args.no_extra:
        t_extra = time.perf_counter()
        probe = pcie_probe(b, h2d, d2h)
        extra["pcie_probe"] = probe
        if e2e:
            e2e["pcie_floor_ms"] = probe["both_ms"]
            e2e["frac_of_pcie_floor"] = probe["both_ms"] / e2e["ms_per_step"]
            e2e["pcie_pattern_ms"] = probe["pattern_ms"]
            e2e["frac_of_pcie_pattern"] = probe["pattern_ms"] / e2e["ms_per_step"]
            sec, _ = e2e_loop(b, circ, eng, batch, head["inputs"], head["dev"], 2, pinned=False)
            p_ms = b.max_over_ranks([sec * 1e3])[0]
            extra["e2e_pageable"] = {"value": n_and * batch * world / (p_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": p_ms,
                                     "how": "the same calls on pageable (malloc) host buffers: one extra host memcpy each way, "
                                            "done by the calling thread"}
        del head["dev"]
        torch.cuda.empty_cache()
        # key shape (ii) of SURVEY 8d: per-instance random 32-byte keys (circuit/garbler.go:47-53, sha2pc/garbler.go:96-101)
        keys32 = np.random.default_rng(99 + rank).integers(0, 256, (batch, 32), dtype=np.uint8)
        k32 = device_loop(b, circ, eng, batch, keys32, max(3, args.steps // 3), 3)
        steps32 = max(3, args.steps // 3)
        extra["aes_128_keys32"] = {
            "workload": workload_name(circ, batch) + ", per-instance random 32-byte keys (AES-256, 14 rounds, key schedule per instance)",
            "value": n_and * batch * world * steps32 / (k32["ms"] * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": k32["ms"] / steps32,
            "roofline": roofline_of(b, circ, eng, batch, k32["garble_ms"], k32["eval_ms"], 14)}
        del k32
        torch.cuda.empty_cache()
        sha = load_circuit("sha256")
        sha_eng = GarbleEngine(sha)
        sb = WORKLOADS["sha256"][1]
        steps_s = max(3, args.steps // 3)
        sh = device_loop(b, sha, sha_eng, sb, KEY, steps_s, 3)
        extra["sha256"] = {
            "workload": workload_name(sha, sb) + ", shared 16-byte key",
            "value": sha.count(2) * sb * world * steps_s / (sh["ms"] * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": sh["ms"] / steps_s,
            "roofline": roofline_of(b, sha, sha_eng, sb, sh["garble_ms"], sh["eval_ms"], 10)}
        del sh
        torch.cuda.empty_cache()
        extra["iknp_2p24"] = extra_iknp(b)
        torch.cuda.empty_cache()
        extra["stream_sha256_step"] = extra_stream_step(b)
        torch.cuda.empty_cache()
        extra["stream_program"] = extra_stream_program(b, args.program_batch)
        torch.cuda.empty_cache()
        if world == 1:
            extra["latency_batch1"] = extra_latency(b)
        extra["host"] = host_topology()
        extra["seconds"] = time.perf_counter() - t_extra

    if rank == 0:
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
            "metric": METRIC.replace("AES-128 circuit", WORKLOADS[args.circuit][0]), "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32 (AES T-tables, 128-bit label XOR)",
            "data": "synthetic",
            "config": config_of(circ, batch),
            "roofline": roofline_of(b, circ, eng, batch, g_ms, e_ms, 10),
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": head["launches"],
            "clocks": clocks,
            "extra": extra or None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        b.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gcb", choices=["gcb", "reference"])
    ap.add_argument("--circuit", default="aes_128", choices=sorted(WORKLOADS), help="aes_128 = the headline workload")
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU (default: 4096 for aes_128, 2368 for sha256)")
    ap.add_argument("--program-batch", type=int, default=148, help="instances per GPU of the streaming program in `extra`")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: device-resident loop only")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary configurations")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gcb(args)


if __name__ == "__main__":
    main()
