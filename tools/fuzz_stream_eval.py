"""Fuzz of the streaming evaluator with mutated record streams (the stream comes from the peer): every call must end in
an error or a result, and -- run under `compute-sanitizer --tool memcheck` -- without a single invalid device access.
  compute-sanitizer --tool memcheck python tools/fuzz_stream_eval.py [iterations]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_circuit  # noqa: E402
from mpc_b200 import _lib  # noqa: E402
from mpc_b200.circuit import GarbleEngine, StreamEval, Streaming  # noqa: E402
from mpc_b200.circuit_io import LABEL_DTYPE  # noqa: E402
from mpc_b200.drbg import DRBG  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 120
circ = load_circuit("sub64")
batch, ids, outs = 2, list(range(128)), list(range(200, 264))
keys = np.stack([DRBG(f"fz/key/{b}").array(32) for b in range(batch)])
lab = np.stack([np.frombuffer(DRBG(f"fz/{b}").read(16 * 129), dtype=">u8").astype("<u8").view(LABEL_DTYPE) for b in range(batch)])
st = Streaming(keys, np.ascontiguousarray(lab[:, 0]), ids, np.ascontiguousarray(lab[:, 1:]))
eng = GarbleEngine(circ)
buf, _, _ = st.garble(eng, ids, outs)
wires = st.get_inputs(ids)
labels = np.ascontiguousarray(wires["l0"]).astype(LABEL_DTYPE)
rng = np.random.default_rng(7)
ok = refused = 0
for it in range(iters):
    bad = buf.copy()
    kind = int(rng.integers(0, 5))
    n = bad.shape[1]
    if kind == 0:                                              # a few flipped bytes, the same in every instance
        for _ in range(int(rng.integers(1, 4))):
            bad[:, int(rng.integers(0, n))] = int(rng.integers(0, 256))
    elif kind == 1:                                            # ... in one instance only (headers that disagree)
        bad[int(rng.integers(0, batch)), int(rng.integers(0, n))] ^= int(rng.integers(1, 256))
    elif kind == 2:                                            # truncated
        bad = np.ascontiguousarray(bad[:, : int(rng.integers(0, n))])
    elif kind == 3:                                            # a wire index turned huge (first bytes of a record header)
        p = int(rng.integers(0, min(n, 4000)))
        bad[:, p: p + 4] = 0xff
    ngates = circ.num_gates if kind != 4 else int(rng.choice([0, 1, circ.num_gates + 1, 1 << 20, 0xffffffff]))
    nwires = int(rng.choice([264, 264, 264, 1, 200, 1 << 16, 0xffffffff]))
    sev = StreamEval(keys, batch)
    sev.set(ids, labels)
    import time
    t0 = time.perf_counter()
    print(f"it {it} kind {kind} ngates {ngates} nwires {nwires} len {bad.shape[1]}", end=" ", flush=True)
    try:
        if bad.shape[1]:
            sev.circuit(bad, ngates, circ.num_wires, nwires)
        sev.get(outs[:4])
        ok += 1
        print(f"ok {time.perf_counter() - t0:.2f}s", flush=True)
    except _lib.GcbError as e:
        refused += 1
        print(f"refused ({str(e)[:50]}) {time.perf_counter() - t0:.2f}s", flush=True)
print(f"fuzz_stream_eval: {ok} evaluated, {refused} refused of {iters}", flush=True)
