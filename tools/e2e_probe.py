"""Host-pointer path timing: gcb_garble alone, gcb_eval alone, both pipelined (pinned buffers)."""
import ctypes as C, os, sys, time, threading, queue
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_circuit
from mpc_b200 import _lib
from mpc_b200.circuit import GarbleEngine
from mpc_b200.circuit_io import LABEL_DTYPE, WIRE_DTYPE
circ = load_circuit("aes_128"); eng = GarbleEngine(circ); L = _lib.lib()
batch, nin, nout, rows = 4096, circ.num_inputs, circ.num_outputs, circ.num_rows
def pinned(shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = L.gcb_host_alloc(n)
    return np.frombuffer((C.c_uint8 * n).from_address(p), dtype=dtype).reshape(shape)
rng = np.random.default_rng(0)
h_r = pinned((batch,), LABEL_DTYPE); h_l0 = pinned((batch, nin), LABEL_DTYPE)
h_r.view(np.uint8).reshape(-1)[:] = rng.integers(0, 256, h_r.nbytes, dtype=np.uint8); h_l0.view(np.uint8).reshape(-1)[:] = rng.integers(0, 256, h_l0.nbytes, dtype=np.uint8)
h_tab = pinned((batch, rows), LABEL_DTYPE); h_io = pinned((batch, nin + nout), WIRE_DTYPE)
h_in = pinned((batch, nin), LABEL_DTYPE); h_out = pinned((batch, nout), LABEL_DTYPE)
KEY = b"0123456789abcdef"
def g(sl=slice(None)): eng.garble_batch(KEY, h_r[sl], h_l0[sl], tables=h_tab[sl], io_wires=h_io[sl])
def e(sl=slice(None)): eng.eval_batch(KEY, h_tab[sl], h_in[sl], out_labels=h_out[sl])
g(); h_in[:] = h_io["l0"][:, :nin]; e()
def t(f, reps=5):
    f(); t0 = time.perf_counter()
    for _ in range(reps): f()
    return (time.perf_counter() - t0) / reps * 1e3
print("garble alone %.2f ms   eval alone %.2f ms" % (t(g), t(e)))
for parts in (1, 2, 4, 8):
    sls = [slice(k * batch // parts, (k + 1) * batch // parts) for k in range(parts)]
    def both():
        q = queue.Queue()
        def gg():
            for sl in sls: g(sl); q.put(sl)
        th = threading.Thread(target=gg); th.start()
        for _ in sls: e(q.get())
        th.join()
    print("pipelined, %d parts: %.2f ms" % (parts, t(both)))

from concurrent.futures import ThreadPoolExecutor
for parts, workers in ((8, 2), (8, 3), (16, 2), (16, 4)):
    sls = [slice(k * batch // parts, (k + 1) * batch // parts) for k in range(parts)]
    gpool, epool = ThreadPoolExecutor(workers), ThreadPoolExecutor(workers)
    def both():
        efs = []
        gfs = [gpool.submit(g, sl) for sl in sls]
        for f, sl in zip(gfs, sls):
            f.result()
            efs.append(epool.submit(e, sl))
        for f in efs: f.result()
    print("pipelined, %d parts, %d+%d worker threads: %.2f ms" % (parts, workers, workers, t(both)))
