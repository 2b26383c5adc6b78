"""Where the host-pointer path loses time: garble-only / eval-only / both, by parts and worker threads.
python tools/e2e_probe2.py"""
import ctypes as C, os, sys, time
from concurrent.futures import ThreadPoolExecutor
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
def t(f, reps=4):
    f(); t0 = time.perf_counter()
    for _ in range(reps): f()
    return (time.perf_counter() - t0) / reps * 1e3
def run(parts, gw, ew, mode):
    sls = [slice(k * batch // parts, (k + 1) * batch // parts) for k in range(parts)]
    gpool, epool = ThreadPoolExecutor(max(gw, 1)), ThreadPoolExecutor(max(ew, 1))
    def both():
        efs = []
        gfs = [gpool.submit(g, sl) for sl in sls] if mode != "e" else []
        if mode == "g":
            for f in gfs: f.result()
            return
        if mode == "e":
            efs = [epool.submit(e, sl) for sl in sls]
        else:
            for f, sl in zip(gfs, sls):
                f.result(); efs.append(epool.submit(e, sl))
        for f in efs: f.result()
    ms = t(both)
    gpool.shutdown(); epool.shutdown()
    print(f"stagger={os.environ.get('GCB_STAGGER','dflt')} mode={mode} parts={parts} workers={gw}+{ew}: {ms:.2f} ms", flush=True)
for stg in (None, "0"):
    if stg is None: os.environ.pop("GCB_STAGGER", None)
    else: os.environ["GCB_STAGGER"] = stg
    for cfg in sys.argv[1:] or ["16,2,2,g", "16,2,2,e", "16,2,2,b", "16,2,3,b", "16,3,3,b", "32,2,2,b", "8,1,1,g", "8,1,1,e"]:
        p_, gw, ew, mode = cfg.split(",")
        run(int(p_), int(gw), int(ew), mode)
