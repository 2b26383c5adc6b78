"""Device-time measurements of the secondary hot-path configurations (BASELINE.json
configs 3-4 and the K5/K6 kernels).  One JSON line per measurement; CUDA events, 3 warm-ups,
best of 5.  Run on one B200:  python tools/bench_paths.py > gpurun_out/paths.jsonl"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_circuit  # noqa: E402
from mpc_b200 import _lib  # noqa: E402
from mpc_b200._lib import Label, check, ptr  # noqa: E402
from mpc_b200.circuit import GarbleEngine, Streaming, select_labels_dev  # noqa: E402

PEAK = 6454.0
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
dev = torch.device("cuda:0")
L = _lib.lib()
s = torch.cuda.current_stream().cuda_stream


def timeit(fn, warm=3, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def emit(**kw):
    print(json.dumps(kw), flush=True)


def rnd(*shape):
    return torch.randint(0, 256, shape, dtype=torch.uint8, device=dev)


def iknp(n, pos):
    k0, k1, delta = rnd(128, 16), rnd(128, 16), rnd(1, 16)
    choice = torch.randint(0, 2, (n,), dtype=torch.uint8, device=dev)
    ul = L.gcb_iknp_u_size(n)
    u = torch.empty(ul, dtype=torch.uint8, device=dev)
    lab = torch.empty((n, 16), dtype=torch.uint8, device=dev)
    ms_r = timeit(lambda: check(L.gcb_iknp_receiver_expand_dev(ptr(k0), ptr(k1), pos, ptr(choice), n, ptr(u), ptr(lab), s)))
    ms_s = timeit(lambda: check(L.gcb_iknp_sender_expand_dev(ptr(k0), ptr(delta), pos, ptr(u), ul, n, ptr(lab), s)))
    for side, ms, blocks in (("receiver", ms_r, 2), ("sender", ms_s, 1)):
        gb = 32.0 * n / 1e9
        emit(path=f"iknp_{side}_expand", n=n, stream_pos=pos, ms=ms, ot_per_s=n / ms * 1e3, aes_blocks_per_s=blocks * n / ms * 1e3,
             algorithmic_gb=gb, achieved_gbs=gb / ms * 1e3, hbm_frac=gb / ms * 1e3 / PEAK)


def mitccrh(n, h):
    blks = rnd(n * h, 16)
    seed = Label(0x0123456789abcdef, 0xfedcba9876543210)
    ms = timeit(lambda: check(L.gcb_mitccrh_hash_dev(C.byref(seed), 0, ptr(blks), n, h, s)))
    gb = 32.0 * n * h / 1e9
    emit(path="mitccrh_hash", nkeys=n, h=h, ms=ms, keys_per_s=n / ms * 1e3, achieved_gbs=gb / ms * 1e3, hbm_frac=gb / ms * 1e3 / PEAK)


def hash_half(n, klen):
    key, x, out = rnd(klen), rnd(n, 16), torch.empty((n, 16), dtype=torch.uint8, device=dev)
    ms = timeit(lambda: check(L.gcb_hash_half_dev(ptr(key), klen, ptr(x), 0, ptr(out), n, s)))
    gb = 32.0 * n / 1e9
    emit(path="hash_half", n=n, keylen=klen, ms=ms, hashes_per_s=n / ms * 1e3, achieved_gbs=gb / ms * 1e3, hbm_frac=gb / ms * 1e3 / PEAK)


def cot(n):
    """COT / ROT post-processing and the malicious-mode sums on labels that stay in HBM."""
    seed, delta = Label(0x0123456789abcdef, 0xfedcba9876543210), Label(0x1111111111111111 | 1 << 63, 0x2222222222222222)
    q, wires, msgs = rnd(n, 16), rnd(n, 32), torch.empty((2 * n, 16), dtype=torch.uint8, device=dev)
    choice = torch.randint(0, 2, (n,), dtype=torch.uint8, device=dev)
    res = torch.empty((n, 16), dtype=torch.uint8, device=dev)
    out = np.zeros(6, dtype=np.uint64)
    runs = (
        ("cot_send", 80, 2, lambda: check(L.gcb_cot_send_dev(C.byref(seed), C.byref(delta), ptr(q), ptr(wires), n, ptr(msgs), 0, s))),
        ("cot_receive", 33, 1, lambda: check(L.gcb_cot_receive_dev(C.byref(seed), ptr(choice), ptr(msgs), ptr(q), n, ptr(res), 0, s))),
        ("rot_send", 48, 2, lambda: check(L.gcb_rot_send_dev(C.byref(seed), C.byref(delta), ptr(q), n, ptr(msgs), s))),
        ("rot_receive", 32, 1, lambda: check(L.gcb_rot_receive_dev(C.byref(seed), ptr(q), n, ptr(res), s))),
        ("iknp_check_sums", 17, 1, lambda: check(L.gcb_iknp_check_sums_dev(C.byref(seed), 0, ptr(q), ptr(choice), n, ptr(out), s))),
    )
    for name, bytes_per_ot, blocks, fn in runs:
        ms = timeit(fn)
        gb = bytes_per_ot * n / 1e9
        emit(path=name, n=n, ms=ms, ot_per_s=n / ms * 1e3, aes_blocks_per_s=blocks * n / ms * 1e3, algorithmic_gb=gb,
             achieved_gbs=gb / ms * 1e3, hbm_frac=gb / ms * 1e3 / PEAK)


def garble_eval(name, batch, klen, per_instance):
    circ = load_circuit(name)
    eng = GarbleEngine(circ)
    nin, nout, rows = circ.num_inputs, circ.num_outputs, circ.num_rows
    keys = rnd(batch, klen) if per_instance else rnd(klen)
    ks = klen if per_instance else 0
    r, l0, bits = rnd(batch, 16), rnd(batch, nin, 16), torch.randint(0, 2, (batch, nin), dtype=torch.uint8, device=dev)
    tab = torch.empty((batch, rows, 16), dtype=torch.uint8, device=dev)
    io = torch.empty((batch, nin + nout, 32), dtype=torch.uint8, device=dev)
    inl = torch.empty((batch, nin, 16), dtype=torch.uint8, device=dev)
    out = torch.empty((batch, nout, 16), dtype=torch.uint8, device=dev)
    ms_g = timeit(lambda: eng.garble_dev(keys, klen, ks, batch, r, l0, tab, io, stream=s))
    select_labels_dev(io, nin + nout, bits, inl, batch, nin, stream=s)
    ms_e = timeit(lambda: eng.eval_dev(keys, klen, ks, batch, tab, inl, out, stream=s))
    n_and = circ.count(2)
    gb_g = batch * (16 * (1 + nin) + 16 * rows + 32 * (nin + nout)) / 1e9
    i = eng.info
    emit(path="garble+eval", circuit=name, batch=batch, keylen=klen, per_instance_keys=per_instance,
         teams=i.teams_per_sm, team_threads=i.team_threads, slots=i.num_slots, ms_garble=ms_g, ms_eval=ms_e,
         m_and_per_s=n_and * batch / (ms_g + ms_e) / 1e3, garble_achieved_gbs=gb_g / ms_g * 1e3,
         garble_hbm_frac=gb_g / ms_g * 1e3 / PEAK,
         aes_blocks_per_s=(i.garble_hashes + i.eval_hashes) * batch / (ms_g + ms_e) * 1e3)


def streaming(name, batch, klen):
    circ = load_circuit(name)
    eng = GarbleEngine(circ)
    nin, nout = circ.num_inputs, circ.num_outputs
    rng = np.random.default_rng(0)
    from mpc_b200.circuit_io import LABEL_DTYPE
    r = rng.integers(0, 2**63, (batch, 2), dtype=np.uint64).view(LABEL_DTYPE).reshape(batch)
    l0 = rng.integers(0, 2**63, (batch, nin, 2), dtype=np.uint64).view(LABEL_DTYPE).reshape(batch, nin)
    st = Streaming(rng.integers(0, 256, klen, dtype=np.uint8).tobytes(), r, list(range(nin)), l0)
    ins, outs = list(range(nin)), list(range(nin, nin + nout))
    st.garble(eng, ins, outs)
    best = (1e18, 0, 0)
    for _ in range(3):
        buf, t0, t1 = st.garble(eng, ins, outs)
        best = min(best, (t0 + t1, t0, t1))
    emit(path="stream_garble_step", circuit=name, batch=batch, keylen=klen, stream_bytes_per_instance=int(buf.shape[1]),
         ms_init=best[1] / 1e6, ms_garble_incl_d2h=best[2] / 1e6,
         m_gates_per_s=circ.num_gates * batch / (best[0] / 1e9) / 1e6)


if __name__ == "__main__":
    what = sys.argv[1:] or ["iknp", "mitccrh", "cot", "hash", "gc", "stream"]
    if "iknp" in what:
        iknp(1 << 24, 0); iknp(1 << 24, 17)
    if "mitccrh" in what:
        mitccrh(1 << 24, 1); mitccrh(1 << 24, 2)
    if "cot" in what:
        cot(1 << 24)
    if "hash" in what:
        hash_half(1 << 24, 16); hash_half(1 << 24, 32)
    if "gc" in what:
        garble_eval("aes_128", 4096, 16, False)
        garble_eval("aes_128", 4096, 32, True)
        garble_eval("sha256", 1184, 16, False)
        garble_eval("sha256xor", 1184, 32, True)
        garble_eval("mul64", 4096, 16, False)
    if "stream" in what:
        streaming("sha256", 256, 32)
