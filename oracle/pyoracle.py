"""ctypes binding of the CPU oracle (oracle/gcb_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from mpc_b200/.
Arrays cross the boundary as numpy arrays with the dtypes of
mpc_b200.circuit_io (GATE_DTYPE, LABEL_DTYPE, WIRE_DTYPE).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from mpc_b200.circuit_io import GATE_DTYPE, LABEL_DTYPE, WIRE_DTYPE, Circuit

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgcb_oracle.so")

ERRORS = {
    -1: "crypto/aes: invalid key size",
    -2: "invalid gate type",
    -3: "corrupted ciruit: AND row length",
    -4: "corrupted circuit: index >= row",
    -5: "buffer too small",
    -6: "invalid chunk size",
    -7: "invalid argument",
}


class OracleError(RuntimeError):
    def __init__(self, rc):
        super().__init__(ERRORS.get(rc, f"oracle error {rc}"))
        self.rc = rc


class _Label(C.Structure):
    _fields_ = [("d0", C.c_uint64), ("d1", C.c_uint64)]


class _Wire(C.Structure):
    _fields_ = [("l0", _Label), ("l1", _Label)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gcb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "--no-print-directory"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, u8p, u32, u64, sz = C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_size_t
        L.orc_have_aesni.restype = C.c_int
        L.orc_set_aesni.argtypes = [C.c_int]
        L.orc_aes_encrypt_block.argtypes = [u8p, u32, u8p, u8p]
        L.orc_label_mul2.restype = _Label
        L.orc_label_mul2.argtypes = [_Label]
        L.orc_label_mul4.restype = _Label
        L.orc_label_mul4.argtypes = [_Label]
        L.orc_encrypt_half.argtypes = [u8p, u32, _Label, u32, C.POINTER(_Label)]
        L.orc_encrypt.argtypes = [u8p, u32, _Label, _Label, _Label, u32, C.POINTER(_Label)]
        L.orc_decrypt.argtypes = [u8p, u32, _Label, _Label, u32, _Label, C.POINTER(_Label)]
        L.orc_garble.argtypes = [vp, u32, u32, u32, u8p, u32, u8p, vp, vp, vp, vp]
        L.orc_eval.argtypes = [vp, u32, u32, u8p, u32, vp, vp, vp]
        L.orc_garble_batch.argtypes = [vp, u32, u32, u32, u32, u8p, u32, u32, u32, u8p, vp, vp, vp, C.c_int]
        L.orc_eval_batch.argtypes = [vp, u32, u32, u32, u32, u8p, u32, u32, u32, vp, vp, vp, C.c_int]
        L.orc_stream_new.restype = vp
        L.orc_stream_new.argtypes = [u8p, u32, u8p, vp, u32]
        L.orc_stream_free.argtypes = [vp]
        L.orc_stream_r.restype = _Label
        L.orc_stream_r.argtypes = [vp]
        L.orc_stream_get_input.restype = _Wire
        L.orc_stream_get_input.argtypes = [vp, u32]
        L.orc_stream_set_wire.argtypes = [vp, u32, _Wire]
        L.orc_stream_garble.argtypes = [vp, vp, u32, u32, vp, u32, vp, u32, u8p, sz, C.POINTER(sz)]
        L.orc_seval_new.restype = vp
        L.orc_seval_new.argtypes = [u8p, u32]
        L.orc_seval_free.argtypes = [vp]
        L.orc_seval_set.argtypes = [vp, u32, _Label]
        L.orc_seval_get.restype = _Label
        L.orc_seval_get.argtypes = [vp, u32]
        L.orc_seval_circuit.argtypes = [vp, u8p, sz, u32, u32, u32, C.POINTER(sz)]
        L.orc_prg.argtypes = [_Label, u64, u8p, sz]
        L.orc_create_labels.argtypes = [vp, sz, u8p, C.c_int]
        L.orc_iknp_receive.argtypes = [vp, vp, C.POINTER(u64), u8p, u64, u8p, sz, C.POINTER(sz), vp]
        L.orc_iknp_send.argtypes = [vp, _Label, C.POINTER(u64), u8p, sz, u64, vp]
        L.orc_iknp_receive_bits.argtypes = [vp, vp, C.POINTER(u64), vp, u64, u8p, sz, C.POINTER(sz), vp]
        L.orc_iknp_send_bits.argtypes = [vp, _Label, C.POINTER(u64), u8p, sz, u64, vp]
        L.orc_cot_send.argtypes = [vp, _Label, _Label, vp, u64, vp]
        L.orc_cot_receive.argtypes = [vp, u8p, _Label, vp, u64]
        L.orc_rot_send.argtypes = [vp, _Label, _Label, vp, u64]
        L.orc_rot_receive.argtypes = [vp, _Label, u64]
        L.orc_mul128.argtypes = [_Label, _Label, C.POINTER(_Label), C.POINTER(_Label)]
        L.orc_inner_product.argtypes = [vp, vp, u64, C.POINTER(_Label), C.POINTER(_Label)]
        L.orc_iknp_check_sums.argtypes = [_Label, u64, vp, vp, u64, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _b(x) -> np.ndarray:
    return np.frombuffer(bytes(x), dtype=np.uint8).copy()


def _chk(rc):
    if rc != 0:
        raise OracleError(rc)


def L_(d0: int, d1: int) -> _Label:
    return _Label(d0 & (2**64 - 1), d1 & (2**64 - 1))


def lab(x) -> _Label:
    """numpy LABEL_DTYPE scalar / (d0,d1) tuple -> ctypes label."""
    if isinstance(x, _Label):
        return x
    return _Label(int(x[0]), int(x[1]))


def t(l: _Label):
    return (int(l.d0), int(l.d1))


# ---- primitives -----------------------------------------------------------
def aes_encrypt_block(key: bytes, block: bytes) -> bytes:
    k, i, o = _b(key), _b(block), np.zeros(16, np.uint8)
    _chk(lib().orc_aes_encrypt_block(_p(k), len(key), _p(i), _p(o)))
    return o.tobytes()


def mul2(l):
    return t(lib().orc_label_mul2(lab(l)))


def mul4(l):
    return t(lib().orc_label_mul4(lab(l)))


def encrypt_half(key: bytes, x, tweak: int):
    out = _Label()
    k = _b(key)
    _chk(lib().orc_encrypt_half(_p(k), len(key), lab(x), tweak, C.byref(out)))
    return t(out)


def encrypt(key: bytes, a, b, c, tw: int):
    out = _Label()
    k = _b(key)
    _chk(lib().orc_encrypt(_p(k), len(key), lab(a), lab(b), lab(c), tw, C.byref(out)))
    return t(out)


def decrypt(key: bytes, a, b, tw: int, c):
    out = _Label()
    k = _b(key)
    _chk(lib().orc_decrypt(_p(k), len(key), lab(a), lab(b), tw, lab(c), C.byref(out)))
    return t(out)


# ---- Garble / Eval ----------------------------------------------------------
def garble(circ: Circuit, key: bytes, rand: bytes):
    """Circuit.Garble.  Returns (R, wires[nwires], slab[rows], row_off[ngates+1])."""
    assert len(rand) >= 16 * (1 + circ.num_inputs)
    gates = np.ascontiguousarray(circ.gates)
    wires = np.zeros(circ.num_wires, WIRE_DTYPE)
    slab = np.zeros(max(circ.num_rows, 1), LABEL_DTYPE)
    off = np.zeros(circ.num_gates + 1, np.uint32)
    r = np.zeros(1, LABEL_DTYPE)
    k, rb = _b(key), _b(rand)
    _chk(lib().orc_garble(_p(gates), circ.num_gates, circ.num_wires, circ.num_inputs, _p(k), len(key),
                          _p(rb), _p(r), _p(wires), _p(slab), _p(off)))
    return r[0], wires, slab[: circ.num_rows], off


def eval_(circ: Circuit, key: bytes, in_labels: np.ndarray, slab: np.ndarray, row_off=None) -> np.ndarray:
    """Circuit.Eval.  Returns all nwires labels."""
    gates = np.ascontiguousarray(circ.gates)
    wires = np.zeros(circ.num_wires, LABEL_DTYPE)
    wires[: circ.num_inputs] = in_labels
    slab = np.ascontiguousarray(slab)
    k = _b(key)
    ro = None if row_off is None else np.ascontiguousarray(row_off, dtype=np.uint32)
    _chk(lib().orc_eval(_p(gates), circ.num_gates, circ.num_wires, _p(k), len(key), _p(wires), _p(slab), _p(ro)))
    return wires


def _keys(keys, batch):
    """bytes (shared) or uint8 [batch, keylen] -> (array, keylen, stride)."""
    if isinstance(keys, (bytes, bytearray)):
        return _b(keys), len(keys), 0
    keys = np.ascontiguousarray(keys, dtype=np.uint8)
    assert keys.shape[0] == batch
    return keys, keys.shape[1], keys.shape[1]


def garble_batch(circ: Circuit, keys, rand: np.ndarray, threads: int = 1):
    """rand: uint8 [batch, 16*(1+ninputs)].  Returns (R[batch], tables[batch,rows], io_wires[batch,nin+nout])."""
    batch = rand.shape[0]
    rand = np.ascontiguousarray(rand, dtype=np.uint8)
    assert rand.shape[1] == 16 * (1 + circ.num_inputs)
    gates = np.ascontiguousarray(circ.gates)
    ka, kl, ks = _keys(keys, batch)
    r = np.zeros(batch, LABEL_DTYPE)
    tables = np.zeros((batch, circ.num_rows), LABEL_DTYPE)
    io = np.zeros((batch, circ.num_inputs + circ.num_outputs), WIRE_DTYPE)
    _chk(lib().orc_garble_batch(_p(gates), circ.num_gates, circ.num_wires, circ.num_inputs, circ.num_outputs,
                                _p(ka), kl, ks, batch, _p(rand), _p(r), _p(tables), _p(io), threads))
    return r, tables, io


def eval_batch(circ: Circuit, keys, tables: np.ndarray, in_labels: np.ndarray, threads: int = 1) -> np.ndarray:
    batch = in_labels.shape[0]
    gates = np.ascontiguousarray(circ.gates)
    ka, kl, ks = _keys(keys, batch)
    tables = np.ascontiguousarray(tables)
    in_labels = np.ascontiguousarray(in_labels)
    out = np.zeros((batch, circ.num_outputs), LABEL_DTYPE)
    _chk(lib().orc_eval_batch(_p(gates), circ.num_gates, circ.num_wires, circ.num_inputs, circ.num_outputs,
                              _p(ka), kl, ks, batch, _p(tables), _p(in_labels), _p(out), threads))
    return out


# ---- streaming ---------------------------------------------------------------
class Streaming:
    """circuit.Streaming restated (NewStreaming / Garble / GetInput)."""

    def __init__(self, key: bytes, rand: bytes, input_ids):
        ids = np.ascontiguousarray(input_ids, dtype=np.uint32)
        assert len(rand) >= 16 * (1 + len(ids))
        k, rb = _b(key), _b(rand)
        self._h = lib().orc_stream_new(_p(k), len(key), _p(rb), _p(ids), len(ids))
        if not self._h:
            raise OracleError(-1)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_stream_free(self._h)
            self._h = None

    @property
    def r(self):
        return t(lib().orc_stream_r(self._h))

    def get_input(self, wid: int):
        w = lib().orc_stream_get_input(self._h, wid)
        return (t(w.l0), t(w.l1))

    def garble(self, circ: Circuit, in_ids, out_ids) -> bytes:
        gates = np.ascontiguousarray(circ.gates)
        i = np.ascontiguousarray(in_ids, dtype=np.uint32)
        o = np.ascontiguousarray(out_ids, dtype=np.uint32)
        cap = 13 * circ.num_gates + 16 * circ.num_rows + 64
        buf = np.zeros(cap, np.uint8)
        n = C.c_size_t(0)
        _chk(lib().orc_stream_garble(self._h, _p(gates), circ.num_gates, circ.num_wires, _p(i), len(i),
                                     _p(o), len(o), _p(buf), cap, C.byref(n)))
        return buf[: n.value].tobytes()


class StreamEval:
    def __init__(self, key: bytes):
        k = _b(key)
        self._h = lib().orc_seval_new(_p(k), len(key))
        if not self._h:
            raise OracleError(-1)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_seval_free(self._h)
            self._h = None

    def set(self, wid: int, label):
        lib().orc_seval_set(self._h, wid, lab(label))

    def get(self, wid: int):
        return t(lib().orc_seval_get(self._h, wid))

    def circuit(self, buf: bytes, ngates: int, ntmp: int, nwires: int) -> int:
        b = _b(buf)
        n = C.c_size_t(0)
        _chk(lib().orc_seval_circuit(self._h, _p(b), len(buf), ngates, ntmp, nwires, C.byref(n)))
        return n.value


# ---- IKNP ----------------------------------------------------------------------
def prg(key, pos: int, n: int) -> bytes:
    buf = np.zeros(max(n, 1), np.uint8)
    lib().orc_prg(lab(key), pos, _p(buf), n)
    return buf[:n].tobytes()


def create_labels(nl: int, buf: bytes, w: int) -> np.ndarray:
    out = np.zeros(nl, LABEL_DTYPE)
    b = _b(buf)
    lib().orc_create_labels(_p(out), nl, _p(b), w)
    return out


def u_size(n: int) -> int:
    """Bytes of U the receiver sends for n OTs (sum of byteRows*128 over chunks)."""
    full, rem = divmod(n, 512)
    return full * 8192 + ((rem + 7) // 8) * 128


def iknp_receive(k0: np.ndarray, k1: np.ndarray, pos: int, choice: np.ndarray):
    """Returns (U bytes, labels[n], new pos)."""
    n = len(choice)
    k0, k1 = np.ascontiguousarray(k0), np.ascontiguousarray(k1)
    ch = np.ascontiguousarray(choice, dtype=np.uint8)
    cap = u_size(n) + 16
    u = np.zeros(cap, np.uint8)
    res = np.zeros(max(n, 1), LABEL_DTYPE)
    p, ul = C.c_uint64(pos), C.c_size_t(0)
    _chk(lib().orc_iknp_receive(_p(k0), _p(k1), C.byref(p), _p(ch), n, _p(u), cap, C.byref(ul), _p(res)))
    return u[: ul.value].copy(), res[:n], p.value


def iknp_send(k: np.ndarray, delta, pos: int, u: np.ndarray, n: int):
    k = np.ascontiguousarray(k)
    u = np.ascontiguousarray(u, dtype=np.uint8)
    res = np.zeros(max(n, 1), LABEL_DTYPE)
    p = C.c_uint64(pos)
    _chk(lib().orc_iknp_send(_p(k), lab(delta), C.byref(p), _p(u), len(u), n, _p(res)))
    return res[:n], p.value


def iknp_receive_bits(k0, k1, pos: int, choices: np.ndarray, n: int):
    k0, k1 = np.ascontiguousarray(k0), np.ascontiguousarray(k1)
    ch = np.ascontiguousarray(choices, dtype=np.uint64)
    cap = u_size(n) + 16
    u = np.zeros(cap, np.uint8)
    res = np.zeros((n + 63) // 64, np.uint64)
    p, ul = C.c_uint64(pos), C.c_size_t(0)
    _chk(lib().orc_iknp_receive_bits(_p(k0), _p(k1), C.byref(p), _p(ch), n, _p(u), cap, C.byref(ul), _p(res)))
    return u[: ul.value].copy(), res, p.value


def iknp_send_bits(k, delta, pos: int, u: np.ndarray, n: int):
    k = np.ascontiguousarray(k)
    u = np.ascontiguousarray(u, dtype=np.uint8)
    res = np.zeros((n + 63) // 64, np.uint64)
    p = C.c_uint64(pos)
    _chk(lib().orc_iknp_send_bits(_p(k), lab(delta), C.byref(p), _p(u), len(u), n, _p(res)))
    return res, p.value


# ---- MiTCCRH / COT / ROT ------------------------------------------------------------
class _Mitccrh(C.Structure):
    _fields_ = [("batch_size", C.c_int), ("start", _Label), ("gid", C.c_uint64),
                ("key_used", C.c_int), ("keys", (C.c_uint8 * 16) * 64)]


class MITCCRH:
    def __init__(self, seed, batch_size: int):
        self._m = _Mitccrh()
        L = lib()
        L.orc_mitccrh_init.argtypes = [C.POINTER(_Mitccrh), _Label, C.c_int]
        L.orc_mitccrh_hash.argtypes = [C.POINTER(_Mitccrh), C.c_void_p, C.c_int, C.c_int]
        L.orc_mitccrh_init(C.byref(self._m), lab(seed), batch_size)

    def hash(self, blks: np.ndarray, k: int, h: int) -> None:
        assert blks.dtype == LABEL_DTYPE and blks.flags.c_contiguous and len(blks) == k * h
        _chk(lib().orc_mitccrh_hash(C.byref(self._m), _p(blks), k, h))


def cot_send(data, delta, seed, wires):
    n = len(data)
    data, wires = np.ascontiguousarray(data), np.ascontiguousarray(wires)
    out = np.zeros(2 * n, LABEL_DTYPE)
    lib().orc_cot_send(_p(data), lab(delta), lab(seed), _p(wires), n, _p(out))
    return out


def cot_receive(t_labels, flags, seed, msgs):
    res = np.ascontiguousarray(t_labels).copy()
    fl = np.ascontiguousarray(flags, dtype=np.uint8)
    msgs = np.ascontiguousarray(msgs)
    lib().orc_cot_receive(_p(res), _p(fl), lab(seed), _p(msgs), len(res))
    return res


def rot_send(data, delta, seed):
    n = len(data)
    data = np.ascontiguousarray(data)
    wires = np.zeros(n, WIRE_DTYPE)
    lib().orc_rot_send(_p(data), lab(delta), lab(seed), _p(wires), n)
    return wires


def rot_receive(t_labels, seed):
    res = np.ascontiguousarray(t_labels).copy()
    lib().orc_rot_receive(_p(res), lab(seed), len(res))
    return res


def mul128(a, b):
    lo, hi = _Label(), _Label()
    lib().orc_mul128(lab(a), lab(b), C.byref(lo), C.byref(hi))
    return t(lo), t(hi)


def inner_product(a: np.ndarray, b: np.ndarray):
    lo, hi = _Label(), _Label()
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    lib().orc_inner_product(_p(a), _p(b), min(len(a), len(b)), C.byref(lo), C.byref(hi))
    return t(lo), t(hi)


def iknp_check_sums(seed2, chi_start: int, labels: np.ndarray, choice=None):
    """(lo, hi, x) of the malicious-mode check (ot/iknp.go:150-173, :408-451)."""
    labels = np.ascontiguousarray(labels)
    out = (_Label * 3)()
    ch = None if choice is None else np.ascontiguousarray(choice, dtype=np.uint8)
    lib().orc_iknp_check_sums(lab(seed2), chi_start, _p(labels), None if ch is None else _p(ch), len(labels), out)
    return t(out[0]), t(out[1]), t(out[2])
