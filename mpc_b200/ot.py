"""Host-side mirror of the reference's ``ot`` package API for the hot path:
the IKNP expansion (ot/iknp.go) and MiTCCRH (ot/mitccrh.go).

Base OTs, framing and I/O stay with the caller, as in the reference split
described in include/gcb200.h; these classes keep the per-object state the Go
objects keep (the 128 / 256 PRG seeds, Delta and the byte-granular keystream
position) and make one device call per Send/Receive.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import GcbError, Label, check, ptr
from .circuit_io import LABEL_DTYPE

__all__ = ["IKNPSender", "IKNPReceiver", "MITCCRH", "u_size", "stream_advance", "GcbError",
           "cot_send", "cot_receive", "rot_send", "rot_receive", "iknp_check_sums", "mul128", "COT_WIRE_BYTES"]


def u_size(n: int) -> int:
    return int(_lib.lib().gcb_iknp_u_size(n))


def stream_advance(n: int) -> int:
    return int(_lib.lib().gcb_iknp_stream_advance(n))


class IKNPReceiver:
    """ot.IKNPReceiver after setup (ot/iknp.go:313-359): k0/k1 are the 128
    (L0, L1) seed pairs that were base-OT *sent*."""

    def __init__(self, k0: np.ndarray, k1: np.ndarray):
        self.k0 = np.ascontiguousarray(k0, dtype=LABEL_DTYPE)
        self.k1 = np.ascontiguousarray(k1, dtype=LABEL_DTYPE)
        assert len(self.k0) == 128 and len(self.k1) == 128
        self.pos = 0

    def receive(self, b: np.ndarray):
        """IKNPReceiver.receive (ot/iknp.go:468-511).  b: bool/uint8[n] choices.
        Returns (u, labels): the bytes handed to io.SendData chunk by chunk
        (concatenated) and result[n]."""
        b = np.ascontiguousarray(b, dtype=np.uint8)
        n = len(b)
        u = np.zeros(max(u_size(n), 16), dtype=np.uint8)
        labels = np.zeros(max(n, 1), dtype=LABEL_DTYPE)
        check(_lib.lib().gcb_iknp_receiver_expand(ptr(self.k0), ptr(self.k1), self.pos, ptr(b), n, ptr(u), ptr(labels)))
        self.pos += stream_advance(n)
        return u[: u_size(n)], labels[:n]


    def receive_bits(self, choices: np.ndarray, n: int):
        """IKNPReceiver.ReceiveBits (ot/iknp.go:554-620).  choices: packed uint64 words.
        Returns (u, result words)."""
        ch = np.ascontiguousarray(choices, dtype=np.uint64)
        if (n + 63) // 64 > len(ch):
            raise ValueError(f"choices buffer len={len(ch)} too short for n={n}")
        u = np.zeros(max(u_size(n), 16), dtype=np.uint8)
        res = np.zeros(max((n + 63) // 64, 1), dtype=np.uint64)
        check(_lib.lib().gcb_iknp_receiver_expand_bits(ptr(self.k0), ptr(self.k1), self.pos, ptr(ch), n, ptr(u), ptr(res)))
        self.pos += stream_advance(n)
        return u[: u_size(n)], res[: (n + 63) // 64]


class IKNPSender:
    """ot.IKNPSender after setup (ot/iknp.go:80-125): k are the 128 seeds
    received by base OT with choice bits Delta."""

    def __init__(self, k: np.ndarray, delta):
        self.k = np.ascontiguousarray(k, dtype=LABEL_DTYPE)
        assert len(self.k) == 128
        self.delta = np.ascontiguousarray(np.asarray(delta, dtype=LABEL_DTYPE).reshape(1))
        self.pos = 0

    def send(self, u: np.ndarray, n: int) -> np.ndarray:
        """IKNPSender.send (ot/iknp.go:197-226) over the received chunks."""
        u = np.ascontiguousarray(u, dtype=np.uint8)
        labels = np.zeros(max(n, 1), dtype=LABEL_DTYPE)
        check(_lib.lib().gcb_iknp_sender_expand(ptr(self.k), ptr(self.delta), self.pos, ptr(u) if len(u) else None,
                                                len(u), n, ptr(labels)))
        self.pos += stream_advance(n)
        return labels[:n]


    def send_bits(self, u: np.ndarray, n: int) -> np.ndarray:
        """IKNPSender.SendBits (ot/iknp.go:259-310): packed result words."""
        u = np.ascontiguousarray(u, dtype=np.uint8)
        res = np.zeros(max((n + 63) // 64, 1), dtype=np.uint64)
        check(_lib.lib().gcb_iknp_sender_expand_bits(ptr(self.k), ptr(self.delta), self.pos, ptr(u) if len(u) else None,
                                                     len(u), n, ptr(res)))
        self.pos += stream_advance(n)
        return res[: (n + 63) // 64]


class MITCCRH:
    """ot.MITCCRH (ot/mitccrh.go:50-128) with keys renewed on demand."""

    def __init__(self, seed, batch_size: int):
        s = np.asarray(seed, dtype=LABEL_DTYPE).reshape(1)
        self.seed = Label(int(s["d0"][0]), int(s["d1"][0]))
        self.batch_size = batch_size
        self.gid = 0
        self.key_used = batch_size           # forces renewal on first Hash (mitccrh.go:61-68)

    def hash(self, blks: np.ndarray, k: int, h: int) -> None:
        """Hash(blks, K, H) in place.  Panics of the reference become ValueError."""
        if k > self.batch_size:
            raise ValueError("MITCCRH.Hash: K > batchSize")
        if k > 0 and self.batch_size % k != 0:
            raise ValueError("MITCCRH.Hash: batchSize % K != 0")
        if len(blks) != k * h:
            raise ValueError("MITCCRH.Hash: len(blks) != K*H")
        if self.key_used == self.batch_size:       # renewKeys, mitccrh.go:70-89
            self.gid_base = self.gid
            self.gid += self.batch_size
            self.key_used = 0
        assert blks.dtype == LABEL_DTYPE and blks.flags["C_CONTIGUOUS"]
        check(_lib.lib().gcb_mitccrh_hash(C.byref(self.seed), self.gid_base + self.key_used, ptr(blks), k, h))
        self.key_used += k


def mitccrh_hash_many(seed, gid_start: int, blks: np.ndarray, nkeys: int, h: int) -> None:
    """One device call over ``nkeys`` consecutive keys (what COT/ROT batches add up to)."""
    s = np.asarray(seed, dtype=LABEL_DTYPE).reshape(1)
    lab = Label(int(s["d0"][0]), int(s["d1"][0]))
    assert blks.dtype == LABEL_DTYPE and blks.flags["C_CONTIGUOUS"] and len(blks) == nkeys * h
    check(_lib.lib().gcb_mitccrh_hash(C.byref(lab), gid_start, ptr(blks), nkeys, h))


# ---- COT / ROT post-processing and the malicious-mode check (ot/cot.go, ot/rot.go, ot/iknp.go) ----
COT_WIRE_BYTES = 1


def _label(x) -> Label:
    s = np.asarray(x, dtype=LABEL_DTYPE).reshape(1)
    return Label(int(s["d0"][0]), int(s["d1"][0]))


def cot_send(seed, delta, q: np.ndarray, wires: np.ndarray, wire_bytes: bool = False) -> np.ndarray:
    """The 2n messages COT.Send puts on the wire after the extension (ot/cot.go:157-181).
    wire_bytes: return the SendLabel byte encoding (uint8[2n, 16]) instead of labels."""
    from .circuit_io import WIRE_DTYPE
    q = np.ascontiguousarray(q, dtype=LABEL_DTYPE)
    wires = np.ascontiguousarray(wires, dtype=WIRE_DTYPE)
    n = len(q)
    assert len(wires) == n
    msgs = np.zeros(max(2 * n, 1), dtype=LABEL_DTYPE)
    s, d = _label(seed), _label(delta)
    check(_lib.lib().gcb_cot_send(C.byref(s), C.byref(d), ptr(q), ptr(wires), n, ptr(msgs), COT_WIRE_BYTES if wire_bytes else 0))
    msgs = msgs[: 2 * n]
    return msgs.view(np.uint8).reshape(-1, 16) if wire_bytes else msgs


def cot_receive(seed, choice: np.ndarray, msgs: np.ndarray, t: np.ndarray, wire_bytes: bool = False) -> np.ndarray:
    """COT.Receive after the extension (ot/cot.go:201-233): result[j] = msgs[2j + b_j] ^ H_j(t_j)."""
    t = np.ascontiguousarray(t, dtype=LABEL_DTYPE)
    n = len(t)
    choice = np.ascontiguousarray(choice, dtype=np.uint8)
    msgs = np.ascontiguousarray(msgs)
    assert msgs.nbytes == 32 * n and len(choice) == n
    res = np.zeros(max(n, 1), dtype=LABEL_DTYPE)
    s = _label(seed)
    check(_lib.lib().gcb_cot_receive(C.byref(s), ptr(choice), ptr(msgs), ptr(t), n, ptr(res), COT_WIRE_BYTES if wire_bytes else 0))
    return res[:n]


def rot_send(seed, delta, q: np.ndarray) -> np.ndarray:
    """ROT.Send after the extension (ot/rot.go:155-172): wires[j] = {H_j(q_j), H_j(q_j ^ delta)}."""
    from .circuit_io import WIRE_DTYPE
    q = np.ascontiguousarray(q, dtype=LABEL_DTYPE)
    n = len(q)
    wires = np.zeros(max(n, 1), dtype=WIRE_DTYPE)
    s, d = _label(seed), _label(delta)
    check(_lib.lib().gcb_rot_send(C.byref(s), C.byref(d), ptr(q), n, ptr(wires)))
    return wires[:n]


def rot_receive(seed, t: np.ndarray) -> np.ndarray:
    """ROT.Receive after the extension (ot/rot.go:192-197): result[j] = H_j(t_j)."""
    t = np.ascontiguousarray(t, dtype=LABEL_DTYPE)
    n = len(t)
    res = np.zeros(max(n, 1), dtype=LABEL_DTYPE)
    s = _label(seed)
    check(_lib.lib().gcb_rot_receive(C.byref(s), ptr(t), n, ptr(res)))
    return res[:n]


def iknp_check_sums(seed2, chi_start: int, labels: np.ndarray, choice=None):
    """(lo, hi, x) sums of the malicious-mode check (ot/iknp.go:150-173, :408-451) as (d0, d1) tuples."""
    labels = np.ascontiguousarray(labels, dtype=LABEL_DTYPE)
    ch = None if choice is None else np.ascontiguousarray(choice, dtype=np.uint8)
    out = np.zeros(3, dtype=LABEL_DTYPE)
    s = _label(seed2)
    check(_lib.lib().gcb_iknp_check_sums(C.byref(s), chi_start, ptr(labels) if len(labels) else None,
                                          None if ch is None else ptr(ch), len(labels), ptr(out)))
    return tuple((int(o["d0"]), int(o["d1"])) for o in out)


def mul128(a, b):
    """ot.mul128 (ot/mul128_generic.go): 256-bit carry-less product of two labels as ((lo.d0, lo.d1), (hi.d0, hi.d1)).
    Host arithmetic on Python integers: one product per check, nothing to put on the device."""
    def poly(l):
        return int(l[0]) | (int(l[1]) << 64)
    x, y, r = poly(a), poly(b), 0
    while y:
        low = y & -y
        r ^= x * low
        y ^= low
    m = (1 << 64) - 1
    return (r & m, (r >> 64) & m), ((r >> 128) & m, (r >> 192) & m)
