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

__all__ = ["IKNPSender", "IKNPReceiver", "MITCCRH", "u_size", "stream_advance", "GcbError"]


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
