"""Deterministic byte source for synthetic bench / test inputs.

``DRBG(tag)`` = AES-128-CTR keystream, key = first 16 bytes of
``SHA-256("gcb200-bench-v1/" + tag)``, zero IV.  Bytes are consumed in the
order the reference's callers read their ``io.Reader`` (key where the caller
draws it, then R, then one L0 per input wire; circuit/garble.go:253-278).
Host-side only; not part of the data path.
"""
from __future__ import annotations

import hashlib

import numpy as np
from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes

PREFIX = "gcb200-bench-v1/"


class DRBG:
    def __init__(self, tag: str):
        key = hashlib.sha256((PREFIX + tag).encode()).digest()[:16]
        self._e = Cipher(algorithms.AES(key), modes.CTR(b"\0" * 16)).encryptor()

    def read(self, n: int) -> bytes:
        return self._e.update(b"\0" * n)

    def array(self, n: int) -> np.ndarray:
        return np.frombuffer(self.read(n), dtype=np.uint8).copy()


def garble_inputs(prefix: str, batch: int, ninputs: int, keylen: int = 0):
    """Per-instance reader streams ``DRBG(prefix/<i>)``.

    Returns (keys uint8[batch, keylen] or None, rand uint8[batch, 16*(1+ninputs)]);
    the key is drawn first when keylen != 0 (circuit/garbler.go:47-53 order).
    """
    rs = 16 * (1 + ninputs)
    rand = np.empty((batch, rs), dtype=np.uint8)
    keys = np.empty((batch, keylen), dtype=np.uint8) if keylen else None
    for i in range(batch):
        d = DRBG(f"{prefix}/{i}")
        if keylen:
            keys[i] = np.frombuffer(d.read(keylen), dtype=np.uint8)
        rand[i] = np.frombuffer(d.read(rs), dtype=np.uint8)
    return keys, rand
