"""Independent pure-Python restatement of the reference algorithms, used to
cross-check the C oracle (a second implementation sharing no code with it).

AES comes from OpenSSL through the ``cryptography`` package; labels are Python
ints V = D0*2^64 + D1, i.e. the big-endian value of ``Label.GetData``
(ot/label.go:105-108).  Only for small inputs: everything is a Python loop.

Reference lines followed: circuit/garble.go:40-143,248-482; circuit/eval.go;
circuit/stream_garble.go:131-157,195-449; ot/iknp.go:197-226,468-511,622-683;
ot/mitccrh.go:70-128.
"""
from __future__ import annotations

import struct

from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes

M128 = (1 << 128) - 1
SBIT = 1 << 127
XOR, XNOR, AND, OR, INV = range(5)


class Aes:
    def __init__(self, key: bytes):
        if len(key) not in (16, 24, 32):
            raise ValueError("crypto/aes: invalid key size")
        self._e = Cipher(algorithms.AES(key), modes.ECB()).encryptor()

    def enc(self, v: int) -> int:
        return int.from_bytes(self._e.update(v.to_bytes(16, "big")), "big")


def S(v: int) -> int:
    return v >> 127


def h1(alg: Aes, x: int, i: int) -> int:
    """encryptHalf: K = 2x ^ i; AES(K) ^ K."""
    k = ((x << 1) & M128) ^ i
    return alg.enc(k) ^ k


def h2(alg: Aes, a: int, b: int, t: int) -> int:
    """encrypt with c = 0: K = 2a ^ 4b ^ t; AES(K) ^ K."""
    k = ((a << 1) & M128) ^ ((b << 2) & M128) ^ t
    return alg.enc(k) ^ k


def to_v(lbl) -> int:
    """numpy LABEL_DTYPE scalar or (d0, d1) -> int."""
    return (int(lbl[0]) << 64) | int(lbl[1])


def from_v(v: int):
    return (v >> 64, v & ((1 << 64) - 1))


def garble_gate(alg: Aes, op: int, a, b, r: int, tid: int):
    """Returns (c0, c1, rows, new tweak id); a, b are (L0, L1) pairs."""
    a0, a1 = a
    if op == XOR:
        c0 = a0 ^ b[0]
        return c0, c0 ^ r, [], tid
    if op == XNOR:
        c0 = a0 ^ b[0]
        return c0 ^ r, c0, [], tid
    if op == AND:
        b0, b1 = b
        pa, pb = S(a0), S(b0)
        j0, j1 = tid, tid + 1
        tg = h1(alg, a0, j0) ^ h1(alg, a1, j0) ^ (r if pb else 0)
        wg0 = h1(alg, a0, j0) ^ (tg if pa else 0)
        te = h1(alg, b0, j1) ^ h1(alg, b1, j1) ^ a0
        we0 = h1(alg, b0, j1) ^ ((te ^ a0) if pb else 0)
        c0 = wg0 ^ we0
        return c0, c0 ^ r, [tg, te], tid + 2
    if op == OR:
        b0, b1 = b
        tab = [0] * 4
        for x in (a0, a1):
            for y in (b0, b1):
                tab[2 * S(x) + S(y)] = h2(alg, x, y, tid)
        l0i = 2 * S(a0) + S(b0)
        c0 = c1 = tab[0]
        if l0i == 0:
            c1 ^= r
        else:
            c0 ^= r
        tab = [tab[i] ^ (c0 if i == l0i else c1) for i in range(4)]
        return c0, c1, tab[1:], tid + 1
    if op == INV:
        tab = [0] * 2
        tab[S(a0)] = h2(alg, a0, 0, tid)
        tab[S(a1)] = h2(alg, a1, 0, tid)
        l0i = S(a0)
        c0 = c1 = tab[0]
        if l0i == 0:
            c0 ^= r
        else:
            c1 ^= r
        tab = [tab[i] ^ (c1 if i == l0i else c0) for i in range(2)]
        return c0, c1, tab[1:], tid + 1
    raise ValueError("invalid gate type")


def eval_gate(alg: Aes, op: int, a: int, b: int, rows, tid: int):
    if op in (XOR, XNOR):
        return a ^ b, tid
    if op == AND:
        if len(rows) != 2:
            raise ValueError("corrupted ciruit: AND row length")
        wg = h1(alg, a, tid) ^ (rows[0] if S(a) else 0)
        we = h1(alg, b, tid + 1) ^ ((rows[1] ^ a) if S(b) else 0)
        return wg ^ we, tid + 2
    if op == OR:
        i = 2 * S(a) + S(b)
        c = rows[i - 1] if i > 0 else 0
        return c ^ h2(alg, a, b, tid), tid + 1
    if op == INV:
        c = rows[0] if S(a) else 0
        return c ^ h2(alg, a, 0, tid), tid + 1
    raise ValueError("invalid operation")


def garble(circ, key: bytes, rand: bytes):
    """Circuit.Garble -> (R, wires [(l0,l1)], per-gate row lists)."""
    alg = Aes(key)
    r = int.from_bytes(rand[:16], "big") | SBIT
    wires = [None] * circ.num_wires
    for i in range(circ.num_inputs):
        l0 = int.from_bytes(rand[16 * (i + 1):16 * (i + 2)], "big")
        wires[i] = (l0, l0 ^ r)
    tid = 0
    tables = []
    g = circ.gates
    for a, b, c, op in zip(g["in0"].tolist(), g["in1"].tolist(), g["out"].tolist(), g["op"].tolist()):
        c0, c1, rows, tid = garble_gate(alg, op, wires[a], wires[b] if op != INV else None, r, tid)
        wires[c] = (c0, c1)
        tables.append(rows)
    return r, wires, tables


def evaluate(circ, key: bytes, in_labels, tables):
    alg = Aes(key)
    wires = [0] * circ.num_wires
    wires[: circ.num_inputs] = list(in_labels)
    tid = 0
    g = circ.gates
    for i, (a, b, c, op) in enumerate(zip(g["in0"].tolist(), g["in1"].tolist(), g["out"].tolist(), g["op"].tolist())):
        wires[c], tid = eval_gate(alg, op, wires[a], wires[b] if op != INV else 0, tables[i], tid)
    return wires


def stream_garble(alg: Aes, r: int, perm: dict, circ, ins, outs) -> bytes:
    """Streaming.Garble for one sub-circuit on the permanent wire dict `perm`."""
    nin, first_out = len(ins), circ.num_wires - len(outs)
    tmp = {}

    def loc(w):
        if w < nin:
            return ins[w], False
        if w >= first_out:
            return outs[w - first_out], False
        return w, True

    def get(w):
        i, t = loc(w)
        return (tmp[i] if t else perm[i]), i, t

    out = bytearray()
    tid = 0
    g = circ.gates
    for a, b, c, op in zip(g["in0"].tolist(), g["in1"].tolist(), g["out"].tolist(), g["op"].tolist()):
        wb = bi = bt = None
        if op != INV:
            wb, bi, bt = get(b)
        wa, ai, at = get(a)
        c0, c1, rows, tid = garble_gate(alg, op, wa, wb, r, tid)
        ci, ct = loc(c)
        (tmp if ct else perm)[ci] = (c0, c1)
        hdr = op | (0x80 if at else 0) | (0x40 if bt else 0) | (0x20 if ct else 0)
        idx = [ai, ci] if op == INV else [ai, bi, ci]
        if all(i <= 0xffff for i in idx):
            out.append(hdr | 0x10)
            out += b"".join(struct.pack(">H", i) for i in idx)
        else:
            out.append(hdr)
            out += b"".join(struct.pack(">I", i) for i in idx)
        for row in rows:
            out += row.to_bytes(16, "big")
    return bytes(out)


# ---- IKNP -----------------------------------------------------------------------
class Prg:
    """newPrg/prg: AES-128-CTR, zero IV, stateful byte stream (Go cipher.NewCTR)."""

    def __init__(self, key_v: int):
        self._e = Cipher(algorithms.AES(key_v.to_bytes(16, "big")), modes.CTR(b"\0" * 16)).encryptor()

    def read(self, n: int) -> bytes:
        return self._e.update(b"\0" * n)


def label_bit(v: int, i: int) -> int:
    """Label.Bit(i): i<64 -> D0>>i, else D1>>(i-64)."""
    d0, d1 = v >> 64, v & ((1 << 64) - 1)
    return ((d1 >> (i - 64)) if i > 63 else (d0 >> i)) & 1


def create_labels(nl: int, buf: bytes, w: int):
    out = []
    for row in range(w * 8):
        if row >= nl:
            break
        d0 = d1 = 0
        for j in range(128):
            if (buf[j * w + row // 8] >> (row % 8)) & 1:
                if j < 64:
                    d0 |= 1 << j
                else:
                    d1 |= 1 << (j - 64)
        out.append((d0 << 64) | d1)
    return out


def iknp_receive(g0, g1, choice):
    """g0/g1: lists of 128 Prg objects (stateful).  Returns (chunks, labels)."""
    n = len(choice)
    bbuf = bytearray((n + 7) // 8)
    for i, f in enumerate(choice):
        if f:
            bbuf[i // 8] |= 1 << (i % 8)
    chunks, labels = [], []
    ofs = 0
    while ofs < n:
        rows = min(512, n - ofs)
        br = (rows + 7) // 8
        t0 = bytearray()
        u = bytearray()
        for i in range(128):
            a = g0[i].read(br)
            b = g1[i].read(br)
            t0 += a
            u += bytes(x ^ y ^ z for x, y, z in zip(a, b, bbuf[ofs // 8: ofs // 8 + br]))
        chunks.append(bytes(u))
        labels += create_labels(n - ofs, bytes(t0), br)
        ofs += rows
    return chunks, labels


def iknp_send(g0, delta_v: int, chunks, n: int):
    labels = []
    ofs = 0
    for chunk in chunks:
        assert len(chunk) % 128 == 0
        br = len(chunk) // 128
        t = bytearray()
        for i in range(128):
            a = g0[i].read(br)
            if label_bit(delta_v, i):
                a = bytes(x ^ y for x, y in zip(a, chunk[i * br:(i + 1) * br]))
            t += a
        labels += create_labels(n - ofs, bytes(t), br)
        ofs += br * 8
    return labels


def mitccrh_key(seed_v: int, gid: int) -> bytes:
    d0, d1 = seed_v >> 64, seed_v & ((1 << 64) - 1)
    return (((d0 ^ gid) << 64) | d1).to_bytes(16, "big")


def mitccrh_hash(seed_v: int, gid0: int, blks, h: int):
    """Key gid0+i hashes blocks i*h .. i*h+h-1: AES(x) ^ x."""
    out = []
    for i in range(len(blks) // h):
        alg = Aes(mitccrh_key(seed_v, gid0 + i))
        for j in range(h):
            x = blks[i * h + j]
            out.append(alg.enc(x) ^ x)
    return out
