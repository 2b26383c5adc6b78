"""Test infrastructure: the reference's sha2pc protocol (sha2pc/garbler.go, evaluator.go, encoding.go, ot/co_helpers.go)
re-run with the deterministic readers of sha2pc/sha2pc_test.go:74-131 (TestDeterministicTranscript), with the garbling
step pluggable.  The reference's test pins SHA-256 hashes of the three encoded rounds; round 3 contains the AES key, all
42,914 garbled rows, the garbler's input labels, both labels of every output wire and the OT ciphertexts of both labels
of every evaluator input -- so reproducing its hash pins the bytes Circuit.Garble produces (rows, row order, input and
output labels) against the reference itself."""
import hashlib

import numpy as np

import gostd

EXP_ROUND1 = "0191a7115a2ae1a1ff5ef7c9dbc5cf1078049b9e8fb77270b6b3c8f033220174"   # sha2pc_test.go:118-123
EXP_ROUND2 = "ff6286651743fff6b5b98857425fd11b9b2f877bb54258230054fdbe16575c84"
EXP_ROUND3 = "ae10edf7fdb70a039b817cbacd9acf5069e3b019cf0d33eb48754536b2a7af39"
EXP_FINAL = "4b2f74579fc7c778745121996f604371a326dc5174f9851706032626668abf2e"


class DeterministicReader:
    """sha2pc_test.go:473-494: math/rand seeded with the first 8 bytes of SHA-256(seed); one Intn(256) per byte."""

    def __init__(self, seed: bytes):
        self.src = gostd.GoRand(int.from_bytes(hashlib.sha256(seed).digest()[:8], "big"))

    def read(self, n: int) -> bytes:
        return bytes(self.src.intn_pow2(256) for _ in range(n))


def _fixed(v: int) -> bytes:
    return v.to_bytes(32, "big")


def _minimal(v: int) -> bytes:                                  # big.Int.Bytes()
    return v.to_bytes((v.bit_length() + 7) // 8, "big")


def _chunk(data: bytes) -> bytes:                               # encoding.go:663-668 (uvarint length; < 128 here)
    assert len(data) < 128
    return bytes([len(data)]) + data


def _mask(x: int, y: int, idx: int) -> bytes:                   # ot/co_helpers.go:217-230
    return hashlib.sha256(_minimal(x) + _minimal(y) + idx.to_bytes(8, "big")).digest()


def _label_bytes(lab) -> bytes:                                 # ot/label.go:105-108 (big-endian D0, D1)
    return int(lab["d0"]).to_bytes(8, "big") + int(lab["d1"]).to_bytes(8, "big")


def _bits_little(data: bytes):                                  # sha2pc/bits.go:4-15
    return [(b >> i) & 1 for b in data for i in range(8)]


def run(garble, evaluate=None):
    """garble(key32, rand_bytes) -> (input wires [512] of {l0,l1}, output wires [256], slab [rows]); rows in gate order.
    evaluate(key32, in_labels [512], slab) -> output labels [256] (optional: checks the final digest).
    Returns the hex hashes (round1, round2, round3, final)."""
    r1_rand, r3_rand, ev_rand = (DeterministicReader(s) for s in (b"garbler-round1", b"garbler-round3", b"evaluator-seed"))
    a = bytes(range(32))
    b = bytes(32 - i for i in range(32))
    # GarblerRound1 (garbler.go:36-74): CO sender setup, then the session id
    sa = gostd.crypto_rand_int(r1_rand.read, gostd.N)
    ax, ay = gostd.scalar_base_mult(sa)
    aax, aay = gostd.scalar_mult(ax, ay, sa)
    ainv = (aax, gostd.P - aay)
    sid = r1_rand.read(8)
    enc1 = b"R1" + sid + _chunk(b"P-256") + _fixed(ax) + _fixed(ay)             # encoding.go:35-49, 604-625
    # EvaluatorRound2 (evaluator.go:27-63, co_helpers.go:124-160)
    ebits = _bits_little(b)
    scalars, points = [], []
    for bit in ebits:
        s = gostd.crypto_rand_int(ev_rand.read, gostd.N)
        scalars.append(s)
        p = gostd.scalar_base_mult(s)
        if bit:
            p = gostd.add(p[0], p[1], ax, ay)
        points.append(p)
    signs = bytearray(32)
    for i, p in enumerate(points):
        if p[1] & 1:
            signs[i // 8] |= 1 << (i % 8)
    enc2 = b"R2" + sid + _chunk(b"P-256") + b"".join(_fixed(p[0]) for p in points) + bytes(signs)   # encoding.go:84-100, 481-503
    # GarblerRound3 (garbler.go:79-136): key, then Circuit.Garble reads R and one label per input wire
    key = r3_rand.read(32)
    rand = r3_rand.read(16 * (1 + 512))
    wires, hints, slab = garble(key, rand)
    gbits = _bits_little(a)
    garbler_labels = [wires[i]["l1"] if gbits[i] else wires[i]["l0"] for i in range(256)]
    cts = []
    for idx, (px, py) in enumerate(points):                                     # co_helpers.go:88-121
        bx, by = gostd.scalar_mult(px, py, sa)
        bax, bay = gostd.add(bx, by, *ainv)
        w = wires[256 + idx]
        m0, m1 = _mask(bx, by, idx), _mask(bax, bay, idx)
        cts.append(bytes(x ^ y for x, y in zip(m0, _label_bytes(w["l0"]))) + bytes(x ^ y for x, y in zip(m1, _label_bytes(w["l1"]))))
    enc3 = (b"R3" + sid + key + b"".join(_label_bytes(l) for l in slab) + b"".join(_label_bytes(l) for l in garbler_labels)
            + b"".join(_label_bytes(w["l0"]) + _label_bytes(w["l1"]) for w in hints) + b"".join(cts))   # encoding.go:149-174
    assert (len(enc1), len(enc2), len(enc3)) == (80, 8240, 707146)              # sha2pc_test.go:236-240 (P-256 sizes)
    final = None
    if evaluate is not None:                                                    # EvaluatorRound4 (evaluator.go:67-114)
        in_labels = np.zeros(512, dtype=slab.dtype)
        in_labels[:256] = garbler_labels
        for idx in range(256):
            sx, sy = gostd.scalar_mult(ax, ay, scalars[idx])
            ct = cts[idx][16:] if ebits[idx] else cts[idx][:16]
            pt = bytes(x ^ y for x, y in zip(_mask(sx, sy, idx), ct))
            in_labels[256 + idx] = (int.from_bytes(pt[:8], "big"), int.from_bytes(pt[8:], "big"))
        out = evaluate(key, in_labels, slab)
        digest = bytearray(32)
        for i in range(256):
            if out[i] == hints[i]["l1"]:
                digest[i // 8] |= 1 << (i % 8)
            else:
                assert out[i] == hints[i]["l0"], "unknown output label"
        final = bytes(digest).hex()
    h = lambda d: hashlib.sha256(d).hexdigest()
    return h(enc1), h(enc2), h(enc3), final
