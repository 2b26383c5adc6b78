"""Test infrastructure: the parts of Go's standard library that the reference's deterministic tests depend on,
restated from their published algorithms so that the reference's own golden hashes can be reproduced without a Go
toolchain (sha2pc/sha2pc_test.go:74-131 draws its randomness from math/rand and runs Chou-Orlandi OT on P-256).

* math/rand's rngSource: additive lagged Fibonacci generator x[n] = x[n-607] + x[n-273] mod 2^64, seeded through the
  Lehmer generator 48271 * x mod (2^31 - 1) and the 607-entry table rngCooked.  The table is not copied: it is "the
  state of the generator after 780e10 iterations" of the same recurrence from srand(1) (Go's gen_cooked.go), which is
  computed here by exponentiating x modulo the characteristic polynomial x^607 - x^334 - 1 over Z / 2^64.  Pinned by
  the well-known first outputs of rand.New(rand.NewSource(1)) (tests/test_reference_transcript.py).
* crypto/rand.Int(reader, max): big-endian rejection sampling with the top bits masked.
* crypto/elliptic P-256: affine results of ScalarBaseMult / ScalarMult / Add (plain Jacobian arithmetic on Python ints).

Nothing under mpc_b200/ imports this."""
import functools

import numpy as np

_L, _TAP = 607, 273
_M31 = (1 << 31) - 1
_MASK64 = (1 << 64) - 1


def _seedrand(x: int) -> int:
    return (48271 * x) % _M31


def _polymulmod(a, b):
    """a * b modulo x^607 - x^334 - 1, coefficients modulo 2^64."""
    out = np.zeros(2 * _L - 1, dtype=np.uint64)
    for i in range(_L):
        if a[i]:
            out[i:i + _L] += a[i] * b
    o = [int(v) for v in out]
    for d in range(2 * _L - 2, _L - 1, -1):
        c = o[d]
        if c:
            o[d - _TAP] = (o[d - _TAP] + c) & _MASK64
            o[d - _L] = (o[d - _L] + c) & _MASK64
    return np.array(o[:_L], dtype=np.uint64)


@functools.lru_cache(maxsize=None)
def rng_cooked():
    """gen_cooked.go: srand(1), 7.8e12 calls of vrand(), the vector as it then lies in memory."""
    with np.errstate(over="ignore"):
        x, vec = 1, [0] * _L
        for i in range(-20, _L):
            x = _seedrand(x)
            if i >= 0:
                u = x << 20
                x = _seedrand(x); u ^= x << 10
                x = _seedrand(x); u ^= x
                vec[i] = u
        n = int(7.8e12)
        # the sequence view: s[i] = s_{i-607}; vrand() step k reads positions feed = 333 - k and tap = 606 - k
        s = np.array([vec[(_L - _TAP - 1 - i) % _L] for i in range(_L)], dtype=np.uint64)
        r = np.zeros(_L, dtype=np.uint64); r[0] = 1
        base = np.zeros(_L, dtype=np.uint64); base[1] = 1
        e = n
        while e:
            if e & 1:
                r = _polymulmod(r, base)
            base = _polymulmod(base, base)
            e >>= 1
        out = []
        for _ in range(_L):                                   # s_{n-607} .. s_{n-1}
            out.append(int((r * s).sum(dtype=np.uint64)))
            c = r[_L - 1]
            r = np.roll(r, 1); r[0] = c
            r[_L - _TAP] += c
        feed = (_L - _TAP - n) % _L                           # holds s_{n-1}; feed + j holds s_{n-1-j}
        cooked = [0] * _L
        for j in range(_L):
            cooked[(feed + j) % _L] = out[_L - 1 - j]
        return tuple(cooked)


class GoRand:
    """rand.New(rand.NewSource(seed)): Int63, Intn for powers of two (what the deterministic readers use)."""

    def __init__(self, seed: int):
        seed = seed - (1 << 64) if seed >> 63 else seed       # int64
        rem = abs(seed) % _M31                                # Go's % truncates towards zero
        if seed < 0:
            rem = -rem
        if rem < 0:
            rem += _M31
        if rem == 0:
            rem = 89482311
        cooked = rng_cooked()
        x = rem
        self.vec = [0] * _L
        self.tap, self.feed = 0, _L - _TAP
        for i in range(-20, _L):
            x = _seedrand(x)
            if i >= 0:
                u = (x << 40) & _MASK64
                x = _seedrand(x); u ^= (x << 20) & _MASK64
                x = _seedrand(x); u ^= x
                self.vec[i] = u ^ cooked[i]

    def uint64(self) -> int:
        self.tap = (self.tap - 1) % _L
        self.feed = (self.feed - 1) % _L
        x = (self.vec[self.feed] + self.vec[self.tap]) & _MASK64
        self.vec[self.feed] = x
        return x

    def int63(self) -> int:
        return self.uint64() & ((1 << 63) - 1)

    def intn_pow2(self, n: int) -> int:
        """Intn(n) for n a power of two <= 2^31: Int31() & (n - 1), Int31 = Int63 >> 32."""
        return (self.int63() >> 32) & (n - 1)


def crypto_rand_int(read, maxv: int) -> int:
    """crypto/rand.Int(reader, max): uniform in [0, max)."""
    bitlen = (maxv - 1).bit_length()
    k = (bitlen + 7) // 8
    b = bitlen % 8 or 8
    while True:
        buf = bytearray(read(k))
        buf[0] &= (1 << b) - 1
        v = int.from_bytes(buf, "big")
        if v < maxv:
            return v


# ---- P-256 (FIPS 186-4 D.1.2.3) ------------------------------------------------------------
P = 0xffffffff00000001000000000000000000000000ffffffffffffffffffffffff
N = 0xffffffff00000000ffffffffffffffffbce6faada7179e84f3b9cac2fc632551
B = 0x5ac635d8aa3a93e7b3ebbd55769886bc651d06b0cc53b0f63bce3c3e27d2604b
GX = 0x6b17d1f2e12c4247f8bce6e563a440f277037d812deb33a0f4a13945d898c296
GY = 0x4fe342e2fe1a7f9b8ee7eb4a7c0f9e162bce33576b315ececbb6406837bf51f5


def _dbl(p):
    x, y, z = p
    if not z or not y:
        return (0, 1, 0)
    s = 4 * x * y * y % P
    m = (3 * (x - z * z) * (x + z * z)) % P                   # a = -3
    x2 = (m * m - 2 * s) % P
    return (x2, (m * (s - x2) - 8 * pow(y, 4, P)) % P, 2 * y * z % P)


def _add(p, q):
    if not p[2]:
        return q
    if not q[2]:
        return p
    x1, y1, z1 = p
    x2, y2, z2 = q
    u1, u2 = x1 * z2 * z2 % P, x2 * z1 * z1 % P
    s1, s2 = y1 * pow(z2, 3, P) % P, y2 * pow(z1, 3, P) % P
    if u1 == u2:
        return _dbl(p) if s1 == s2 else (0, 1, 0)
    h, r = (u2 - u1) % P, (s2 - s1) % P
    x3 = (r * r - h * h * h - 2 * u1 * h * h) % P
    return (x3, (r * (u1 * h * h - x3) - s1 * h * h * h) % P, h * z1 * z2 % P)


def _affine(p):
    if not p[2]:
        return (0, 0)
    zi = pow(p[2], -1, P)
    return (p[0] * zi * zi % P, p[1] * zi * zi * zi % P)


def scalar_mult(x: int, y: int, k: int):
    acc, q = (0, 1, 0), (x, y, 1)
    k %= N
    while k:
        if k & 1:
            acc = _add(acc, q)
        q = _dbl(q)
        k >>= 1
    return _affine(acc)


def scalar_base_mult(k: int):
    return scalar_mult(GX, GY, k)


def add(x1, y1, x2, y2):
    return _affine(_add((x1, y1, 1), (x2, y2, 1)))


def on_curve(x, y) -> bool:
    return (y * y - (x * x * x - 3 * x + B)) % P == 0
