"""GPU parity for the OT-extension path: IKNP expansion (AES-CTR column PRG +
bit-matrix transpose), MiTCCRH and the gate hash, against the CPU oracle."""
import numpy as np
import pytest

from mpc_b200 import _lib
from mpc_b200.circuit import hash_half
from mpc_b200.circuit_io import LABEL_DTYPE
from mpc_b200.ot import MITCCRH, IKNPReceiver, IKNPSender, mitccrh_hash_many, stream_advance, u_size
from oracle import pyoracle as O
from util import DRBG, drbg_labels, eq

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("klen", [16, 24, 32])
def test_hash_half_bit_exact(klen):
    key = DRBG(f"hh/{klen}").read(klen)
    x = drbg_labels(f"hh/x/{klen}", 5000)
    got = hash_half(key, x, tweak0=0xfffffff0)          # tweak wraps as uint32
    for i in list(range(40)) + [4999]:
        want = O.encrypt_half(key, x[i], (0xfffffff0 + i) & 0xffffffff)
        assert (int(got[i]["d0"]), int(got[i]["d1"])) == want


def test_mitccrh_golden_blocks_from_reference():
    """ot/mitccrh_test.go:16-33: seed 0, batch 8, K=8, H=2 on zero blocks; block 2i = AES_{BE(i||0)}(0)."""
    m = MITCCRH(np.zeros(1, LABEL_DTYPE), 8)
    blks = np.zeros(16, dtype=LABEL_DTYPE)
    m.hash(blks, 8, 2)
    hexes = [f"{int(b['d0']):016x}{int(b['d1']):016x}" for b in blks]
    assert hexes[0] == "66e94bd4ef8a2c3b884cfa59ca342b2e"
    assert hexes[2] == "f6b7bdd1caeebab574683893c4475484"
    om = O.MITCCRH(np.zeros(1, LABEL_DTYPE)[0], 8)
    ob = np.zeros(16, dtype=LABEL_DTYPE)
    om.hash(ob, 8, 2)
    assert eq(blks, ob)


@pytest.mark.parametrize("h", [1, 2, 3])
def test_mitccrh_many_keys_and_renewal(h):
    seed = drbg_labels("mit/seed", 1)
    n = 4099
    blks = drbg_labels(f"mit/blks/{h}", n * h)
    want = blks.copy()
    om = O.MITCCRH(seed[0], 1)
    for i in range(0, n, 1):                             # oracle: one key per Hash call
        om.hash(want[i * h:(i + 1) * h], 1, h)
    got = blks.copy()
    mitccrh_hash_many(seed, 0, got, n, h)
    assert eq(got, want)
    # the stateful mirror with batch 8, K = 8 (ot/cot.go:47)
    m, om8 = MITCCRH(seed, 8), O.MITCCRH(seed[0], 8)
    a, b = blks[: 64 * h].copy(), blks[: 64 * h].copy()
    for i in range(0, 64, 8):
        m.hash(a[i * h:(i + 8) * h], 8, h)
        om8.hash(b[i * h:(i + 8) * h], 8, h)
    assert eq(a, b)


def _keys(tag):
    k0, k1 = drbg_labels(tag + "/k0", 128), drbg_labels(tag + "/k1", 128)
    delta = drbg_labels(tag + "/delta", 1)
    dbits = [(int(delta["d0"][0]) >> i) & 1 if i < 64 else (int(delta["d1"][0]) >> (i - 64)) & 1 for i in range(128)]
    ks = np.where(np.array(dbits, dtype=bool), k1, k0).astype(LABEL_DTYPE)    # base OT outcome
    return k0, k1, delta, ks


@pytest.mark.parametrize("sizes", [[1], [129, 7, 512], [511, 512, 513], [1100, 64, 4096], [17, 1 << 16],
                                   [(1 << 16) + 300]])
def test_iknp_bit_exact_with_persistent_streams(sizes):
    k0, k1, delta, ks = _keys("iknp")
    rcv, snd = IKNPReceiver(k0, k1), IKNPSender(ks, delta)
    pos = 0
    for n in sizes:
        b = (DRBG(f"iknp/b/{n}").array(n) & 1).astype(np.uint8)
        u, t = rcv.receive(b)
        o_u, o_t, o_pos = O.iknp_receive(k0, k1, pos, b)
        assert len(u) == u_size(n) and eq(u, np.frombuffer(o_u, dtype=np.uint8)), f"U differs (n={n}, pos={pos})"
        assert eq(t, o_t), f"receiver labels differ (n={n}, pos={pos})"
        q = snd.send(u, n)
        o_q, o_pos2 = O.iknp_send(ks, delta[0], pos, np.frombuffer(o_u, dtype=np.uint8), n)
        assert eq(q, o_q), f"sender labels differ (n={n}, pos={pos})"
        # IKNP correlation t_j = q_j ^ b_j * Delta (ot/iknp_test.go:17-116)
        want = q.copy()
        want["d0"] ^= np.where(b.astype(bool), delta["d0"][0], 0).astype(np.uint64)
        want["d1"] ^= np.where(b.astype(bool), delta["d1"][0], 0).astype(np.uint64)
        assert eq(t, want)
        pos += stream_advance(n)
        assert pos == o_pos == o_pos2 == rcv.pos == snd.pos


def test_iknp_rejects_bad_chunk_size():
    k0, k1, delta, ks = _keys("iknp")
    snd = IKNPSender(ks, delta)
    with pytest.raises(_lib.GcbError, match="invalid chunk size") as e:
        snd.send(np.zeros(100, np.uint8), 512)
    assert e.value.rc == _lib.E_CHUNK


@pytest.mark.parametrize("n", [1, 64, 100, 511, 512, 1000, 4096 + 77, 1 << 16])
def test_iknp_bit_cot_bit_exact(n):
    """ReceiveBits / SendBits (gmw/triples.go's bit-COT): U bytes and packed result words vs the oracle,
    including the reference's whole-word choice XOR on short last chunks."""
    k0, k1, delta, ks = _keys("bits")
    rcv, snd = IKNPReceiver(k0, k1), IKNPSender(ks, delta)
    rcv.pos = snd.pos = 5
    words = (n + 63) // 64
    choices = np.frombuffer(DRBG(f"bits/c/{n}").read(8 * words), dtype=np.uint64).copy()
    u, r = rcv.receive_bits(choices, n)
    o_u, o_r, o_pos = O.iknp_receive_bits(k0, k1, 5, choices, n)
    assert eq(u, o_u), "U differs"
    assert np.array_equal(r, o_r), "receiver bits differ"
    s = snd.send_bits(u, n)
    o_s, _ = O.iknp_send_bits(ks, delta[0], 5, o_u, n)
    assert np.array_equal(s, o_s), "sender bits differ"
    assert rcv.pos == snd.pos == o_pos
