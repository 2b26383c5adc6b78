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


# ---- COT / ROT post-processing and the malicious-mode check on the device -------------
def _rand_wires(tag, n):
    from mpc_b200.circuit_io import WIRE_DTYPE
    w = np.zeros(n, WIRE_DTYPE)
    a, b = drbg_labels(tag + "/l0", n), drbg_labels(tag + "/l1", n)
    w["l0"], w["l1"] = a, b
    return w


@pytest.mark.parametrize("n", [1, 7, 8, 9, 64, 1000, 4096 + 5, (1 << 16) + 3])
def test_cot_rot_bit_exact(n):
    """COT.Send / Receive and ROT.Send / Receive after the extension (ot/cot.go:157-233,
    ot/rot.go:155-197) against the oracle's batch-of-eight loops, incl. partial last batches."""
    from mpc_b200.ot import cot_receive, cot_send, rot_receive, rot_send
    k0, k1, delta, ks = _keys("cot")
    b = (DRBG(f"cot/b/{n}").array(n) & 1).astype(np.uint8)
    u, t = IKNPReceiver(k0, k1).receive(b)
    q = IKNPSender(ks, delta).send(u, n)
    seed = drbg_labels(f"cot/seed/{n}", 1)
    wires = _rand_wires(f"cot/w/{n}", n)
    msgs = cot_send(seed, delta, q, wires)
    assert eq(msgs, O.cot_send(q, delta[0], seed[0], wires)), "COT messages differ"
    got = cot_receive(seed, b, msgs, t)
    assert eq(got, O.cot_receive(t, b, seed[0], msgs)), "COT results differ"
    want = np.where(b.astype(bool), wires["l1"], wires["l0"])
    assert eq(got, want), "COT does not deliver the chosen labels"
    # SendLabel byte encoding on both sides (ot/label.go:105-114)
    mb = cot_send(seed, delta, q, wires, wire_bytes=True)
    be = np.stack([msgs["d0"].astype(">u8").view(np.uint8).reshape(-1, 8),
                   msgs["d1"].astype(">u8").view(np.uint8).reshape(-1, 8)], axis=1).reshape(-1, 16)
    assert eq(mb, be), "wire-byte messages differ"
    assert eq(cot_receive(seed, b, mb, t, wire_bytes=True), want)
    rw = rot_send(seed, delta, q)
    assert eq(rw, O.rot_send(q, delta[0], seed[0])), "ROT wires differ"
    rr = rot_receive(seed, t)
    assert eq(rr, O.rot_receive(t, seed[0])), "ROT results differ"
    assert eq(rr, np.where(b.astype(bool), rw["l1"], rw["l0"]))


@pytest.mark.parametrize("n", [1, 2, 100, 1024, 1025, 5000, (1 << 16) + 77])
def test_iknp_check_sums_bit_exact(n):
    """prgLabels + vectorInnPrdtSumNoRed / mul128 sums (ot/iknp.go:150-173, :408-451) vs the oracle."""
    from mpc_b200.ot import iknp_check_sums
    labels = drbg_labels(f"chk/l/{n}", n)
    choice = (DRBG(f"chk/b/{n}").array(n) & 1).astype(np.uint8)
    seed2 = drbg_labels(f"chk/seed/{n}", 1)
    for start in (0, 12345, (1 << 32) - 3):
        assert iknp_check_sums(seed2, start, labels, choice) == O.iknp_check_sums(seed2[0], start, labels, choice)
    lo, hi, x = iknp_check_sums(seed2, 9, labels)          # sender side: no choice vector
    assert (lo, hi) == O.iknp_check_sums(seed2[0], 9, labels)[:2] and x == (0, 0)


def test_iknp_malicious_check_full_size():
    """The whole malicious-mode exchange at 2^20 OTs on the device: extension, the 256 choice-vector OTs,
    both sides' sums, and the sender's verdict q ^ mul128(x, Delta) == t (ot/iknp.go:175-191)."""
    from mpc_b200.ot import iknp_check_sums, mul128
    n = 1 << 20
    k0, k1, delta, ks = _keys("mal")
    rcv, snd = IKNPReceiver(k0, k1), IKNPSender(ks, delta)
    b = (DRBG("mal/b").array(n) & 1).astype(np.uint8)
    bcv = (DRBG("mal/bcv").array(256) & 1).astype(np.uint8)
    u, t = rcv.receive(b)
    u2, t2 = rcv.receive(bcv)
    q = snd.send(u, n)
    q2 = snd.send(u2, 256)
    seed2 = drbg_labels("mal/seed2", 1)

    def fold(a, c):
        return tuple((p[0] ^ r[0], p[1] ^ r[1]) for p, r in zip(a, c))
    t0, t1, x = fold(iknp_check_sums(seed2, 0, t, b), iknp_check_sums(seed2, n, t2, bcv))
    q0, q1, _ = fold(iknp_check_sums(seed2, 0, q), iknp_check_sums(seed2, n, q2))
    d = (int(delta["d0"][0]), int(delta["d1"][0]))
    r0, r1 = mul128(x, d)
    assert mul128(x, d) == O.mul128(x, d)
    assert (q0[0] ^ r0[0], q0[1] ^ r0[1]) == t0 and (q1[0] ^ r1[0], q1[1] ^ r1[1]) == t1
    # a receiver that lies about one choice bit is caught
    b[12345] ^= 1
    c0, c1, cx = fold(iknp_check_sums(seed2, 0, t, b), iknp_check_sums(seed2, n, t2, bcv))
    r0, r1 = mul128(cx, d)
    assert ((q0[0] ^ r0[0], q0[1] ^ r0[1]), (q1[0] ^ r1[0], q1[1] ^ r1[1])) != (c0, c1)


def test_cot_chain_stays_on_the_device_through_the_c_abi_alone():
    """IKNP expansion -> COT post-processing on both sides with every intermediate label left in HBM, using
    only C-ABI calls (gcb_dev_alloc / upload / download, the _dev entry points and one library stream) --
    what a Go caller without a CUDA binding does.  The receiver ends up with the chosen wire labels."""
    import ctypes as C
    from mpc_b200._lib import Label, check, ptr
    from mpc_b200.ot import stream_advance, u_size
    L = _lib.lib()
    n = 5000
    k0, k1, delta, ks = _keys("chain")
    b = (DRBG("chain/b").array(n) & 1).astype(np.uint8)
    wires = _rand_wires("chain/w", n)
    seed = drbg_labels("chain/seed", 1)
    st = C.c_void_p()
    check(L.gcb_dev_stream_create(C.byref(st)))

    def dev(nbytes, src=None):
        p = L.gcb_dev_alloc(nbytes)
        assert p
        if src is not None:
            check(L.gcb_dev_upload(p, ptr(np.ascontiguousarray(src)), nbytes, st))
        return p

    d_k0, d_k1, d_ks, d_delta = dev(2048, k0), dev(2048, k1), dev(2048, ks), dev(16, delta)
    d_b, d_w = dev(n, b), dev(32 * n, wires)
    ul = u_size(n)
    d_u, d_t, d_q, d_msgs, d_res = dev(ul), dev(16 * n), dev(16 * n), dev(32 * n), dev(16 * n)
    s_lab, d_lab = Label(int(seed["d0"][0]), int(seed["d1"][0])), Label(int(delta["d0"][0]), int(delta["d1"][0]))
    check(L.gcb_iknp_receiver_expand_dev(d_k0, d_k1, 0, d_b, n, d_u, d_t, st))
    check(L.gcb_iknp_sender_expand_dev(d_ks, d_delta, 0, d_u, ul, n, d_q, st))
    check(L.gcb_cot_send_dev(C.byref(s_lab), C.byref(d_lab), d_q, d_w, n, d_msgs, 0, st))
    check(L.gcb_cot_receive_dev(C.byref(s_lab), d_b, d_msgs, d_t, n, d_res, 0, st))
    res = np.zeros(n, dtype=LABEL_DTYPE)
    check(L.gcb_dev_download(ptr(res), d_res, 16 * n, st))
    check(L.gcb_dev_sync(st))
    assert eq(res, np.where(b.astype(bool), wires["l1"], wires["l0"]))
    # and the host-pointer calls agree with the chain (same OT numbering, same seed)
    u, t = IKNPReceiver(k0, k1).receive(b)
    q = IKNPSender(ks, delta).send(u, n)
    from mpc_b200.ot import cot_receive, cot_send
    assert eq(cot_receive(seed, b, cot_send(seed, delta, q, wires), t), res)
    for p in (d_k0, d_k1, d_ks, d_delta, d_b, d_w, d_u, d_t, d_q, d_msgs, d_res):
        L.gcb_dev_free(p)
    L.gcb_dev_stream_destroy(st)
