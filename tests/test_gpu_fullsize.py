"""BASELINE.json's full sizes on the GPU, checked through size-independent properties and
sampled oracle comparisons: AES-128 x 4096 (config 2), SHA-256 streaming KAT (config 3),
IKNP 2^24 (config 4)."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import load_circuit
from mpc_b200.circuit import GarbleEngine, Streaming
from mpc_b200.circuit_io import LABEL_DTYPE
from mpc_b200.ot import IKNPReceiver, IKNPSender, stream_advance
from oracle import pyoracle as O
from util import DRBG, decode, drbg_labels, eq, garble_inputs, rand_to_labels, select

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("klen,shared", [(16, True), (32, False)])
def test_aes128_batch_4096(klen, shared):
    """Every instance decodes to OpenSSL AES(key, index); sampled instances equal the oracle bit for bit."""
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
    circ = load_circuit("aes_128")
    eng = GarbleEngine(circ)
    batch, nin = 4096, circ.num_inputs
    rng = np.random.default_rng(klen)
    rand = rng.integers(0, 256, (batch, 16 * (1 + nin)), dtype=np.uint8)
    keys = b"0123456789abcdef" if shared else rng.integers(0, 256, (batch, klen), dtype=np.uint8)
    r, l0 = rand_to_labels(rand, nin)
    tables, io = eng.garble_batch(keys, r, l0)
    pt_key = int.from_bytes(bytes(range(16)), "big")
    bits = np.zeros((batch, nin), dtype=np.uint8)
    bits[:, :128] = [(pt_key >> b) & 1 for b in range(128)]
    idx = np.arange(batch, dtype=np.uint64)
    for b in range(16):
        bits[:, 128 + b] = (idx >> np.uint64(b)) & np.uint64(1)
    out = eng.eval_batch(keys, tables, select(io[:, :nin], bits))
    ob = decode(io[:, nin:], out)
    assert ob.max() <= 1, "an output label is neither L0 nor L1"
    enc = Cipher(algorithms.AES(bytes(range(16))), modes.ECB()).encryptor()
    want = np.frombuffer(b"".join(enc.update(int(i).to_bytes(16, "big")) for i in range(batch)), dtype=np.uint8)
    got = np.packbits(ob[:, ::-1], axis=1)               # wire i = bit i of the big-endian integer
    assert np.array_equal(got.reshape(-1), want)
    sample = [0, 1, 591, 592, 2047, 4095]
    ks = keys if shared else keys[sample]
    _, o_tables, o_io = O.garble_batch(circ, ks, rand[sample])
    assert eq(tables[sample], o_tables) and eq(io[sample], o_io)
    # linearity of Free-XOR: L1 = L0 ^ R on every I/O wire, S(R) = 1
    R = io["l0"]["d0"][:, 0] ^ io["l1"]["d0"][:, 0]
    assert np.all(R >> np.uint64(63) == 1)
    assert np.all((io["l0"]["d0"] ^ io["l1"]["d0"]) == R[:, None])


def test_sha256_streaming_kat_abc():
    """sha256.circ(block = "abc" || pad, state = IV) decodes to ba7816bf...f20015ad after one streamed step."""
    circ = load_circuit("sha256")
    eng = GarbleEngine(circ)
    key = DRBG("kat256").read(32)
    ids = list(range(768))
    st = Streaming.new(DRBG("kat256/r").read(16 * 769), key, ids)
    out_ids = list(range(1000, 1256))
    buf, _, _ = st.garble(eng, ids, out_ids)
    inw, outw = st.get_inputs(ids)[0], st.get_inputs(out_ids)[0]
    block = b"abc" + b"\x80" + b"\0" * 52 + (24).to_bytes(8, "big")
    iv = bytes.fromhex("6a09e667bb67ae853c6ef372a54ff53a510e527f9b05688c1f83d9ab5be0cd19")
    bi, si = int.from_bytes(block, "big"), int.from_bytes(iv, "big")
    bits = [(bi >> i) & 1 for i in range(512)] + [(si >> i) & 1 for i in range(256)]
    se = O.StreamEval(key)
    for i, b in enumerate(bits):
        lab = inw[i]["l1"] if b else inw[i]["l0"]
        se.set(i, (int(lab["d0"]), int(lab["d1"])))
    used = se.circuit(buf[0].tobytes(), circ.num_gates, circ.num_wires, 1256)
    assert used == buf.shape[1]
    digest = 0
    for i, wid in enumerate(out_ids):
        l = se.get(wid)
        w0 = (int(outw[i]["l0"]["d0"]), int(outw[i]["l0"]["d1"]))
        w1 = (int(outw[i]["l1"]["d0"]), int(outw[i]["l1"]["d1"]))
        assert l in (w0, w1)
        digest |= (1 if l == w1 else 0) << i
    assert digest.to_bytes(32, "big") == hashlib.sha256(b"abc").digest()


@pytest.mark.parametrize("pos", [0, 17])
def test_iknp_2_24(pos):
    n = 1 << 24
    k0, k1, delta = drbg_labels("big/k0", 128), drbg_labels("big/k1", 128), drbg_labels("big/d", 1)
    db = [(int(delta["d0"][0]) >> i) & 1 if i < 64 else (int(delta["d1"][0]) >> (i - 64)) & 1 for i in range(128)]
    ks = np.where(np.array(db, dtype=bool), k1, k0).astype(LABEL_DTYPE)
    rcv, snd = IKNPReceiver(k0, k1), IKNPSender(ks, delta)
    rcv.pos = snd.pos = pos
    b = (np.random.default_rng(pos).integers(0, 2, n)).astype(np.uint8)
    u, t = rcv.receive(b)
    q = snd.send(u, n)
    assert rcv.pos == snd.pos == pos + stream_advance(n)
    # correlation over all 2^24 rows: t = q ^ b * Delta
    m = b.astype(bool)
    assert np.array_equal(t["d0"], q["d0"] ^ np.where(m, delta["d0"][0], 0).astype(np.uint64))
    assert np.array_equal(t["d1"], q["d1"] ^ np.where(m, delta["d1"][0], 0).astype(np.uint64))
    # sampled chunks against the oracle (the CTR stream is random access by byte position)
    for c0, rows in ((0, 1536), (12345, 1024), (32768 - 2, 1024)):
        lo = 512 * c0
        o_u, o_t, _ = O.iknp_receive(k0, k1, pos + 64 * c0, b[lo:lo + rows])
        assert eq(u[8192 * c0: 8192 * c0 + len(o_u)], o_u) and eq(t[lo:lo + rows], o_t)
        o_q, _ = O.iknp_send(ks, delta[0], pos + 64 * c0, o_u, rows)
        assert eq(q[lo:lo + rows], o_q)


def test_stream_program_over_1e8_gates():
    """BASELINE config 5 at its stated size (>= 10^8 gates per program instance, compiler/ssa/streamer.go:664-699,
    benchmarks.md:677-703) for a small batch: 377 chained sha512 / mul64 steps through the streaming garbler and
    evaluator; the first record streams equal the oracle's bytes, and the final state of the evaluator decodes to
    the plaintext evaluation of the whole program."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import stream_program as sp
    steps = sp.steps_for_gates(1e8)
    assert steps >= 300
    torch.cuda.set_device(0)
    res = sp.run_program(steps, batch=3, check=2, warm=2, oracle_steps=2)
    assert res["gates_per_instance"] >= 100_000_000
    assert res["checks_ok"]


def test_sha512_batch_that_overflows_the_resident_instances():
    """A batch larger than one wave of the all-hot sha512 plan (3 instances per SM) runs on the plan that keeps only a
    hot subset of the labels in shared memory (8 instances per SM), built on demand: sampled instances equal the oracle,
    every instance decodes to the plaintext function."""
    from mpc_b200.circuit import HostCircuit
    circ = load_circuit("sha512")
    eng = GarbleEngine(circ)
    batch, nin = 470, circ.num_inputs
    assert batch > eng.info.teams_per_sm * 148
    rng = np.random.default_rng(512)
    rand = rng.integers(0, 256, (batch, 16 * (1 + nin)), dtype=np.uint8)
    key = rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
    r, l0 = rand_to_labels(rand, nin)
    tables, io = eng.garble_batch(key, r, l0)
    bits = rng.integers(0, 2, (batch, nin), dtype=np.uint8)
    out = eng.eval_batch(key, tables, select(io[:, :nin], bits))
    assert np.array_equal(decode(io[:, nin:], out), HostCircuit(circ).compute_bits(bits))
    sample = [0, 147, 148, 469]
    _, o_tables, o_io = O.garble_batch(circ, key, rand[sample])
    assert eq(tables[sample], o_tables) and eq(io[sample], o_io)


def test_sha256_batch_that_overflows_the_resident_instances():
    """sha256 keeps 8 all-hot instances per SM; a batch beyond 8 x 148 runs on the plan with split live ranges (16
    instances per SM, labels evicted to / reloaded from the per-instance scratch by per-phase copy lists), built on
    demand: every instance decodes to SHA-256's compression function, sampled instances equal the oracle."""
    from mpc_b200.circuit import HostCircuit
    circ = load_circuit("sha256")
    eng = GarbleEngine(circ)
    batch, nin = 16 * 148 + 5, circ.num_inputs
    assert batch > eng.info.teams_per_sm * 148
    rng = np.random.default_rng(256)
    rand = rng.integers(0, 256, (batch, 16 * (1 + nin)), dtype=np.uint8)
    key = rng.integers(0, 256, 16, dtype=np.uint8).tobytes()
    r, l0 = rand_to_labels(rand, nin)
    tables, io = eng.garble_batch(key, r, l0)
    bits = rng.integers(0, 2, (batch, nin), dtype=np.uint8)
    out = eng.eval_batch(key, tables, select(io[:, :nin], bits))
    assert np.array_equal(decode(io[:, nin:], out), HostCircuit(circ).compute_bits(bits))
    sample = [0, 1183, 1184, 2367, batch - 1]
    _, o_tables, o_io = O.garble_batch(circ, key, rand[sample])
    assert eq(tables[sample], o_tables) and eq(io[sample], o_io)
