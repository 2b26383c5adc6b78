"""GPU parity of the asynchronous job form of the host entry points (gcb_garble_begin / gcb_eval_begin /
gcb_job_wait) and of the multi-device fan-out inside the library (gcb_set_devices, GCB_FLAG_FANOUT):
the bytes must equal the oracle's and the one-device call's.  Tests that need two devices skip below that."""
import numpy as np
import pytest
import torch

from conftest import load_circuit, mixed_circuit
from mpc_b200 import _lib
from mpc_b200 import circuit as gc
from mpc_b200.circuit import FLAG_FANOUT, GarbleEngine, host_alloc, host_free
from mpc_b200.circuit_io import LABEL_DTYPE, WIRE_DTYPE
from mpc_b200.ot import (IKNPReceiver, IKNPSender, cot_receive, cot_send, iknp_check_sums, mitccrh_hash_many,
                         rot_receive, rot_send)
from oracle import pyoracle as O
from util import drbg_labels, eq, garble_inputs, rand_to_labels, select

pytestmark = pytest.mark.gpu


def n_devices() -> int:
    return int(_lib.lib().gcb_device_count())


two_devices = pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2,
                                 reason="needs two CUDA devices")


@pytest.fixture
def all_devices():
    """Fan-out over every device of the box for the duration of one test."""
    gc.set_devices(list(range(n_devices())))
    yield n_devices()
    gc.set_devices([])


def _case(name, batch, klen, per_instance):
    circ = mixed_circuit(2, 500, 30, 12) if name == "mixed2" else load_circuit(name)
    eng = GarbleEngine(circ)
    keys, rand = garble_inputs(f"multi/{name}/{klen}", batch, circ.num_inputs, klen)
    if not per_instance:
        keys = keys[0].tobytes()
    r, l0 = rand_to_labels(rand, circ.num_inputs)
    return circ, eng, keys, rand, r, l0


@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("name,batch,klen,per_instance", [("aes_128", 23, 16, False), ("mixed2", 70, 32, True),
                                                          ("add64", 300, 24, False)])
def test_jobs_in_flight_match_oracle(name, batch, klen, per_instance, pinned):
    """Several garble and eval jobs in flight from one host thread; pinned (DMA in place) and pageable buffers."""
    circ, eng, keys, rand, r, l0 = _case(name, batch, klen, per_instance)
    nin, nout, rows = circ.num_inputs, circ.num_outputs, circ.num_rows
    alloc = host_alloc if pinned else (lambda shape, dt: np.zeros(shape, dtype=dt))
    h_r, h_l0 = alloc((batch,), LABEL_DTYPE), alloc((batch, nin), LABEL_DTYPE)
    h_tab, h_io = alloc((batch, max(rows, 1)), LABEL_DTYPE), alloc((batch, nin + nout), WIRE_DTYPE)
    h_in, h_out = alloc((batch, nin), LABEL_DTYPE), alloc((batch, nout), LABEL_DTYPE)
    h_r[:] = r
    h_l0[:] = l0
    tab = h_tab[:, :rows] if rows == h_tab.shape[1] else h_tab
    parts = [slice(k * batch // 4, (k + 1) * batch // 4) for k in range(4)]
    ks = (lambda sl: keys[sl]) if per_instance else (lambda sl: keys)
    jobs = [eng.garble_begin(ks(sl), h_r[sl], h_l0[sl], tab[sl], h_io[sl]) for sl in parts]
    assert all(isinstance(j.done(), bool) for j in jobs)
    _, o_tables, o_io = O.garble_batch(circ, keys, rand, threads=4)
    bits = np.random.default_rng(7).integers(0, 2, (batch, nin), dtype=np.uint8)
    ejobs = []
    for j, sl in zip(jobs, parts):                      # evaluate each part as soon as its tables are on the host
        j.wait()
        h_in[sl] = select(h_io[sl][:, :nin], bits[sl])
        ejobs.append(eng.eval_begin(ks(sl), tab[sl], h_in[sl], h_out[sl]))
    for j in ejobs:
        j.wait()
    assert eq(tab[:, :rows], o_tables) and eq(h_io, o_io)
    assert eq(h_out, O.eval_batch(circ, keys, o_tables, np.ascontiguousarray(h_in), threads=4))
    if pinned:
        for a in (h_r, h_l0, h_tab, h_io, h_in, h_out):
            host_free(a)


def test_job_error_paths():
    circ = load_circuit("add64")
    eng = GarbleEngine(circ)
    keys, rand = garble_inputs("multi/err", 4, circ.num_inputs, 16)
    r, l0 = rand_to_labels(rand, circ.num_inputs)
    tab = np.zeros((4, circ.num_rows), dtype=LABEL_DTYPE)
    with pytest.raises(_lib.GcbError, match="invalid key size"):
        eng.garble_begin(b"x" * 17, r, l0, tab)
    j = eng.garble_begin(keys[0].tobytes(), r, l0, tab)
    j.wait()
    j.wait()                                            # idempotent
    assert _lib.lib().gcb_job_wait(None) == _lib.E_ARG
    assert _lib.lib().gcb_job_done(None) == 1


@two_devices
@pytest.mark.parametrize("name,batch,klen,per_instance", [("aes_128", 41, 16, False), ("mixed2", 9, 32, True),
                                                          ("sha256", 5, 16, False), ("and", 1, 16, False)])
def test_fanout_garble_eval_matches_oracle(all_devices, name, batch, klen, per_instance):
    circ, eng, keys, rand, r, l0 = _case(name, batch, klen, per_instance)
    assert len(gc.get_devices()) == all_devices >= 2
    tables, io = eng.garble_batch(keys, r, l0)
    _, o_tables, o_io = O.garble_batch(circ, keys, rand, threads=4)
    assert eq(tables, o_tables) and eq(io, o_io)
    bits = np.random.default_rng(3).integers(0, 2, (batch, circ.num_inputs), dtype=np.uint8)
    inl = select(io[:, : circ.num_inputs], bits)
    out = eng.eval_batch(keys, tables, inl)
    assert eq(out, O.eval_batch(circ, keys, o_tables, inl, threads=4))
    gc.set_devices([])
    t1, io1 = eng.garble_batch(keys, r, l0)             # the one-device call
    assert eq(t1, tables) and eq(io1, io)


@two_devices
@pytest.mark.parametrize("n", [1, 511, 513, 5000, 70001])
def test_fanout_iknp_and_post_processing(all_devices, n):
    k0, k1, delta = drbg_labels("m/k0", 128), drbg_labels("m/k1", 128), drbg_labels("m/d", 1)
    db = [(int(delta["d0"][0]) >> i) & 1 if i < 64 else (int(delta["d1"][0]) >> (i - 64)) & 1 for i in range(128)]
    ks = np.where(np.array(db, dtype=bool), k1, k0)
    rcv, snd = IKNPReceiver(k0, k1), IKNPSender(ks, delta)
    rcv.pos = snd.pos = 17                              # unaligned stream position
    b = (np.arange(n) * 5 % 7 < 3).astype(np.uint8)
    u, t = rcv.receive(b)
    o_u, o_t, _ = O.iknp_receive(k0, k1, 17, b)
    assert eq(u, o_u) and eq(t, o_t)
    q = snd.send(u, n)
    o_q, _ = O.iknp_send(ks, delta[0], 17, o_u, n)
    assert eq(q, o_q)
    # post-processing on OT ranges: each device hashes its block under MiTCCRH keys lo..hi
    seed = drbg_labels("m/seed", 1)
    wires = drbg_labels("m/w", 2 * n).view(WIRE_DTYPE).reshape(n) if n else np.zeros(0, WIRE_DTYPE)
    msgs = cot_send(seed, delta, q, wires)
    got = cot_receive(seed, b, msgs, t)
    assert eq(got, np.where(b.astype(bool), wires["l1"], wires["l0"]).astype(LABEL_DTYPE))
    rw, rr = rot_send(seed, delta, q), rot_receive(seed, t)
    assert eq(rr, np.where(b.astype(bool), rw["l1"], rw["l0"]).astype(LABEL_DTYPE))
    sums = iknp_check_sums(seed, 5, t, b)
    blks = t.copy()
    mitccrh_hash_many(seed, 3, blks, n, 1)
    gc.set_devices([0])
    assert eq(msgs, cot_send(seed, delta, q, wires)) and eq(rw, rot_send(seed, delta, q))
    assert eq(np.asarray(sums), np.asarray(iknp_check_sums(seed, 5, t, b)))
    b1 = t.copy()
    mitccrh_hash_many(seed, 3, b1, n, 1)
    assert eq(blks, b1)


@two_devices
def test_fanout_bit_cot(all_devices):
    n = 3000
    k0, k1, delta = drbg_labels("mb/k0", 128), drbg_labels("mb/k1", 128), drbg_labels("mb/d", 1)
    db = [(int(delta["d0"][0]) >> i) & 1 if i < 64 else (int(delta["d1"][0]) >> (i - 64)) & 1 for i in range(128)]
    ks = np.where(np.array(db, dtype=bool), k1, k0)
    words = np.random.default_rng(5).integers(0, 2**63, (n + 63) // 64, dtype=np.uint64)
    rcv, snd = IKNPReceiver(k0, k1), IKNPSender(ks, delta)
    u, res = rcv.receive_bits(words, n)
    sres = snd.send_bits(u, n)
    gc.set_devices([0])
    rcv1, snd1 = IKNPReceiver(k0, k1), IKNPSender(ks, delta)
    u1, res1 = rcv1.receive_bits(words, n)
    assert eq(u, u1) and eq(res, res1) and eq(sres, snd1.send_bits(u1, n))


@two_devices
@pytest.mark.parametrize("name,batch,klen,per_instance", [("aes_128", 37, 16, False), ("mixed2", 11, 32, True)])
def test_dev_fanout_scatter_gather_over_peers(all_devices, name, batch, klen, per_instance):
    """GCB_FLAG_FANOUT: operands on device 0, the batch split over all devices, tables gathered back by peer copies."""
    circ, eng, keys, rand, r, l0 = _case(name, batch, klen, per_instance)
    nin, nout, rows = circ.num_inputs, circ.num_outputs, circ.num_rows
    dev = torch.device("cuda", 0)

    def to_dev(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)

    karr = np.frombuffer(keys, dtype=np.uint8) if isinstance(keys, bytes) else keys
    d_key, d_r, d_l0 = to_dev(karr), to_dev(r), to_dev(l0)
    d_tab = torch.zeros(batch * rows * 16, dtype=torch.uint8, device=dev)
    d_io = torch.zeros(batch * (nin + nout) * 32, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.lib().gcb_set_device(0))
    try:
        eng.garble_dev(d_key, klen, klen if per_instance else 0, batch, d_r, d_l0, d_tab, d_io, stream=s, flags=FLAG_FANOUT)
        torch.cuda.synchronize(dev)
        _, o_tables, o_io = O.garble_batch(circ, keys, rand, threads=4)
        assert d_tab.cpu().numpy().tobytes() == o_tables.tobytes()
        assert d_io.cpu().numpy().tobytes() == o_io.tobytes()
        bits = np.random.default_rng(9).integers(0, 2, (batch, nin), dtype=np.uint8)
        inl = select(o_io[:, :nin], bits)
        d_in = to_dev(inl)
        d_out = torch.zeros(batch * nout * 16, dtype=torch.uint8, device=dev)
        eng.eval_dev(d_key, klen, klen if per_instance else 0, batch, d_tab, d_in, d_out, stream=s, flags=FLAG_FANOUT)
        torch.cuda.synchronize(dev)
        assert d_out.cpu().numpy().tobytes() == O.eval_batch(circ, keys, o_tables, inl, threads=4).tobytes()
    finally:
        _lib.check(_lib.lib().gcb_set_device(-1))
