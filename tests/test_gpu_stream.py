"""GPU parity for the streaming garbler: the emitted record stream (headers +
rows), byte for byte, and the permanent wires, against the oracle's emitter."""
import numpy as np
import pytest

from conftest import load_circuit, mixed_circuit
from mpc_b200.circuit import GarbleEngine, Streaming
from mpc_b200.circuit_io import LABEL_DTYPE
from oracle import pyoracle as O
from util import DRBG, eq

pytestmark = pytest.mark.gpu


def _wire_tuple(w):
    return ((int(w["l0"]["d0"]), int(w["l0"]["d1"])), (int(w["l1"]["d0"]), int(w["l1"]["d1"])))


def _run(circ, key, in_ids, out_ids, input_ids=None, tag="s"):
    input_ids = list(in_ids) if input_ids is None else input_ids
    uniq = list(dict.fromkeys(input_ids))
    rand = DRBG(tag).read(16 * (1 + len(uniq)))
    st = Streaming.new(rand, key, uniq)
    ost = O.Streaming(key, rand, uniq)
    eng = GarbleEngine(circ)
    buf, t0, t1 = st.garble(eng, in_ids, out_ids)
    want = ost.garble(circ, in_ids, out_ids)
    assert buf.shape == (1, len(want))
    assert buf[0].tobytes() == want, "record stream differs from the oracle"
    ids = list(dict.fromkeys(list(in_ids) + list(out_ids)))
    got = st.get_inputs(ids)[0]
    for k, wid in enumerate(ids):
        assert _wire_tuple(got[k]) == ost.get_input(wid), f"permanent wire {wid} differs"
    assert t1 > 0
    return st, ost, eng


@pytest.mark.parametrize("klen", [16, 32])
def test_sha256_one_step_bit_exact(klen):
    """BASELINE config 3: sha256.circ as one Streaming.Garble step, in = 0..767, fresh out ids."""
    circ = load_circuit("sha256")
    key = b"\0" * 16 if klen == 16 else DRBG("s256key").read(32)      # stream_garble_test.go:47 uses a zero key
    _run(circ, key, list(range(768)), list(range(768, 1024)), tag=f"sha256/{klen}")


def test_long_and_short_index_records():
    circ = mixed_circuit(7, 600, 24, 8)
    key = DRBG("k").read(24)
    # ids above 0xffff force the 32-bit record form for gates that touch them
    in_ids = [70000 + 3 * i for i in range(12)] + list(range(5, 17))
    out_ids = [65535, 65536, 9, 300000, 1, 2, 3, 4]
    _run(circ, key, in_ids, out_ids, tag="longshort")


def test_chained_steps_and_aliased_ids():
    """Steps chained through permanent wires; an output id that is also an input id, a
    repeated input id and a repeated output id follow the reference's sequential semantics."""
    circ = load_circuit("add64")
    key = DRBG("chain").read(16)
    ins = list(range(100, 228))
    rand = DRBG("chain/r").read(16 * (1 + 128))
    st, ost, eng = Streaming.new(rand, key, ins), O.Streaming(key, rand, ins), GarbleEngine(circ)
    steps = [
        (ins, list(range(300, 364))),
        (list(range(300, 364)) + list(range(100, 164)), list(range(400, 464))),       # consume a result
        (list(range(400, 464)) + list(range(400, 464)), list(range(400, 464))),       # x = x + x in place
        (list(range(400, 464)) + list(range(164, 228)), [500] * 32 + list(range(501, 533))),   # repeated out id
    ]
    for i, o in steps:
        buf, _, _ = st.garble(eng, i, o)
        assert buf[0].tobytes() == ost.garble(circ, i, o)
        ids = list(dict.fromkeys(i + o))
        got = st.get_inputs(ids)[0]
        for k, wid in enumerate(ids):
            assert _wire_tuple(got[k]) == ost.get_input(wid)


def test_streaming_batch_matches_single_instances():
    circ = load_circuit("mul64")
    batch, ids = 5, list(range(128))
    keys = np.stack([DRBG(f"sb/key/{b}").array(32) for b in range(batch)])
    rands = [DRBG(f"sb/{b}").read(16 * 129) for b in range(batch)]
    lab = np.stack([np.frombuffer(r, dtype=">u8").astype("<u8").view(LABEL_DTYPE) for r in rands])
    st = Streaming(keys, np.ascontiguousarray(lab[:, 0]), ids, np.ascontiguousarray(lab[:, 1:]))
    eng = GarbleEngine(circ)
    buf, _, _ = st.garble(eng, ids, list(range(1000, 1064)))
    for b in range(batch):
        ost = O.Streaming(keys[b].tobytes(), rands[b], ids)
        assert buf[b].tobytes() == ost.garble(circ, ids, list(range(1000, 1064))), f"instance {b}"


# ------------------------------------------------------------------ streaming evaluator
from mpc_b200.circuit import StreamEval  # noqa: E402


def _lab_tuple(l):
    return (int(l["d0"]), int(l["d1"]))


def _eval_step(sev, osev, buf, circ, ins, outs, nwires):
    """Feed one sub-circuit's record stream to the device evaluator and to the oracle's."""
    used = sev.circuit(buf, circ.num_gates, circ.num_wires, nwires)
    assert used == buf.shape[1]
    o_used = osev.circuit(buf[0].tobytes(), circ.num_gates, circ.num_wires, nwires)
    assert o_used == used
    ids = list(dict.fromkeys(list(ins) + list(outs)))
    got = sev.get(ids)[0]
    for k, wid in enumerate(ids):
        assert _lab_tuple(got[k]) == osev.get(wid), f"evaluator wire {wid} differs from the oracle"


@pytest.mark.parametrize("name,klen", [("sha256", 16), ("aes_128", 32), ("mul64", 24)])
def test_stream_eval_one_step_bit_exact(name, klen):
    circ = load_circuit(name)
    key = DRBG(f"se/{name}").read(klen)
    nin, nout = circ.num_inputs, circ.num_outputs
    ins, outs = list(range(nin)), list(range(nin + 5, nin + 5 + nout))
    rand = DRBG(f"se/r/{name}").read(16 * (1 + nin))
    st, eng = Streaming.new(rand, key, ins), GarbleEngine(circ)
    buf, _, _ = st.garble(eng, ins, outs)
    wires = st.get_inputs(ins)[0]
    bits = np.random.default_rng(3).integers(0, 2, nin).astype(bool)
    labels = np.where(bits, wires["l1"], wires["l0"]).astype(LABEL_DTYPE)
    sev, osev = StreamEval(key), O.StreamEval(key)
    sev.set(ins, labels.reshape(1, -1))
    for i, wid in enumerate(ins):
        osev.set(wid, _lab_tuple(labels[i]))
    _eval_step(sev, osev, buf, circ, ins, outs, nin + 5 + nout)
    # and the evaluated labels decode to the plaintext circuit
    ow = st.get_inputs(outs)[0]
    got = sev.get(outs)[0]
    dec = np.where(got == ow["l1"], 1, np.where(got == ow["l0"], 0, 2))
    assert np.array_equal(dec, circ.compute_bits(bits.astype(int).tolist()))


def test_stream_eval_chained_aliased_steps_and_long_indices():
    circ = load_circuit("add64")
    key = DRBG("sechain").read(16)
    ins = list(range(70000, 70128))                   # 32-bit record form
    rand = DRBG("sechain/r").read(16 * 129)
    st, eng = Streaming.new(rand, key, ins), GarbleEngine(circ)
    wires = st.get_inputs(ins)[0]
    bits = np.random.default_rng(5).integers(0, 2, 128).astype(bool)
    labels = np.where(bits, wires["l1"], wires["l0"]).astype(LABEL_DTYPE)
    sev, osev = StreamEval(key), O.StreamEval(key)
    sev.set(ins, labels.reshape(1, -1))
    for i, wid in enumerate(ins):
        osev.set(wid, _lab_tuple(labels[i]))
    steps = [
        (ins, list(range(300, 364))),
        (list(range(300, 364)) + ins[:64], list(range(400, 464))),
        (list(range(400, 464)) + list(range(400, 464)), list(range(400, 464))),           # in place
        (list(range(400, 464)) + ins[64:], [500] * 32 + list(range(501, 533))),            # repeated out id
    ]
    for i, o in steps:
        buf, _, _ = st.garble(eng, i, o)
        _eval_step(sev, osev, buf, circ, i, o, 70200)


def test_stream_eval_batch_and_malformed_stream():
    circ = load_circuit("sub64")
    batch, ids = 4, list(range(128))
    keys = np.stack([DRBG(f"seb/key/{b}").array(32) for b in range(batch)])
    rands = [DRBG(f"seb/{b}").read(16 * 129) for b in range(batch)]
    lab = np.stack([np.frombuffer(r, dtype=">u8").astype("<u8").view(LABEL_DTYPE) for r in rands])
    st = Streaming(keys, np.ascontiguousarray(lab[:, 0]), ids, np.ascontiguousarray(lab[:, 1:]))
    eng = GarbleEngine(circ)
    outs = list(range(200, 264))
    buf, _, _ = st.garble(eng, ids, outs)
    wires = st.get_inputs(ids)
    bits = np.random.default_rng(9).integers(0, 2, (batch, 128)).astype(bool)
    labels = np.where(bits, wires["l1"], wires["l0"]).astype(LABEL_DTYPE)
    sev = StreamEval(keys, batch)
    sev.set(ids, labels)
    assert sev.circuit(buf, circ.num_gates, circ.num_wires, 264) == buf.shape[1]
    got = sev.get(outs)
    for b in range(batch):
        osev = O.StreamEval(keys[b].tobytes())
        for i, wid in enumerate(ids):
            osev.set(wid, _lab_tuple(labels[b, i]))
        osev.circuit(buf[b].tobytes(), circ.num_gates, circ.num_wires, 264)
        for k, wid in enumerate(outs):
            assert _lab_tuple(got[b, k]) == osev.get(wid)
    from mpc_b200 import _lib
    with pytest.raises(_lib.GcbError) as e:
        sev.circuit(buf[:, : buf.shape[1] // 2], circ.num_gates, circ.num_wires, 264)
    assert e.value.rc == _lib.E_BUFFER
    bad = buf.copy()
    bad[:, 0] = 0x0f
    with pytest.raises(_lib.GcbError) as e:
        sev.circuit(bad, circ.num_gates, circ.num_wires, 264)
    assert e.value.rc == _lib.E_BADOP


@pytest.mark.parametrize("variant", [dict(GCB_NT="4", GCB_ILP="2"), dict(GCB_NT="4", GCB_ILP="1"),
                                     dict(GCB_NT="2", GCB_ILP="1"), dict(GCB_NT="2", GCB_ILP="2")])
def test_streaming_under_every_kernel_variant(variant, monkeypatch):
    """The streaming garbler and evaluator run the same gate kernels in wire-file mode: every variant
    (resident T-tables x AES blocks per thread) emits the oracle's bytes and evaluates them back."""
    from mpc_b200.circuit import StreamEval
    for k, v in variant.items():
        monkeypatch.setenv(k, v)
    circ = mixed_circuit(11, 1500, 40, 16)
    key = DRBG("variant/stream").read(32)
    in_ids, out_ids = list(range(40)), list(range(100, 116))
    st, ost, eng = _run(circ, key, in_ids, out_ids, tag="variant/stream")
    buf, _, _ = st.garble(eng, in_ids, out_ids)               # second step on the same ids: same bytes again
    bits = np.random.default_rng(2).integers(0, 2, 40).astype(bool)
    w = st.get_inputs(in_ids)[0]
    sev = StreamEval(key, 1)
    sev.set(in_ids, np.where(bits, w["l1"], w["l0"]).astype(LABEL_DTYPE).reshape(1, -1))
    sev.circuit(buf, circ.num_gates, circ.num_wires, 116)
    got = sev.get(out_ids)[0]
    ow = st.get_inputs(out_ids)[0]
    dec = np.where(got == ow["l1"], 1, np.where(got == ow["l0"], 0, 2))
    assert np.array_equal(dec, circ.compute_bits(bits.astype(np.uint8).tolist()))


def test_exact_size_buffer_and_too_small_buffer():
    """gcb_stream_garble with a buffer of exactly step_size bytes (the layout is built before the launch)
    gives the same bytes as the over-sized buffer path (layout built while the kernel runs); a buffer that is
    too small is refused before anything is garbled."""
    import ctypes as C
    from mpc_b200 import _lib
    from mpc_b200._lib import ptr
    circ = mixed_circuit(13, 900, 32, 12)
    key = DRBG("exact/key").read(16)
    in_ids, out_ids = list(range(32)), list(range(50, 62))
    rand = DRBG("exact/rand").read(16 * 33)
    eng = GarbleEngine(circ)
    want = O.Streaming(key, rand, in_ids).garble(circ, in_ids, out_ids)
    st = Streaming.new(rand, key, in_ids)
    n = st.step_size(eng, in_ids, out_ids)
    assert n == len(want)
    i32, o32 = np.array(in_ids, np.uint32), np.array(out_ids, np.uint32)
    small = np.zeros((1, n - 1), np.uint8)
    w, t0, t1 = C.c_size_t(), C.c_uint64(), C.c_uint64()
    rc = _lib.lib().gcb_stream_garble(st._h, eng.handle, ptr(i32), 32, ptr(o32), 12, ptr(small), n - 1,
                                      C.byref(w), C.byref(t0), C.byref(t1))
    assert rc == _lib.E_BUFFER and int(w.value) == n
    exact = np.zeros((1, n), np.uint8)
    _lib.check(_lib.lib().gcb_stream_garble(st._h, eng.handle, ptr(i32), 32, ptr(o32), 12, ptr(exact), n,
                                            C.byref(w), C.byref(t0), C.byref(t1)))
    assert exact[0].tobytes() == want
    st2 = Streaming.new(rand, key, in_ids)
    buf, _, _ = st2.garble(eng, in_ids, out_ids)
    assert buf[0].tobytes() == want


def test_garble_begin_wait_keeps_a_step_in_flight():
    """gcb_stream_garble_begin / _wait: chained steps issued back to back (the second begun before the first
    one's bytes were waited for) give the oracle's bytes and wire file, exactly like the blocking call."""
    circ = load_circuit("add64")
    key = DRBG("async").read(16)
    ins = list(range(100, 228))
    rand = DRBG("async/r").read(16 * (1 + 128))
    st, ost, eng = Streaming.new(rand, key, ins), O.Streaming(key, rand, ins), GarbleEngine(circ)
    st.stream_buffers = 3
    steps = [
        (ins, list(range(300, 364))),
        (list(range(300, 364)) + list(range(100, 164)), list(range(400, 464))),
        (list(range(400, 464)) + list(range(164, 228)), list(range(500, 564))),
        (list(range(500, 564)) + list(range(300, 364)), list(range(600, 664))),
    ]
    bufs = []
    for k, (i, o) in enumerate(steps):
        bufs.append(st.garble_begin(eng, i, o))
        st.garble_wait(1)                                   # everything but the newest step is in place
        if k:
            assert bufs[k - 1][0].tobytes() == ost.garble(circ, *steps[k - 1])
    st.garble_wait(0)
    assert bufs[-1][0].tobytes() == ost.garble(circ, *steps[-1])
    ids = list(range(600, 664))
    got = st.get_inputs(ids)[0]
    for k, wid in enumerate(ids):
        assert _wire_tuple(got[k]) == ost.get_input(wid)


def test_streaming_step_that_spills_labels():
    """A sub-circuit with more live labels than fit on chip (9,000 inputs kept live through 12,000 gates of all
    types) through the streaming garbler and evaluator: the spilling kernel variant in wire-file mode."""
    from mpc_b200.circuit import StreamEval
    circ = mixed_circuit(31, 12000, 9000, 32)
    key = DRBG("spill/stream").read(16)
    in_ids, out_ids = list(range(9000)), list(range(20000, 20032))
    st, ost, eng = _run(circ, key, in_ids, out_ids, tag="spill/stream")
    assert eng.info.num_slots > 6400
    buf, _, _ = st.garble(eng, in_ids, out_ids)
    bits = np.random.default_rng(6).integers(0, 2, 9000).astype(bool)
    w = st.get_inputs(in_ids)[0]
    sev = StreamEval(key, 1)
    sev.set(in_ids, np.where(bits, w["l1"], w["l0"]).astype(LABEL_DTYPE).reshape(1, -1))
    sev.circuit(buf, circ.num_gates, circ.num_wires, 20032)
    got = sev.get(out_ids)[0]
    ow = st.get_inputs(out_ids)[0]
    dec = np.where(got == ow["l1"], 1, np.where(got == ow["l0"], 0, 2))
    assert np.array_equal(dec, circ.compute_bits(bits.astype(np.uint8).tolist()))


def test_stream_eval_refuses_wire_counts_it_cannot_hold():
    """The wire count of an OpCircuit comes from the peer (the reference sizes a slice by it): counts beyond 2^28, or
    beyond the free device memory, are refused at once -- not after allocating page by page towards them."""
    import time
    from mpc_b200 import _lib
    circ = load_circuit("sub64")
    key = DRBG("big/key").read(16)
    sev = StreamEval(key, 1)
    stream = np.zeros((1, 64), dtype=np.uint8)
    t0 = time.time()
    for nwires in (0xffffffff, (1 << 28) + 1):
        with pytest.raises(_lib.GcbError) as e:
            sev.circuit(stream, 1, circ.num_wires, nwires)
        assert e.value.rc in (_lib.E_TOO_LARGE, _lib.E_BUFFER, _lib.E_BADOP, _lib.E_WIRE)
    assert time.time() - t0 < 5
