"""GPU parity: batched garble / eval through the C ABI against the CPU oracle,
bit-exact (integer / byte work), on the same seeded inputs.  Also the
reference's single-instance API (Garble / Eval mirrors) and its error paths."""
import numpy as np
import pytest
import torch

from conftest import load_circuit, millionaire_circuit, mixed_circuit
from mpc_b200 import _lib
from mpc_b200.circuit import Garbled, GarbleEngine, decode_bits_dev, select_labels_dev
from mpc_b200.circuit_io import LABEL_DTYPE, WIRE_DTYPE
from oracle import pyoracle as O
from util import DRBG, decode, eq, garble_inputs, rand_to_labels, select

pytestmark = pytest.mark.gpu

_engines = {}


def get(name):
    if name not in _engines:
        circ = {"mixed1": lambda: mixed_circuit(1), "mixed2": lambda: mixed_circuit(2, 500, 30, 12),
                "mixed3": lambda: mixed_circuit(5, 3000, 64, 32),
                "millionaire": millionaire_circuit}.get(name, lambda: load_circuit(name))()
        _engines[name] = (circ, GarbleEngine(circ))
    return _engines[name]


CASES = [("and", 70, 16), ("not", 33, 32), ("add64", 40, 16), ("sub64", 17, 24), ("mixed1", 37, 16),
         ("mixed2", 21, 32), ("mixed3", 9, 24), ("millionaire", 12, 32), ("mul64", 5, 16), ("div64", 3, 32),
         ("aes_128", 13, 16), ("aes_128", 7, 32), ("aes_256", 3, 24), ("sha256", 3, 16), ("sha256xor", 2, 32),
         ("chacha20block", 2, 16), ("sha512", 2, 32)]


@pytest.mark.parametrize("per_instance_keys", [False, True])
@pytest.mark.parametrize("name,batch,klen", CASES)
def test_garble_eval_batch_bit_exact(name, batch, klen, per_instance_keys, monkeypatch):
    # per-instance keys: the library's own launch shape (a small batch is spread over the SMs, one team per CTA);
    # shared key: full CTAs, every team of an SM racing for the instances
    monkeypatch.setenv("GCB_SPREAD", "1" if per_instance_keys else "0")
    circ, eng = get(name)
    keys, rand = garble_inputs(f"gpu/{name}/{klen}", batch, circ.num_inputs, klen)
    if not per_instance_keys:
        keys = keys[0].tobytes()
    r, l0 = rand_to_labels(rand, circ.num_inputs)
    tables, io = eng.garble_batch(keys, r, l0)
    o_r, o_tables, o_io = O.garble_batch(circ, keys, rand, threads=4)
    assert eq(tables, o_tables), "garbled tables differ from the oracle"
    assert eq(io, o_io), "input / output wires differ from the oracle"
    # evaluate on random input bits
    rng = np.random.default_rng(batch * 1000 + klen)
    bits = rng.integers(0, 2, (batch, circ.num_inputs), dtype=np.uint8)
    inl = select(io[:, : circ.num_inputs], bits)
    out = eng.eval_batch(keys, tables, inl)
    o_out = O.eval_batch(circ, keys, o_tables, inl, threads=4)
    assert eq(out, o_out), "output labels differ from the oracle"
    got = decode(io[:, circ.num_inputs:], out)
    for i in range(min(batch, 3)):
        assert np.array_equal(got[i], circ.compute_bits(bits[i].tolist()))


@pytest.mark.parametrize("name,klen", [("and", 16), ("not", 32), ("mixed1", 24), ("add64", 32), ("sha256xor", 32)])
def test_reference_api_single_instance(name, klen):
    """(*Circuit).Garble(rand, key) / (*Circuit).Eval(key, wires, garbled) mirrors: every wire and row."""
    circ, eng = get(name)
    key = DRBG(f"key/{name}").read(klen)
    rand = DRBG(f"rand/{name}").read(16 * (1 + circ.num_inputs))
    g = eng.garble(rand, key)
    o_r, o_wires, o_slab, o_off = O.garble(circ, key, rand)
    assert g.R == o_r and int(g.R["d0"]) >> 63 == 1
    assert eq(g.Wires, o_wires) and eq(g.slab, o_slab)
    gates = g.Gates
    for i, op in enumerate(circ.gates["op"].tolist()):
        n = {0: 0, 1: 0, 2: 2, 3: 3, 4: 1}[op]
        assert (gates[i] is None) == (n == 0) and (n == 0 or len(gates[i]) == n)
    bits = np.random.default_rng(1).integers(0, 2, circ.num_inputs)
    wires = np.zeros(circ.num_wires, dtype=LABEL_DTYPE)
    wires[: circ.num_inputs] = np.where(bits.astype(bool), g.Wires["l1"][: circ.num_inputs], g.Wires["l0"][: circ.num_inputs])
    o_w = O.eval_(circ, key, wires[: circ.num_inputs].copy(), o_slab, o_off)
    eng.eval(key, wires, gates)
    assert eq(wires, o_w)


def test_eval_error_paths_match_reference():
    circ, eng = get("mixed1")
    key = b"k" * 16
    g = eng.garble(DRBG("e").read(16 * (1 + circ.num_inputs)), key)
    wires = np.zeros(circ.num_wires, dtype=LABEL_DTYPE)
    gates = list(g.Gates)
    and_i = int(np.nonzero(circ.gates["op"] == 2)[0][0])
    gates[and_i] = gates[and_i][:1]
    with pytest.raises(_lib.GcbError, match="AND row length"):
        eng.eval(key, wires, gates)
    gates = list(g.Gates)
    inv_i = int(np.nonzero(circ.gates["op"] == 4)[0][0])
    gates[inv_i] = None
    with pytest.raises(_lib.GcbError, match="index 0 >= row 0"):
        eng.eval(key, wires, gates)
    with pytest.raises(_lib.GcbError, match="invalid key size"):
        eng.eval(b"x" * 17, wires, g)


def test_empty_batch_and_ragged_last_wave():
    circ, eng = get("add64")
    t, io = eng.garble_batch(b"\0" * 16, np.zeros(0, LABEL_DTYPE), np.zeros((0, circ.num_inputs), LABEL_DTYPE))
    assert t.shape == (0, circ.num_rows) and io.shape[0] == 0
    # a batch that is not a multiple of the resident team count, larger than one wave
    batch = 16 * 148 + 5
    keys, rand = garble_inputs("ragged", batch, circ.num_inputs, 16)
    r, l0 = rand_to_labels(rand, circ.num_inputs)
    tables, io = eng.garble_batch(keys, r, l0)
    _, o_tables, o_io = O.garble_batch(circ, keys, rand, threads=8)
    assert eq(tables, o_tables) and eq(io, o_io)


def test_device_resident_path_with_label_plumbing():
    """gcb_garble_dev -> gcb_select_labels_dev -> gcb_eval_dev -> gcb_decode_bits_dev on torch buffers."""
    circ, eng = get("aes_128")
    batch, nin, nout = 64, circ.num_inputs, circ.num_outputs
    _, rand = garble_inputs("dev", batch, nin, 0)
    key = b"0123456789abcdef"                       # circuit/garble_bench_test.go:34
    r, l0 = rand_to_labels(rand, nin)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a.view(np.uint8).reshape(a.shape + (-1,))).to(dev)
    d_key, d_r, d_l0 = torch.frombuffer(bytearray(key), dtype=torch.uint8).to(dev), t(r), t(l0)
    d_tab = torch.zeros((batch, circ.num_rows, 16), dtype=torch.uint8, device=dev)
    d_io = torch.zeros((batch, nin + nout, 32), dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    eng.garble_dev(d_key, 16, 0, batch, d_r, d_l0, d_tab, d_io, stream=s)
    # plaintext: AES key 000102..0f, block = instance index (big-endian integer, wire i = bit i)
    pt_key = int.from_bytes(bytes(range(16)), "big")
    bits = np.zeros((batch, nin), dtype=np.uint8)
    for i in range(batch):
        bits[i, :128] = [(pt_key >> b) & 1 for b in range(128)]
        bits[i, 128:] = [(i >> b) & 1 for b in range(128)]
    d_bits = torch.from_numpy(bits).to(dev)
    d_in = torch.zeros((batch, nin, 16), dtype=torch.uint8, device=dev)
    select_labels_dev(d_io, nin + nout, d_bits, d_in, batch, nin, stream=s)
    d_out = torch.zeros((batch, nout, 16), dtype=torch.uint8, device=dev)
    eng.eval_dev(d_key, 16, 0, batch, d_tab, d_in, d_out, stream=s)
    d_obits = torch.zeros((batch, nout), dtype=torch.uint8, device=dev)
    decode_bits_dev(d_io[:, nin:].contiguous(), nout, d_out, d_obits, batch, nout, stream=s)
    torch.cuda.synchronize()
    ob = d_obits.cpu().numpy()
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
    enc = Cipher(algorithms.AES(bytes(range(16))), modes.ECB()).encryptor()
    for i in range(batch):
        want = int.from_bytes(enc.update(i.to_bytes(16, "big")), "big")
        assert sum(int(b) << k for k, b in enumerate(ob[i])) == want
    # and the tables equal the oracle's
    _, o_tables, _ = O.garble_batch(circ, key, rand, threads=4)
    assert d_tab.cpu().numpy().tobytes() == o_tables.tobytes()


def test_fips197_kat_through_garble_eval():
    """aes_128.circ(key=000102..0f, pt=00112233..ff) = 69c4e0d8...c55a (FIPS-197 C.1) through GPU garble+eval."""
    circ, eng = get("aes_128")
    key = DRBG("kat").read(32)
    g = eng.garble(DRBG("kat/r").read(16 * 257), key)
    k = int.from_bytes(bytes(range(16)), "big")
    p = int.from_bytes(bytes.fromhex("00112233445566778899aabbccddeeff"), "big")
    bits = [(k >> i) & 1 for i in range(128)] + [(p >> i) & 1 for i in range(128)]
    wires = np.zeros(circ.num_wires, dtype=LABEL_DTYPE)
    wires[:256] = np.where(np.array(bits, dtype=bool), g.Wires["l1"][:256], g.Wires["l0"][:256])
    eng.eval(key, wires, g)
    ob = decode(g.Wires[-128:], wires[-128:])
    assert sum(int(b) << i for i, b in enumerate(ob)) == 0x69c4e0d86a7b0430d8cdb78070b4c55a


def _wire_ref(circ, tables_one):
    """numpy restatement of circuit/garbler.go:69-82 for one instance."""
    out = bytearray(int(circ.num_gates).to_bytes(4, "big"))
    off = circ.row_offsets()
    for g in range(circ.num_gates):
        rows = tables_one[off[g]: off[g + 1]]
        out += len(rows).to_bytes(4, "big")
        for r in rows:
            out += int(r["d0"]).to_bytes(8, "big") + int(r["d1"]).to_bytes(8, "big")
    return bytes(out)


@pytest.mark.parametrize("name", ["add64", "aes_128", "mixed"])
def test_tables_wire_format_round_trip(name):
    """Garbler's table stream (garbler.go:69-82) and Evaluator's parse (evaluator.go:40-66) on the device."""
    from conftest import mixed_circuit
    circ = mixed_circuit(3) if name == "mixed" else load_circuit(name)
    eng = GarbleEngine(circ)
    batch = 5
    keys, rand = garble_inputs(f"wire/{name}", batch, circ.num_inputs, 16)
    r, l0 = rand_to_labels(rand, circ.num_inputs)
    tables, _ = eng.garble_batch(keys, r, l0)
    wire = eng.tables_to_wire(tables)
    assert wire.shape == (batch, 4 + 4 * circ.num_gates + 16 * circ.num_rows)
    for b in (0, batch - 1):
        assert wire[b].tobytes() == _wire_ref(circ, tables[b]), "wire bytes differ from the reference layout"
    back = eng.tables_from_wire(wire)
    assert eq(back, tables)
    bad = wire.copy()
    bad[1, 3] ^= 1                                       # NumGates
    with pytest.raises(_lib.GcbError, match="wrong number of gates"):
        eng.tables_from_wire(bad)
    if circ.num_rows:
        bad = wire.copy()
        first = next(g for g in range(circ.num_gates) if circ.row_offsets()[g + 1] > circ.row_offsets()[g])
        pos = 4 + 4 * first + 16 * int(circ.row_offsets()[first])
        bad[2, pos + 3] ^= 1                             # row count of the first ciphered gate
        with pytest.raises(_lib.GcbError, match="corrupted circuit") as e:
            eng.tables_from_wire(bad)
        assert e.value.rc == _lib.E_CORRUPT


@pytest.mark.parametrize("variant", [dict(GCB_NT="4", GCB_ILP="2"), dict(GCB_NT="4", GCB_ILP="1"),
                                     dict(GCB_NT="2", GCB_ILP="1"), dict(GCB_NT="2", GCB_ILP="2"),
                                     dict(GCB_NT="2", GCB_ILP="1", GCB_TEAM_THREADS="32", GCB_TEAMS="3"),
                                     dict(GCB_NT="4", GCB_ILP="2", GCB_TEAM_THREADS="64", GCB_STAGGER="0")])
def test_every_kernel_variant_is_bit_exact(variant, monkeypatch):
    """The geometry (resident T-tables, AES blocks per thread, team width, node rows in flight) must not change
    a bit: every kernel variant, plain and full-wire mode, against the oracle on a circuit with all gate types
    and on mul64 (deep carry chains: many waves per phase)."""
    monkeypatch.setenv("GCB_SPREAD", "0")                # full CTAs (every team of an SM) although the batch is small
    for k, v in variant.items():
        monkeypatch.setenv(k, v)
    for circ, batch in ((mixed_circuit(7, 2500, 48, 24), 7), (load_circuit("mul64"), 3)):
        eng = GarbleEngine(circ)                                   # geometry is read when the plan is built
        keys, rand = garble_inputs(f"variant/{circ.name}", batch, circ.num_inputs, 24)
        r, l0 = rand_to_labels(rand, circ.num_inputs)
        tables, io = eng.garble_batch(keys, r, l0)
        _, o_tables, o_io = O.garble_batch(circ, keys, rand)
        assert eq(tables, o_tables) and eq(io, o_io)
        bits = np.random.default_rng(1).integers(0, 2, (batch, circ.num_inputs), dtype=np.uint8)
        inl = select(io[:, : circ.num_inputs], bits)
        out = eng.eval_batch(keys, tables, inl)
        assert eq(out, O.eval_batch(circ, keys, o_tables, inl))
        # full-wire mode (Garbled.Wires / Eval's in-place wires) through the single-instance mirror
        key = DRBG(f"variant/key/{circ.name}").read(16)
        rb = DRBG(f"variant/rand/{circ.name}").read(16 * (1 + circ.num_inputs))
        g = eng.garble(rb, key)
        _, o_wires, o_slab, _ = O.garble(circ, key, rb)
        assert eq(g.Wires, o_wires) and eq(g.slab, o_slab), "full-wire garble differs"
        wires = np.zeros(circ.num_wires, dtype=LABEL_DTYPE)
        wires[: circ.num_inputs] = np.where(bits[0].astype(bool), g.Wires["l1"][: circ.num_inputs], g.Wires["l0"][: circ.num_inputs])
        eng.eval(key, wires, g.Gates)
        want = np.where(np.array(circ.compute_wires(bits[0].tolist()), dtype=bool), o_wires["l1"], o_wires["l0"]) \
            if hasattr(circ, "compute_wires") else None
        if want is not None:
            assert eq(wires, want.astype(LABEL_DTYPE)), "full-wire eval differs"
        else:
            no = circ.num_outputs
            assert np.array_equal(decode(g.Wires[-no:], wires[-no:]), circ.compute_bits(bits[0].tolist()))


@pytest.mark.parametrize("name,batch,teams", [("sha256", 5, "16"), ("aes_128", 9, "10"), ("sha512", 2, "8"), ("mul64", 6, "32")])
def test_hot_cold_plans_are_bit_exact(name, batch, teams, monkeypatch):
    """Plans that keep only a hot subset of the labels in shared memory (the rest in the per-instance L2 scratch, moved
    by per-phase evict / reload lists), so that more instances are resident per SM: forced here for several circuits
    and targets, garble and eval against the oracle."""
    monkeypatch.setenv("GCB_HOT_TEAMS", teams)
    monkeypatch.setenv("GCB_SPREAD", "0")
    circ = load_circuit(name)
    eng = GarbleEngine(circ)
    if name != "mul64":                                            # mul64 already holds 16 instances: nothing to gain
        assert eng.info.num_hot_slots < eng.info.num_slots and eng.info.teams_per_sm >= int(teams) // 2
    keys, rand = garble_inputs(f"hotcold/{name}", batch, circ.num_inputs, 32)
    r, l0 = rand_to_labels(rand, circ.num_inputs)
    tables, io = eng.garble_batch(keys, r, l0)
    _, o_tables, o_io = O.garble_batch(circ, keys, rand, threads=4)
    assert eq(tables, o_tables) and eq(io, o_io)
    bits = np.random.default_rng(2).integers(0, 2, (batch, circ.num_inputs), dtype=np.uint8)
    inl = select(io[:, : circ.num_inputs], bits)
    out = eng.eval_batch(keys, tables, inl)
    assert eq(out, O.eval_batch(circ, keys, o_tables, inl, threads=4))


@pytest.mark.parametrize("batch", [1, 2, 7, 300])
def test_twin_teams_are_bit_exact(batch, monkeypatch):
    """One-warp teams run as lock-step pairs (two instances claimed at once, shared 64-thread barriers); an odd tail is
    run by both warps of the last pair.  Forced here on sha256's eight all-hot teams."""
    monkeypatch.setenv("GCB_TWIN", "1")
    monkeypatch.setenv("GCB_SPREAD", "0")
    circ = load_circuit("sha256")
    eng = GarbleEngine(circ)
    keys, rand = garble_inputs(f"twin/{batch}", batch, circ.num_inputs, 16)
    r, l0 = rand_to_labels(rand, circ.num_inputs)
    tables, io = eng.garble_batch(keys, r, l0)
    sample = sorted({0, batch // 2, batch - 1})
    _, o_tables, o_io = O.garble_batch(circ, keys[sample], rand[sample], threads=4)
    assert eq(tables[sample], o_tables) and eq(io[sample], o_io)
    bits = np.random.default_rng(4).integers(0, 2, (batch, circ.num_inputs), dtype=np.uint8)
    inl = select(io[:, : circ.num_inputs], bits)
    out = eng.eval_batch(keys, tables, inl)
    assert eq(out[sample], O.eval_batch(circ, keys[sample], o_tables, inl[sample], threads=4))
    got = decode(io[:, circ.num_inputs:], out)
    assert got.max() <= 1
    assert np.array_equal(got[batch - 1], circ.compute_bits(bits[batch - 1].tolist()))


def test_default_plan_of_sha256_keeps_eight_instances_with_the_balanced_schedule():
    """The schedule that fills the warp passes needs 1,272 live labels per instance; eight such label blocks fit either
    side of the T-tables because the team headers are kept apart from them.  Everything stays in shared memory."""
    for name in ("sha256", "aes_128"):
        i = GarbleEngine(load_circuit(name)).info
        assert i.num_hot_slots == i.num_slots
    i = GarbleEngine(load_circuit("sha256")).info
    assert i.teams_per_sm == 8 and i.garble_passes < 3100


def _wide_circuit(n_pairs: int):
    """2 * n_pairs inputs, all live until the single AND level: out[i] = in[i] & in[n_pairs + i]."""
    lines = [f"2 1 {i} {n_pairs + i} {2 * n_pairs + i} AND" for i in range(n_pairs)]
    text = f"{n_pairs} {3 * n_pairs}\n2 {n_pairs} {n_pairs}\n1 {n_pairs}\n\n" + "\n".join(lines) + "\n"
    from mpc_b200.circuit_io import parse_bristol
    return parse_bristol(text, f"wide{n_pairs}")


def test_circuit_that_only_fits_beside_two_tables():
    """6,000 labels live at once do not fit beside four T-tables (one team's block must fit below or above the
    tables: about 4,000 slots) but do beside two (about 6,400): the plan falls back to the two-table,
    two-block variant and stays bit-exact."""
    circ = _wide_circuit(2000)
    eng = GarbleEngine(circ)
    assert eng.info.num_slots >= 6000 and eng.info.teams_per_sm == 1
    batch = 3
    keys, rand = garble_inputs("wide", batch, circ.num_inputs, 16)
    r, l0 = rand_to_labels(rand, circ.num_inputs)
    tables, io = eng.garble_batch(keys, r, l0)
    _, o_tables, o_io = O.garble_batch(circ, keys, rand)
    assert eq(tables, o_tables) and eq(io, o_io)
    bits = np.random.default_rng(3).integers(0, 2, (batch, circ.num_inputs), dtype=np.uint8)
    inl = select(io[:, : circ.num_inputs], bits)
    out = eng.eval_batch(keys, tables, inl)
    assert eq(out, O.eval_batch(circ, keys, o_tables, inl))
    got = decode(io[:, circ.num_inputs:], out)
    assert np.array_equal(got, bits[:, :2000] & bits[:, 2000:])


@pytest.mark.parametrize("which", ["wide", "deep"])
def test_circuit_that_spills_labels_to_global_memory(which):
    """More live labels than fit on chip (12,000 in one AND level; 9,000 inputs that stay live through 20,000
    gates of all types): the spilling variant keeps the low slots in shared memory and the rest in a
    global-memory scratch, bit-exact in plain and full-wire mode."""
    circ = _wide_circuit(4000) if which == "wide" else mixed_circuit(21, 20000, 9000, 64)
    eng = GarbleEngine(circ)
    assert eng.info.num_slots > 6400 and eng.info.teams_per_sm == 1
    batch = 4
    keys, rand = garble_inputs(f"spill/{which}", batch, circ.num_inputs, 32)
    r, l0 = rand_to_labels(rand, circ.num_inputs)
    tables, io = eng.garble_batch(keys, r, l0)
    _, o_tables, o_io = O.garble_batch(circ, keys, rand)
    assert eq(tables, o_tables) and eq(io, o_io)
    bits = np.random.default_rng(4).integers(0, 2, (batch, circ.num_inputs), dtype=np.uint8)
    inl = select(io[:, : circ.num_inputs], bits)
    out = eng.eval_batch(keys, tables, inl)
    assert eq(out, O.eval_batch(circ, keys, o_tables, inl))
    assert np.array_equal(decode(io[:, circ.num_inputs:], out)[0], circ.compute_bits(bits[0].tolist()))
    key = DRBG(f"spill/key/{which}").read(16)
    rb = DRBG(f"spill/rand/{which}").read(16 * (1 + circ.num_inputs))
    g = eng.garble(rb, key)
    _, o_wires, o_slab, _ = O.garble(circ, key, rb)
    assert eq(g.Wires, o_wires) and eq(g.slab, o_slab), "full-wire garble differs"


def test_reference_garble_fixture_gpu():
    """The same pin as tests/test_oracle.py::test_reference_garble_fixture, on the CUDA path (skips without the fixture)."""
    from test_oracle import fixture_digests, reference_fixture
    fx = reference_fixture()
    circ, eng = get("sha256xor")
    key, rand = bytes.fromhex(fx["key"]), np.frombuffer(bytes.fromhex(fx["rand"]), dtype=np.uint8)
    r, l0 = rand_to_labels(rand.reshape(1, -1), circ.num_inputs)
    tables, io = eng.garble_batch(key, r, l0)
    assert fixture_digests(tables[0], io[0], circ.num_inputs) == (
        fx["tables_sha256"], fx["input_wires_sha256"], fx["output_wires_sha256"])


@pytest.mark.parametrize("name,batch,klen", [("aes_128", 149, 16), ("aes_128", 300, 32), ("aes_128", 700, 16),
                                             ("sha256", 300, 16), ("mul64", 700, 24), ("sha512", 150, 32)])
def test_batches_below_one_wave_are_spread_over_the_sms(name, batch, klen):
    """A batch smaller than one wave runs with fewer, wider teams per CTA (2-5 teams of up to 256 threads for aes_128,
    64-thread teams for the narrow circuits) on all SMs: sampled instances against the oracle, every instance decoded."""
    circ, eng = get(name)
    assert batch < eng.info.teams_per_sm * 148
    keys, rand = garble_inputs(f"spread/{name}/{batch}", batch, circ.num_inputs, klen)
    r, l0 = rand_to_labels(rand, circ.num_inputs)
    tables, io = eng.garble_batch(keys, r, l0)
    sample = sorted({0, 1, 147, 148, batch // 2, batch - 1})
    _, o_tables, o_io = O.garble_batch(circ, keys[sample], rand[sample], threads=4)
    assert eq(tables[sample], o_tables) and eq(io[sample], o_io)
    bits = np.random.default_rng(batch).integers(0, 2, (batch, circ.num_inputs), dtype=np.uint8)
    inl = select(io[:, : circ.num_inputs], bits)
    out = eng.eval_batch(keys, tables, inl)
    assert eq(out[sample], O.eval_batch(circ, keys[sample], o_tables, inl[sample], threads=4))
    got = decode(io[:, circ.num_inputs:], out)
    assert got.max() <= 1
    for i in (0, batch - 1):
        assert np.array_equal(got[i], circ.compute_bits(bits[i].tolist()))
