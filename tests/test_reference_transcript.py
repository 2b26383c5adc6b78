"""The reference's own golden hashes for Circuit.Garble's bytes (sha2pc/sha2pc_test.go:74-131, TestDeterministicTranscript),
reproduced without a Go toolchain: tests/gostd.py restates math/rand, crypto/rand.Int and P-256, tests/sha2pc_transcript.py
the four protocol rounds and their encodings.  Round 3's hash covers the 32-byte AES key, all 42,914 garbled rows in gate
order, the garbler's input labels, both labels of every output wire and (through the OT ciphertexts) both labels of
every evaluator input wire of sha256xor.mpclc -- the byte-level pin of the oracle's Garble (CPU) and of the CUDA path
(GPU) against the reference."""
import numpy as np
import pytest

import gostd
import sha2pc_transcript as T
from conftest import load_circuit
from oracle import pyoracle as O

EXPECTED = (T.EXP_ROUND1, T.EXP_ROUND2, T.EXP_ROUND3, T.EXP_FINAL)


def test_go_math_rand_known_answers():
    """rand.New(rand.NewSource(1)): the sequence every Go program without a seed printed before Go 1.20, and the first
    entries of rngCooked as they stand in math/rand/rng.go -- the derived table is the real one."""
    g = gostd.GoRand(1)
    assert [g.int63() for _ in range(5)] == [5577006791947779410, 8674665223082153551, 6129484611666145821,
                                            4037200794235010051, 3916589616287113937]
    c = gostd.rng_cooked()
    assert [v - (1 << 64) if v >> 63 else v for v in c[:2]] == [-4181792142133755926, -4576982950128230565]
    g = gostd.GoRand(1)
    assert [g.intn_pow2(256) for _ in range(2)] == [(5577006791947779410 >> 32) & 255, (8674665223082153551 >> 32) & 255]


def test_p256_known_answers():
    """NIST point-multiplication vectors for P-256 (k = 2, k = n - 1) and the group law."""
    assert gostd.scalar_base_mult(2) == (0x7CF27B188D034F7E8A52380304B51AC3C08969E277F21B35A60B48FC47669978,
                                         0x07775510DB8ED040293D9AC69F7430DBBA7DADE63CE982299E04B79D227873D1)
    x, y = gostd.scalar_base_mult(gostd.N - 1)
    assert (x, y) == (gostd.GX, gostd.P - gostd.GY)
    p5 = gostd.scalar_base_mult(5)
    assert gostd.add(*gostd.scalar_base_mult(2), *gostd.scalar_base_mult(3)) == p5 and gostd.on_curve(*p5)
    assert gostd.scalar_mult(*gostd.scalar_base_mult(7), 11) == gostd.scalar_base_mult(77)


def test_oracle_garble_reproduces_the_reference_transcript_hashes():
    circ = load_circuit("sha256xor")

    def garble(key, rand):
        _, wires, slab, _ = O.garble(circ, key, rand)
        return wires[:512], wires[-256:], slab

    def evaluate(key, in_labels, slab):
        return O.eval_(circ, key, in_labels, slab)[-256:]

    assert T.run(garble, evaluate) == EXPECTED


def test_a_wrong_row_breaks_the_round3_hash():
    """The pin is sensitive to a single bit of a single row."""
    circ = load_circuit("sha256xor")

    def garble(key, rand):
        _, wires, slab, _ = O.garble(circ, key, rand)
        slab = slab.copy()
        slab["d1"][31337] ^= np.uint64(1)
        return wires[:512], wires[-256:], slab

    r1, r2, r3, _ = T.run(garble)
    assert (r1, r2) == EXPECTED[:2] and r3 != T.EXP_ROUND3


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["single", "batch"])
def test_gpu_garble_reproduces_the_reference_transcript_hashes(path):
    """The same transcript with the CUDA garbler and evaluator in the middle: through the single-instance mirror of
    Circuit.Garble / Eval (full Wires) and through the batch entry points (I/O wires only; the instance sits at an odd
    position of a batch whose other instances use other keys and randomness)."""
    from mpc_b200.circuit import GarbleEngine
    from mpc_b200.circuit_io import LABEL_DTYPE
    from util import rand_to_labels
    circ = load_circuit("sha256xor")
    eng = GarbleEngine(circ)

    if path == "single":
        def garble(key, rand):
            g = eng.garble(rand, key)
            return g.Wires[:512], g.Wires[-256:], g.slab

        def evaluate(key, in_labels, slab):
            wires = np.zeros(circ.num_wires, dtype=LABEL_DTYPE)
            wires[:512] = in_labels
            off = eng.row_off
            eng.eval(key, wires, [slab[off[i]:off[i + 1]] for i in range(circ.num_gates)])
            return wires[-256:]
    else:
        batch, at = 5, 3
        rng = np.random.default_rng(7)

        def garble(key, rand):
            keys = rng.integers(0, 256, (batch, 32), dtype=np.uint8)
            rands = rng.integers(0, 256, (batch, 16 * 513), dtype=np.uint8)
            keys[at] = np.frombuffer(key, dtype=np.uint8)
            rands[at] = np.frombuffer(rand, dtype=np.uint8)
            r, l0 = rand_to_labels(rands, 512)
            tables, io = eng.garble_batch(keys, r, l0)
            garble.keys, garble.tables = keys, tables
            return io[at, :512], io[at, 512:], tables[at]

        def evaluate(key, in_labels, slab):
            inl = np.zeros((batch, 512), dtype=LABEL_DTYPE)
            inl[at] = in_labels
            return eng.eval_batch(garble.keys, garble.tables, inl)[at]

    assert T.run(garble, evaluate) == EXPECTED
