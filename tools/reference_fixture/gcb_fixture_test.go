// gcb_fixture_test.go -- run ONCE with the reference and a Go toolchain to pin the oracle to the reference's own
// Garble bytes (SURVEY.md section 8c, item 4: the only reference test that pins Circuit.Garble's output draws its
// randomness from Go's math/rand, which cannot be reproduced without Go).
//
//   cp tools/reference_fixture/gcb_fixture_test.go <mpc checkout>/sha2pc/
//   cd <mpc checkout> && go test ./sha2pc -run TestGCBFixture
//   cp sha2pc/gcb_garble_fixture.json <this repo>/tests/golden/reference_garble_fixture.json
//
// The fixture holds what the deterministic transcript test feeds Circuit.Garble (the 32-byte key, the 16 bytes read
// for R, the 16 bytes read per input wire) and SHA-256 digests of what Garble returned, in the encodings of
// sha2pc/encoding.go: all table rows (encodeGarbledTables order), the input wires and the output wires.
// tests/test_oracle.py::test_reference_garble_fixture and tests/test_gpu_garble.py::test_reference_garble_fixture_gpu
// consume it; without the file they skip and DESIGN.md keeps saying "parity unpinned" for Garble's table bytes.
package sha2pc

import (
	"bytes"
	"crypto/sha256"
	"encoding/hex"
	"encoding/json"
	"io"
	"os"
	"testing"

	"github.com/markkurossi/mpc/ot"
)

type recordingReader struct {
	r   io.Reader
	buf bytes.Buffer
}

func (r *recordingReader) Read(p []byte) (int, error) {
	n, err := r.r.Read(p)
	r.buf.Write(p[:n])
	return n, err
}

func TestGCBFixture(t *testing.T) {
	rng := &recordingReader{r: newDeterministicReader([]byte("garbler-round3"))}
	var key [32]byte
	if _, err := io.ReadFull(rng, key[:]); err != nil {
		t.Fatal(err)
	}
	rng.buf.Reset()
	garbled, err := sha256xorCircuit.Garble(rng, key[:])
	if err != nil {
		t.Fatal(err)
	}
	randBytes := append([]byte(nil), rng.buf.Bytes()...) // 16 for R, then 16 per input wire
	var tmp ot.LabelData
	tables := sha256.New()
	rows := 0
	for _, row := range garbled.Gates {
		for _, l := range row {
			l.GetData(&tmp)
			tables.Write(tmp[:])
			rows++
		}
	}
	wireHash := func(ws []ot.Wire) string {
		h := sha256.New()
		for _, w := range ws {
			w.L0.GetData(&tmp)
			h.Write(tmp[:])
			w.L1.GetData(&tmp)
			h.Write(tmp[:])
		}
		return hex.EncodeToString(h.Sum(nil))
	}
	nin, nout := sha256xorCircuit.Inputs.Size(), sha256xorCircuit.Outputs.Size()
	out := map[string]interface{}{
		"circuit":            "sha256xor.mpclc",
		"key":                hex.EncodeToString(key[:]),
		"rand":               hex.EncodeToString(randBytes),
		"rows":               rows,
		"tables_sha256":      hex.EncodeToString(tables.Sum(nil)),
		"input_wires_sha256": wireHash(garbled.Wires[:nin]),
		"output_wires_sha256": wireHash(garbled.Wires[len(garbled.Wires)-nout:]),
	}
	data, _ := json.MarshalIndent(out, "", " ")
	if err := os.WriteFile("gcb_garble_fixture.json", data, 0o644); err != nil {
		t.Fatal(err)
	}
}
