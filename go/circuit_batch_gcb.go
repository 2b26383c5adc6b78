// Batched and streaming additions to package circuit over libgcb200: GarbleBatch / EvalBatch (the new throughput
// entry points SURVEY.md 8b asks for beside the unchanged Garble / Eval), NewStreaming and the evaluator's OpCircuit
// loop.  Illustrative: not compiled here (no Go toolchain in the build image).
package circuit

import (
	"io"
	"unsafe"

	"github.com/markkurossi/mpc/ot"

	"gcb200/go/gcb"
)

// FullWires selects what Garble puts into Garbled.Wires.  true (the default) is the reference's contract: every one of
// the NumWires entries.  false fills only the entries callers read -- the input wires and the last Outputs.Size() wires
// (circuit/garbler.go:85-102,153; sha2pc/garbler.go:113-126) -- which lets the engine run its flattened Free-XOR
// plan and skips a 32 * NumWires byte copy per instance (4.3 MB for sha256): see README "batch = 1 latency".
var FullWires = true

// GarbledBatch is the result of GarbleBatch: page-locked slabs, released with Release (the role of Garbled.Release
// and the per-circuit scratch pool, circuit/garble.go:193-244).
type GarbledBatch struct {
	Batch   int
	R       []ot.Label // [batch], S bit set
	Tables  []ot.Label // [batch][rows]: rows of gate g of instance i at i*rows + RowOff[g]
	IOWires []ot.Wire  // [batch][inputs+outputs]
	RowOff  []uint32
}

// Release returns the slabs.
func (g *GarbledBatch) Release() {
	gcb.HostFree(unsafe.Pointer(unsafe.SliceData(g.Tables)))
	gcb.HostFree(unsafe.Pointer(unsafe.SliceData(g.IOWires)))
	g.Tables, g.IOWires = nil, nil
}

// GarbleBatch garbles `batch` independent instances of the circuit with one call (fanned out over gcb.SetDevices).
// Randomness is read per instance in the reference's order: R, then one L0 per input wire.
func (c *Circuit) GarbleBatch(rand io.Reader, key []byte, batch int) (*GarbledBatch, error) {
	p, err := c.plan()
	if err != nil {
		return nil, err
	}
	nin, nout, rows := c.Inputs.Size(), c.Outputs.Size(), int(p.Info.num_rows)
	r := make([]ot.Label, batch)
	l0 := make([]ot.Label, batch*nin)
	for i := 0; i < batch; i++ {
		if r[i], err = ot.NewLabel(rand); err != nil {
			return nil, err
		}
		for k := 0; k < nin; k++ {
			if l0[i*nin+k], err = ot.NewLabel(rand); err != nil {
				return nil, err
			}
		}
	}
	g := &GarbledBatch{Batch: batch, R: r, RowOff: p.RowOff}
	tables, io := gcb.HostLabels(batch*rows), gcb.HostWires(batch*(nin+nout))
	g.Tables = unsafe.Slice((*ot.Label)(unsafe.Pointer(unsafe.SliceData(tables))), len(tables))
	g.IOWires = unsafe.Slice((*ot.Wire)(unsafe.Pointer(unsafe.SliceData(io))), len(io))
	err = p.Garble(key, 0, batch, unsafe.Slice((*gcb.Label)(unsafe.Pointer(&r[0])), batch),
		unsafe.Slice((*gcb.Label)(unsafe.Pointer(unsafe.SliceData(l0))), len(l0)), tables, io, nil)
	if err != nil {
		g.Release()
		return nil, err
	}
	for i := range r {
		r[i].SetS(true)
	}
	return g, nil
}

// EvalBatch evaluates `batch` instances: tables [batch][rows], in [batch][inputs] -> out [batch][outputs].
func (c *Circuit) EvalBatch(key []byte, batch int, tables, in, out []ot.Label) error {
	p, err := c.plan()
	if err != nil {
		return err
	}
	cast := func(l []ot.Label) []gcb.Label {
		return unsafe.Slice((*gcb.Label)(unsafe.Pointer(unsafe.SliceData(l))), len(l))
	}
	return p.Eval(key, 0, batch, cast(tables), cast(in), cast(out), nil)
}

// NewStreaming (circuit/stream_garble.go:41-76): R and the input labels are drawn in Go; the wire file lives on the device.
func NewStreaming(cfg *env.Config, key []byte, inputs []Wire, conn *p2p.Conn) (*Streaming, error) {
	r, err := ot.NewLabel(cfg.GetRandom())
	if err != nil {
		return nil, err
	}
	ids := make([]uint32, len(inputs))
	l0 := make([]gcb.Label, len(inputs))
	for i, w := range inputs {
		ids[i] = uint32(w)
		l, err := ot.NewLabel(cfg.GetRandom())
		if err != nil {
			return nil, err
		}
		l0[i] = *(*gcb.Label)(unsafe.Pointer(&l))
	}
	dev, err := gcb.NewStream(key, *(*gcb.Label)(unsafe.Pointer(&r)), ids, l0)
	if err != nil {
		return nil, err
	}
	r.SetS(true)
	return &Streaming{conn: conn, r: r, dev: dev}, nil
}

// GetInputs (circuit/stream_garble.go:117-128) reads the wires back from the device wire file.
func (stream *Streaming) GetInputs(inputs []Wire) ([]ot.Wire, error) {
	ids := unsafe.Slice((*uint32)(unsafe.Pointer(unsafe.SliceData(inputs))), len(inputs))
	w, err := stream.dev.GetWires(ids)
	return unsafe.Slice((*ot.Wire)(unsafe.Pointer(unsafe.SliceData(w))), len(w)), err
}

// evalCircuit is the OpCircuit case of StreamEvaluator (circuit/stream_evaluator.go:226-432): the header has been read
// (step, numGates, numTmpWires, numWires); the body is buffered once -- its length follows from the gate records -- and
// evaluated on the device.  The wire file is read with streaming.dev.Get when outputs are decoded (:181-199).
func (streaming *StreamEval) evalCircuit(conn *p2p.Conn, numGates, numTmpWires, numWires int) error {
	body, err := readGateRecords(conn, numGates) // buffers exactly the records of this circuit
	if err != nil {
		return err
	}
	_, err = streaming.dev.Circuit(body, numGates, numTmpWires, numWires)
	return err
}
