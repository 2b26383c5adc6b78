// Replacement bodies for package circuit (markkurossi/mpc) over libgcb200.
// Signatures are the reference's (circuit/garble.go:248, circuit/eval.go:17,
// circuit/stream_garble.go:161); only the bodies change.  Illustrative: not
// compiled here (no Go toolchain in the build image).
package circuit

import (
	"fmt"
	"io"
	"time"
	"unsafe"

	"github.com/markkurossi/mpc/ot"

	"gcb200/go/gcb"
)

// plan returns the circuit's compiled plan, built once (replaces garblePool).
func (c *Circuit) plan() (*gcb.Plan, error) {
	if p := c.gcbPlan.Load(); p != nil {
		return p, nil
	}
	gates := unsafe.Slice((*gcb.Gate)(unsafe.Pointer(&c.Gates[0])), len(c.Gates))
	p, err := gcb.NewPlan(gates, c.NumWires, c.Inputs.Size(), c.Outputs.Size())
	if err != nil {
		return nil, err
	}
	c.gcbPlan.CompareAndSwap(nil, p)
	return c.gcbPlan.Load(), nil
}

// Garble garbles the circuit (same contract as circuit/garble.go:248-308).
func (c *Circuit) Garble(rand io.Reader, key []byte) (*Garbled, error) {
	p, err := c.plan()
	if err != nil {
		return nil, err
	}
	// Randomness is consumed by Go in the reference's order: R, then one L0 per input wire.
	r, err := ot.NewLabel(rand)
	if err != nil {
		return nil, err
	}
	nin := c.Inputs.Size()
	l0 := make([]ot.Label, nin)
	for i := range l0 {
		if l0[i], err = ot.NewLabel(rand); err != nil {
			return nil, err
		}
	}
	wires := make([]ot.Wire, c.NumWires)
	slab := make([]ot.Label, p.Info.num_rows)
	err = p.Garble(key, 0, 1,
		[]gcb.Label{*(*gcb.Label)(unsafe.Pointer(&r))},
		unsafe.Slice((*gcb.Label)(unsafe.Pointer(&l0[0])), nin),
		unsafe.Slice((*gcb.Label)(unsafe.Pointer(&slab[0])), len(slab)),
		nil, unsafe.Slice((*gcb.Wire)(unsafe.Pointer(&wires[0])), len(wires)))
	if err != nil {
		return nil, err
	}
	r.SetS(true)
	// Gates[i] is a sub-slice of the slab, nil for XOR/XNOR (garble.go:292-298).
	gates := make([][]ot.Label, c.NumGates)
	for i := range gates {
		if lo, hi := p.RowOff[i], p.RowOff[i+1]; hi > lo {
			gates[i] = slab[lo:hi:hi]
		}
	}
	return &Garbled{R: r, Wires: wires, Gates: gates}, nil
}

// Eval evaluates the circuit in place on wires (circuit/eval.go:17-115).
func (c *Circuit) Eval(key []byte, wires []ot.Label, garbled [][]ot.Label) error {
	p, err := c.plan()
	if err != nil {
		return err
	}
	slab := make([]ot.Label, 0, p.Info.num_rows)
	for i := range c.Gates {
		rows := garbled[i]
		switch c.Gates[i].Op {
		case AND:
			if len(rows) != 2 {
				return fmt.Errorf("corrupted ciruit: AND row length: %d", len(rows))
			}
			slab = append(slab, rows...)
		case OR:
			if len(rows) < 3 {
				return fmt.Errorf("corrupted circuit: index %d >= row %d", 2, len(rows))
			}
			slab = append(slab, rows[:3]...)
		case INV:
			if len(rows) < 1 {
				return fmt.Errorf("corrupted circuit: index %d >= row %d", 0, len(rows))
			}
			slab = append(slab, rows[0])
		}
	}
	nin, nout := c.Inputs.Size(), c.Outputs.Size()
	full := unsafe.Slice((*gcb.Label)(unsafe.Pointer(&wires[0])), len(wires))
	in := append([]gcb.Label(nil), full[:nin]...) // wires is rewritten in place
	out := make([]gcb.Label, nout)
	return p.Eval(key, 0, 1, unsafe.Slice((*gcb.Label)(unsafe.Pointer(&slab[0])), len(slab)), in, out, full)
}

// Garble garbles one sub-circuit and streams it (circuit/stream_garble.go:161-191).
func (stream *Streaming) Garble(c *Circuit, in, out []Wire) (time.Duration, time.Duration, error) {
	p, err := c.plan()
	if err != nil {
		return 0, 0, err
	}
	ids := func(w []Wire) []uint32 { return unsafe.Slice((*uint32)(unsafe.Pointer(unsafe.SliceData(w))), len(w)) }
	n, err := stream.dev.StepSize(p, ids(in), ids(out))
	if err != nil {
		return 0, 0, err
	}
	if cap(stream.scratch) < n {
		stream.scratch = make([]byte, n)
	}
	_, t0, t1, err := stream.dev.Garble(p, ids(in), ids(out), stream.scratch[:n])
	if err != nil {
		return 0, 0, err
	}
	// Hand the records to the unmodified transport in <= 64 KiB pieces (p2p.Conn.NeedSpace).
	for buf := stream.scratch[:n]; len(buf) > 0; {
		if err := stream.conn.NeedSpace(512); err != nil {
			return 0, 0, err
		}
		k := copy(stream.conn.WriteBuf[stream.conn.WritePos:], buf)
		stream.conn.WritePos += k
		buf = buf[k:]
	}
	return time.Duration(t0), time.Duration(t1), nil
}
