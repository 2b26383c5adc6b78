// Replacement bodies for package ot (markkurossi/mpc) over libgcb200:
// IKNPSender.send (ot/iknp.go:197-226), IKNPReceiver.receive (:468-511) and
// MITCCRH.Hash (ot/mitccrh.go:93-128).  Base OTs, Delta, the seed labels and
// all ot.IO traffic stay in Go; the objects keep the 128 / 256 seeds and the
// byte position of the (stateful) CTR streams instead of cipher.Stream values.
// Illustrative: not compiled here (no Go toolchain in the build image).
package ot

import (
	"fmt"
	"unsafe"

	"gcb200/go/gcb"
)

func (s *IKNPSender) send(n int) ([]Label, error) {
	result := make([]Label, n)
	u := make([]byte, 0, gcb.IKNPUSize(n))
	for ofs := 0; ofs < n; { // same framing: one ReceiveData per chunk
		chunk, err := s.io.ReceiveData()
		if err != nil {
			return nil, err
		}
		if len(chunk)%K != 0 {
			return nil, fmt.Errorf("invalid chunk size: %v", len(chunk))
		}
		u = append(u, chunk...)
		ofs += len(chunk) / K * 8
	}
	err := gcb.IKNPSenderExpand((*[128]gcb.Label)(unsafe.Pointer(&s.k0)), (*gcb.Label)(unsafe.Pointer(&s.Delta)),
		s.pos, u, n, unsafe.Slice((*gcb.Label)(unsafe.Pointer(&result[0])), n))
	s.pos += gcb.IKNPStreamAdvance(n)
	return result, err
}

func (r *IKNPReceiver) receive(b []bool, result []Label) error {
	if len(b) != len(result) {
		panic("len(b) != len(result)")
	}
	u := make([]byte, gcb.IKNPUSize(len(b)))
	err := gcb.IKNPReceiverExpand((*[128]gcb.Label)(unsafe.Pointer(&r.k0)), (*[128]gcb.Label)(unsafe.Pointer(&r.k1)),
		r.pos, b, u, unsafe.Slice((*gcb.Label)(unsafe.Pointer(&result[0])), len(result)))
	if err != nil {
		return err
	}
	r.pos += gcb.IKNPStreamAdvance(len(b))
	for ofs := 0; ofs < len(u); { // one SendData per chunk, as on the reference wire
		end := ofs + chunkSize
		if end > len(u) {
			end = len(u)
		}
		if err := r.io.SendData(u[ofs:end]); err != nil {
			return err
		}
		ofs = end
	}
	return r.io.Flush()
}

// Hash hashes blks in place; key i of the call is global key gid+keyUsed+i.
func (m *MITCCRH) Hash(blks []Label, k, h int) {
	if k > m.batchSize || m.batchSize%k != 0 || len(blks) != k*h {
		panic("MITCCRH.Hash: invalid arguments")
	}
	if m.keyUsed == m.batchSize {
		m.gidBase = m.gid // renewKeys: keys gid .. gid+batchSize-1
		m.gid += uint64(m.batchSize)
		m.keyUsed = 0
	}
	if err := gcb.MITCCRHHash((*gcb.Label)(unsafe.Pointer(&m.startPoint)), m.gidBase+uint64(m.keyUsed),
		unsafe.Slice((*gcb.Label)(unsafe.Pointer(&blks[0])), len(blks)), k, h); err != nil {
		panic(err)
	}
	m.keyUsed += k
}

// Send (ot/cot.go:136-181) with the per-batch MiTCCRH loop replaced by one call; the 2n messages
// come back in the SendLabel encoding and go out with a single Write.
func (cot *COT) Send(wires []Wire) error {
	if cot.iknpS == nil {
		return fmt.Errorf("not initialized as sender")
	}
	data, err := cot.iknpS.Send(len(wires), cot.malicious)
	if err != nil {
		return err
	}
	seed, err := NewLabel(cot.r)
	if err != nil {
		return err
	}
	var ld LabelData
	if err := cot.io.SendLabel(seed, &ld); err != nil {
		return err
	}
	if err := cot.io.Flush(); err != nil {
		return err
	}
	msgs := make([]byte, 32*len(wires))
	if err := gcb.COTSend((*gcb.Label)(unsafe.Pointer(&seed)), (*gcb.Label)(unsafe.Pointer(&cot.iknpS.Delta)),
		unsafe.Slice((*gcb.Label)(unsafe.Pointer(&data[0])), len(data)),
		unsafe.Slice((*gcb.Wire)(unsafe.Pointer(&wires[0])), len(wires)), msgs); err != nil {
		return err
	}
	if err := cot.io.SendBytes(msgs); err != nil { // same bytes as 2n SendLabel calls
		return err
	}
	return cot.io.Flush()
}

// The chi loop of IKNPSender.Send's malicious branch (ot/iknp.go:150-173): result first, then the
// 256 choice-vector OTs on the same chi stream.
func (s *IKNPSender) checkSums(seed2 Label, result, choiceVector []Label) (q0, q1 Label, err error) {
	a, err := gcb.IKNPCheckSums((*gcb.Label)(unsafe.Pointer(&seed2)), 0,
		unsafe.Slice((*gcb.Label)(unsafe.Pointer(&result[0])), len(result)), nil)
	if err != nil {
		return
	}
	b, err := gcb.IKNPCheckSums((*gcb.Label)(unsafe.Pointer(&seed2)), uint64(len(result)),
		unsafe.Slice((*gcb.Label)(unsafe.Pointer(&choiceVector[0])), len(choiceVector)), nil)
	q0 = Label{D0: a[0].D0 ^ b[0].D0, D1: a[0].D1 ^ b[0].D1}
	q1 = Label{D0: a[1].D0 ^ b[1].D0, D1: a[1].D1 ^ b[1].D1}
	return
}
