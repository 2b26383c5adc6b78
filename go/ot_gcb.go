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

// ReceiveBits (ot/iknp.go:554-620): the same U exchange as receive; the result is Bit(0) of every label, packed.
func (r *IKNPReceiver) ReceiveBits(choices []uint64, result []uint64, n int) error {
	if (n+63)/64 > len(choices) {
		return fmt.Errorf("choices buffer len=%d too short for n=%d", len(choices), n)
	}
	u := make([]byte, gcb.IKNPUSize(n))
	err := gcb.IKNPReceiverExpandBits((*[128]gcb.Label)(unsafe.Pointer(&r.k0)), (*[128]gcb.Label)(unsafe.Pointer(&r.k1)),
		r.pos, choices, n, u, result)
	if err != nil {
		return err
	}
	r.pos += gcb.IKNPStreamAdvance(n)
	for ofs := 0; ofs < len(u); {
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

// SendBits (ot/iknp.go:259-310): column 0 of the matrix, packed LSB-first.
func (s *IKNPSender) SendBits(result []uint64, n int) error {
	u := make([]byte, 0, gcb.IKNPUSize(n))
	for ofs := 0; ofs < n; {
		chunk, err := s.io.ReceiveData()
		if err != nil {
			return err
		}
		if len(chunk)%K != 0 {
			return fmt.Errorf("invalid chunk size: %v", len(chunk))
		}
		u = append(u, chunk...)
		ofs += len(chunk) / K * 8
	}
	err := gcb.IKNPSenderExpandBits((*[128]gcb.Label)(unsafe.Pointer(&s.k0)), (*gcb.Label)(unsafe.Pointer(&s.Delta)), s.pos, u, n, result)
	s.pos += gcb.IKNPStreamAdvance(n)
	return err
}

// Receive (ot/cot.go:184-235): the extension, then one call instead of the per-batch MiTCCRH loop; the 2n messages
// are read off the connection in the SendLabel encoding with one ReceiveBytes.
func (cot *COT) Receive(flags []bool, result []Label) error {
	if cot.iknpR == nil {
		return fmt.Errorf("not initialized as receiver")
	}
	if err := cot.iknpR.Receive(flags, result, cot.malicious); err != nil {
		return err
	}
	var ld LabelData
	var seed Label
	if err := cot.io.ReceiveLabel(&seed, &ld); err != nil {
		return err
	}
	msgs := make([]byte, 32*len(result))
	if err := cot.io.ReceiveBytes(msgs); err != nil {
		return err
	}
	return gcb.COTReceive((*gcb.Label)(unsafe.Pointer(&seed)), flags, msgs,
		unsafe.Slice((*gcb.Label)(unsafe.Pointer(&result[0])), len(result)))
}

// ROT.Send / ROT.Receive after the extension (ot/rot.go:155-172, 192-197).
func (rot *ROT) hashSend(seed Label, q []Label, wires []Wire) error {
	return gcb.ROTSend((*gcb.Label)(unsafe.Pointer(&seed)), (*gcb.Label)(unsafe.Pointer(&rot.iknpS.Delta)),
		unsafe.Slice((*gcb.Label)(unsafe.Pointer(&q[0])), len(q)), unsafe.Slice((*gcb.Wire)(unsafe.Pointer(&wires[0])), len(wires)))
}
func (rot *ROT) hashReceive(seed Label, result []Label) error {
	return gcb.ROTReceive((*gcb.Label)(unsafe.Pointer(&seed)), unsafe.Slice((*gcb.Label)(unsafe.Pointer(&result[0])), len(result)))
}
