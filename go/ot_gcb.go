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
