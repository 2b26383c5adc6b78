// gcb_batch.go -- the entry points added for throughput callers: device lists, asynchronous jobs, the streaming
// evaluator and the bit-COT variants.  Same rules as gcb.go: every C call and its error fetch run on one locked OS
// thread (call), every handle is kept alive across the call (runtime.KeepAlive), and memory that C keeps using
// after a call returns (the _Begin forms) must be C memory from HostAlloc -- never Go-managed memory.
//
// NOT COMPILED IN THIS REPOSITORY'S CI (no Go toolchain in the build image).
package gcb

/*
#include "gcb200.h"
*/
import "C"

import (
	"runtime"
	"unsafe"
)

// SetDevices is the engine-level multi-GPU setting (SURVEY.md 8b): one GarbleBatch / EvalBatch / IKNP call is
// fanned out over these devices inside the library.  Process-wide; nil clears it.
func SetDevices(ids []int) error {
	c := make([]C.int, len(ids))
	for i, v := range ids {
		c[i] = C.int(v)
	}
	var p *C.int
	if len(c) > 0 {
		p = &c[0]
	}
	return call(func() C.int { return C.gcb_set_devices(p, C.int(len(c))) })
}

// DeviceCount returns the number of CUDA devices.
func DeviceCount() int { return int(C.gcb_device_count()) }

// HostLabels / HostWires view HostAlloc memory as slices (C memory: it may be handed to the _Begin forms).
func HostLabels(n int) []Label { return unsafe.Slice((*Label)(HostAlloc(16*n)), n) }
func HostWires(n int) []Wire   { return unsafe.Slice((*Wire)(HostAlloc(32*n)), n) }

// Job is a queued GarbleBegin / EvalBegin call.
type Job struct {
	h *C.gcb_job
	p *Plan
}

// GarbleBegin queues gcb_garble for `batch` instances and returns; every slice must be HostAlloc memory and stay
// untouched until Wait.  One goroutine keeps all selected devices and both PCIe directions busy this way.
func (p *Plan) GarbleBegin(key []byte, keyStride, batch int, r, l0, tables []Label, ioWires []Wire) (*Job, error) {
	j := &Job{p: p}
	err := call(func() C.int {
		return C.gcb_garble_begin(p.h, (*C.uint8_t)(unsafe.Pointer(&key[0])), C.uint32_t(len(key)/max(1, batchIf(keyStride, batch))),
			C.uint32_t(keyStride), C.uint32_t(batch), lp(r), lp(l0), lp(tables), wp(ioWires), nil, 0, &j.h)
	})
	runtime.KeepAlive(p)
	if err != nil {
		return nil, err
	}
	return j, nil
}

// EvalBegin queues gcb_eval; same memory rules as GarbleBegin.
func (p *Plan) EvalBegin(key []byte, keyStride, batch int, tables, in, out []Label) (*Job, error) {
	j := &Job{p: p}
	err := call(func() C.int {
		return C.gcb_eval_begin(p.h, (*C.uint8_t)(unsafe.Pointer(&key[0])), C.uint32_t(len(key)/max(1, batchIf(keyStride, batch))),
			C.uint32_t(keyStride), C.uint32_t(batch), lp(tables), lp(in), lp(out), nil, 0, &j.h)
	})
	runtime.KeepAlive(p)
	if err != nil {
		return nil, err
	}
	return j, nil
}

// Wait blocks until the job's results are in the caller's buffers; the job is consumed.
func (j *Job) Wait() error {
	h := j.h
	j.h = nil
	if h == nil {
		return nil
	}
	err := call(func() C.int { return C.gcb_job_wait(h) })
	runtime.KeepAlive(j.p)
	return err
}

// Done polls without blocking.
func (j *Job) Done() bool { return j.h == nil || C.gcb_job_done(j.h) != 0 }

// StreamEval wraps gcb_seval: circuit.StreamEval plus the OpCircuit gate loop of StreamEvaluator.
type StreamEval struct{ h *C.gcb_seval }

// NewStreamEval mirrors circuit.NewStreamEval (stream_evaluator.go:37-55) for batch = 1.
func NewStreamEval(key []byte) (*StreamEval, error) {
	s := &StreamEval{}
	if err := call(func() C.int {
		return C.gcb_seval_create((*C.uint8_t)(unsafe.Pointer(&key[0])), C.uint32_t(len(key)), 0, 1, &s.h)
	}); err != nil {
		return nil, err
	}
	runtime.SetFinalizer(s, func(s *StreamEval) { C.gcb_seval_destroy(s.h) })
	return s, nil
}

// Set / Get mirror StreamEval.Set / SetInputs / Get (stream_evaluator.go:57-96).
func (s *StreamEval) Set(ids []uint32, labels []Label) error {
	defer runtime.KeepAlive(s)
	return call(func() C.int { return C.gcb_seval_set_wires(s.h, u32p(ids), C.uint32_t(len(ids)), lp(labels)) })
}
func (s *StreamEval) Get(ids []uint32, labels []Label) error {
	defer runtime.KeepAlive(s)
	return call(func() C.int { return C.gcb_seval_get_wires(s.h, u32p(ids), C.uint32_t(len(ids)), lp(labels)) })
}

// Circuit evaluates one OpCircuit body whose `ngates` gate records are in src (the bytes read off the connection
// after the header); returns the bytes consumed.
func (s *StreamEval) Circuit(src []byte, ngates, ntmp, nwires int) (int, error) {
	var used C.size_t
	err := call(func() C.int {
		return C.gcb_seval_circuit(s.h, (*C.uint8_t)(unsafe.Pointer(&src[0])), C.size_t(len(src)), C.size_t(len(src)),
			C.uint32_t(ngates), C.uint32_t(ntmp), C.uint32_t(nwires), &used)
	})
	runtime.KeepAlive(s)
	return int(used), err
}

// IKNPReceiverExpandBits / IKNPSenderExpandBits replace ReceiveBits / SendBits (ot/iknp.go:554-620, 259-310).
func IKNPReceiverExpandBits(k0, k1 *[128]Label, pos uint64, choices []uint64, n int, u []byte, result []uint64) error {
	return call(func() C.int {
		return C.gcb_iknp_receiver_expand_bits(lp(k0[:]), lp(k1[:]), C.uint64_t(pos), (*C.uint64_t)(unsafe.Pointer(&choices[0])),
			C.uint64_t(n), (*C.uint8_t)(unsafe.Pointer(&u[0])), (*C.uint64_t)(unsafe.Pointer(&result[0])))
	})
}
func IKNPSenderExpandBits(k *[128]Label, delta *Label, pos uint64, u []byte, n int, result []uint64) error {
	return call(func() C.int {
		return C.gcb_iknp_sender_expand_bits(lp(k[:]), (*C.gcb_label)(unsafe.Pointer(delta)), C.uint64_t(pos),
			(*C.uint8_t)(unsafe.Pointer(&u[0])), C.size_t(len(u)), C.uint64_t(n), (*C.uint64_t)(unsafe.Pointer(&result[0])))
	})
}

// SetWires mirrors Streaming.Set for the garbler's wire file is not needed: Streaming.Garble writes its outputs itself
// (stream_garble.go:144-157).
