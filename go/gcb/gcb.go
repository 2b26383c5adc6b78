// Package gcb is the cgo binding of libgcb200 (include/gcb200.h), the
// B200-native garble / eval / OT-extension engine.  It is the only file that
// touches C; go/circuit_gcb.go and go/ot_gcb.go show the bodies that replace
// the reference's Go loops behind unchanged signatures.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain
// (see DESIGN.md section 1).  The calls, argument order and memory rules are
// exactly the ones the Python mirror (mpc_b200/circuit.py, ot.py) exercises in
// the parity tests.
package gcb

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../mpc_b200 -lgcb200 -Wl,-rpath,${SRCDIR}/../../mpc_b200
#include <stdlib.h>
#include "gcb200.h"
*/
import "C"

import (
	"errors"
	"runtime"
	"unsafe"
)

// Label mirrors ot.Label: {D0, D1 uint64}, 16 bytes, D0 the high half.
type Label struct{ D0, D1 uint64 }

// Wire mirrors ot.Wire.
type Wire struct{ L0, L1 Label }

// Gate mirrors circuit.Gate (20 bytes, circuit/circuit_test.go:14-19).
type Gate struct {
	Input0, Input1, Output uint32
	Op                     uint8
	_                      [3]uint8
	Level                  uint32
}

// call runs one C entry point and fetches its error text on the SAME OS thread: gcb_last_error() is
// thread-local in C, and a goroutine may migrate between two cgo calls unless it is locked.
func call(f func() C.int) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	if rc := f(); rc != 0 {
		return errors.New(C.GoString(C.gcb_last_error()))
	}
	return nil
}

// lastError is kept for the one-line wrappers below; they go through call() so that the message is read on the
// thread that produced it.
func lastError(rc C.int) error {
	if rc == 0 {
		return nil
	}
	return errors.New(C.GoString(C.gcb_last_error()))
}

// Plan is the compiled, immutable form of one circuit.Circuit; safe for
// concurrent use (sha256xorCircuit is a package singleton in sha2pc).
type Plan struct {
	h      *C.gcb_plan
	Info   C.gcb_plan_info
	RowOff []uint32 // slab offset of every gate's first row (len = ngates+1)
}

// NewPlan replaces the lazily built scratch of Circuit.Garble
// (circuit/garble.go:193-224).  gates is Circuit.Gates as is.
func NewPlan(gates []Gate, numWires, numInputs, numOutputs int) (*Plan, error) {
	p := &Plan{}
	var g *C.gcb_gate
	if len(gates) > 0 {
		g = (*C.gcb_gate)(unsafe.Pointer(&gates[0]))
	}
	if err := call(func() C.int {
		return C.gcb_plan_create(g, C.uint32_t(len(gates)), C.uint32_t(numWires),
			C.uint32_t(numInputs), C.uint32_t(numOutputs), &p.h)
	}); err != nil {
		return nil, err
	}
	C.gcb_plan_get_info(p.h, &p.Info)
	p.RowOff = make([]uint32, len(gates)+1)
	C.gcb_plan_row_offsets(p.h, (*C.uint32_t)(unsafe.Pointer(&p.RowOff[0])))
	runtime.SetFinalizer(p, func(p *Plan) { C.gcb_plan_destroy(p.h) })
	return p, nil
}

// InfoForBatch reports the plan a call with `batch` instances per device runs on: deep, narrow circuits get a second
// plan with split live ranges (more resident instances per SM) when the batch overflows the default one.
func (p *Plan) InfoForBatch(batch int) (C.gcb_plan_info, error) {
	defer runtime.KeepAlive(p)
	var info C.gcb_plan_info
	err := call(func() C.int { return C.gcb_plan_get_info_for_batch(p.h, C.uint64_t(batch), &info) })
	return info, err
}

// HostAlloc returns page-locked memory for slabs that are pooled per circuit
// (replaces garbleScratchPool); such buffers are DMA'd in place.
func HostAlloc(n int) unsafe.Pointer { return C.gcb_host_alloc(C.size_t(n)) }

// HostFree releases HostAlloc memory.
func HostFree(p unsafe.Pointer) { C.gcb_host_free(p) }

// Garble runs gcb_garble for `batch` instances.  r: batch raw R draws;
// l0: batch*numInputs input L0 draws (reader order, garble.go:253-278);
// tables: batch*rows; ioWires: batch*(in+out) or nil; wiresFull: batch*numWires or nil.
func (p *Plan) Garble(key []byte, keyStride, batch int, r, l0, tables []Label, ioWires, wiresFull []Wire) error {
	defer runtime.KeepAlive(p) // the finalizer must not free the plan while C still uses it
	return call(func() C.int {
		return C.gcb_garble(p.h, (*C.uint8_t)(unsafe.Pointer(&key[0])), C.uint32_t(len(key)/max(1, batchIf(keyStride, batch))),
			C.uint32_t(keyStride), C.uint32_t(batch), lp(r), lp(l0), lp(tables), wp(ioWires), wp(wiresFull), 0)
	})
}

// Eval runs gcb_eval.  tables: batch*rows; in: batch*numInputs; out: batch*numOutputs;
// wiresFull: batch*numWires or nil.
func (p *Plan) Eval(key []byte, keyStride, batch int, tables, in, out, wiresFull []Label) error {
	defer runtime.KeepAlive(p)
	return call(func() C.int {
		return C.gcb_eval(p.h, (*C.uint8_t)(unsafe.Pointer(&key[0])), C.uint32_t(len(key)/max(1, batchIf(keyStride, batch))),
			C.uint32_t(keyStride), C.uint32_t(batch), lp(tables), lp(in), lp(out), lp(wiresFull), 0)
	})
}

func batchIf(stride, batch int) int {
	if stride == 0 {
		return 1
	}
	return batch
}

func lp(l []Label) *C.gcb_label {
	if len(l) == 0 {
		return nil
	}
	return (*C.gcb_label)(unsafe.Pointer(&l[0]))
}
func wp(w []Wire) *C.gcb_wire {
	if len(w) == 0 {
		return nil
	}
	return (*C.gcb_wire)(unsafe.Pointer(&w[0]))
}

// Stream wraps gcb_stream (circuit.Streaming's device state).
type Stream struct{ h *C.gcb_stream }

// NewStream mirrors NewStreaming (stream_garble.go:41-76) for batch = 1.
func NewStream(key []byte, r Label, ids []uint32, l0 []Label) (*Stream, error) {
	s := &Stream{}
	var idp *C.uint32_t
	if len(ids) > 0 {
		idp = (*C.uint32_t)(unsafe.Pointer(&ids[0]))
	}
	if err := lastError(C.gcb_stream_create((*C.uint8_t)(unsafe.Pointer(&key[0])), C.uint32_t(len(key)), 0, 1,
		(*C.gcb_label)(unsafe.Pointer(&r)), idp, C.uint32_t(len(ids)), lp(l0), &s.h)); err != nil {
		return nil, err
	}
	runtime.SetFinalizer(s, func(s *Stream) { C.gcb_stream_destroy(s.h) })
	return s, nil
}

// GetWires mirrors Streaming.GetInput(s).
func (s *Stream) GetWires(ids []uint32) ([]Wire, error) {
	out := make([]Wire, len(ids))
	if len(ids) == 0 {
		return out, nil
	}
	err := lastError(C.gcb_stream_get_wires(s.h, (*C.uint32_t)(unsafe.Pointer(&ids[0])), C.uint32_t(len(ids)), wp(out)))
	return out, err
}

// Garble mirrors Streaming.Garble: returns the record stream and the two durations.
func (s *Stream) Garble(p *Plan, in, out []uint32, dst []byte) (n int, nsInit, nsGarble uint64, err error) {
	var w C.size_t
	var t0, t1 C.uint64_t
	err = lastError(C.gcb_stream_garble(s.h, p.h, u32p(in), C.uint32_t(len(in)), u32p(out), C.uint32_t(len(out)),
		(*C.uint8_t)(unsafe.Pointer(&dst[0])), C.size_t(len(dst)), &w, &t0, &t1))
	return int(w), uint64(t0), uint64(t1), err
}

// StepSize is the number of stream bytes Garble will produce.
func (s *Stream) StepSize(p *Plan, in, out []uint32) (int, error) {
	var n C.size_t
	err := lastError(C.gcb_stream_step_size(s.h, p.h, u32p(in), C.uint32_t(len(in)), u32p(out), C.uint32_t(len(out)), &n))
	return int(n), err
}

func u32p(v []uint32) *C.uint32_t {
	if len(v) == 0 {
		return nil
	}
	return (*C.uint32_t)(unsafe.Pointer(&v[0]))
}

// IKNPReceiverExpand / IKNPSenderExpand replace the chunk loops of ot/iknp.go.
func IKNPReceiverExpand(k0, k1 *[128]Label, pos uint64, choice []bool, u []byte, labels []Label) error {
	return lastError(C.gcb_iknp_receiver_expand(lp(k0[:]), lp(k1[:]), C.uint64_t(pos),
		(*C.uint8_t)(unsafe.Pointer(&choice[0])), C.uint64_t(len(choice)), (*C.uint8_t)(unsafe.Pointer(&u[0])), lp(labels)))
}
func IKNPSenderExpand(k *[128]Label, delta *Label, pos uint64, u []byte, n int, labels []Label) error {
	return lastError(C.gcb_iknp_sender_expand(lp(k[:]), (*C.gcb_label)(unsafe.Pointer(delta)), C.uint64_t(pos),
		(*C.uint8_t)(unsafe.Pointer(&u[0])), C.size_t(len(u)), C.uint64_t(n), lp(labels)))
}
func IKNPUSize(n int) int            { return int(C.gcb_iknp_u_size(C.uint64_t(n))) }
func IKNPStreamAdvance(n int) uint64 { return uint64(C.gcb_iknp_stream_advance(C.uint64_t(n))) }

// MITCCRHHash replaces MITCCRH.Hash over nkeys consecutive keys.
func MITCCRHHash(seed *Label, gidStart uint64, blks []Label, nkeys, h int) error {
	return lastError(C.gcb_mitccrh_hash((*C.gcb_label)(unsafe.Pointer(seed)), C.uint64_t(gidStart), lp(blks),
		C.uint64_t(nkeys), C.uint32_t(h)))
}

// COT / ROT post-processing (ot/cot.go:157-233, ot/rot.go:155-197): OT j is hashed under MiTCCRH key j.
// wireBytes: msgs is the SendLabel byte encoding, ready for conn.Write / filled by io.ReadFull.
func COTSend(seed, delta *Label, q []Label, wires []Wire, msgs []byte) error {
	return lastError(C.gcb_cot_send((*C.gcb_label)(unsafe.Pointer(seed)), (*C.gcb_label)(unsafe.Pointer(delta)), lp(q), wp(wires),
		C.uint64_t(len(q)), (*C.gcb_label)(unsafe.Pointer(&msgs[0])), C.GCB_COT_WIRE_BYTES))
}
func COTReceive(seed *Label, flags []bool, msgs []byte, result []Label) error {
	return lastError(C.gcb_cot_receive((*C.gcb_label)(unsafe.Pointer(seed)), (*C.uint8_t)(unsafe.Pointer(&flags[0])),
		(*C.gcb_label)(unsafe.Pointer(&msgs[0])), lp(result), C.uint64_t(len(result)), lp(result), C.GCB_COT_WIRE_BYTES))
}
func ROTSend(seed, delta *Label, q []Label, wires []Wire) error {
	return lastError(C.gcb_rot_send((*C.gcb_label)(unsafe.Pointer(seed)), (*C.gcb_label)(unsafe.Pointer(delta)), lp(q),
		C.uint64_t(len(q)), wp(wires)))
}
func ROTReceive(seed *Label, result []Label) error {
	return lastError(C.gcb_rot_receive((*C.gcb_label)(unsafe.Pointer(seed)), lp(result), C.uint64_t(len(result)), lp(result)))
}

// IKNPCheckSums replaces the chi loops of the malicious-mode check (ot/iknp.go:150-173, :408-451):
// out = {lo, hi, x}; choice is nil on the sender side.
func IKNPCheckSums(seed2 *Label, chiStart uint64, labels []Label, choice []bool) (out [3]Label, err error) {
	var cp *C.uint8_t
	if choice != nil {
		cp = (*C.uint8_t)(unsafe.Pointer(&choice[0]))
	}
	err = lastError(C.gcb_iknp_check_sums((*C.gcb_label)(unsafe.Pointer(seed2)), C.uint64_t(chiStart), lp(labels), cp,
		C.uint64_t(len(labels)), (*C.gcb_label)(unsafe.Pointer(&out[0]))))
	return
}

// TablesToWire / TablesFromWire convert a batch of row slabs to and from the byte stream of
// circuit/garbler.go:69-82 / circuit/evaluator.go:40-66 (stride = bytes per instance in wire).
func (p *Plan) TablesWireSize() int {
	var n C.size_t
	C.gcb_tables_wire_size(p.h, &n)
	return int(n)
}
func (p *Plan) TablesToWire(batch int, tables []Label, wire []byte, stride int) error {
	return lastError(C.gcb_tables_to_wire(p.h, C.uint32_t(batch), lp(tables), (*C.uint8_t)(unsafe.Pointer(&wire[0])), C.size_t(stride)))
}
func (p *Plan) TablesFromWire(batch int, wire []byte, stride int, tables []Label) error {
	return lastError(C.gcb_tables_from_wire(p.h, C.uint32_t(batch), (*C.uint8_t)(unsafe.Pointer(&wire[0])), C.size_t(stride), lp(tables)))
}

// GarbleBegin / GarbleWait split Stream.Garble so that one step stays in flight: the kernel of step k+1
// runs while the record bytes of step k cross PCIe.  dst must be page-locked C memory from HostAlloc (the copy
// outlives the call, so Go-managed memory is not allowed here by the cgo pointer rules) and must stay untouched
// until a GarbleWait covers it.
func (s *Stream) GarbleBegin(p *Plan, in, out []uint32, dst []byte) (n int, err error) {
	var w C.size_t
	err = lastError(C.gcb_stream_garble_begin(s.h, p.h, u32p(in), C.uint32_t(len(in)), u32p(out), C.uint32_t(len(out)),
		(*C.uint8_t)(unsafe.Pointer(&dst[0])), C.size_t(len(dst)), &w))
	return int(w), err
}

// GarbleWait(0): every begun step is complete; GarbleWait(1): all but the one begun last.
func (s *Stream) GarbleWait(leaveInFlight int) error {
	return lastError(C.gcb_stream_garble_wait(s.h, C.uint32_t(leaveInFlight)))
}

// Device buffers and streams (the operands of the _dev entry points) for callers without a CUDA binding.
func DevAlloc(n int) unsafe.Pointer { return C.gcb_dev_alloc(C.size_t(n)) }
func DevFree(p unsafe.Pointer)      { C.gcb_dev_free(p) }
func DevUpload(dst unsafe.Pointer, src []byte, stream unsafe.Pointer) error {
	return lastError(C.gcb_dev_upload(dst, unsafe.Pointer(&src[0]), C.size_t(len(src)), stream))
}
func DevDownload(dst []byte, src unsafe.Pointer, stream unsafe.Pointer) error {
	return lastError(C.gcb_dev_download(unsafe.Pointer(&dst[0]), src, C.size_t(len(dst)), stream))
}
func DevSync(stream unsafe.Pointer) error { return lastError(C.gcb_dev_sync(stream)) }
