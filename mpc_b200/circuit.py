"""Host-side mirror of the reference's ``circuit`` package API for the hot path.

``GarbleEngine`` wraps one ``circuit.Circuit`` (mpc_b200.circuit_io.Circuit)
and exposes, with the reference's names and argument meaning:

* ``garble(rand, key) -> Garbled``     -- (*Circuit).Garble, circuit/garble.go:248
* ``eval(key, wires, garbled)``        -- (*Circuit).Eval,   circuit/eval.go:17
* ``garble_batch`` / ``eval_batch``    -- the batched throughput entry points
  (host numpy buffers through gcb_garble / gcb_eval, or device tensors
  through gcb_garble_dev / gcb_eval_dev).

All compute happens in libgcb200.so (CUDA, sm_100a).  Nothing here imports
the CPU oracle; without the library or a GPU the calls raise.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np

from . import _lib
from ._lib import GcbError, PlanInfo, check, ptr
from .circuit_io import AND, INV, LABEL_DTYPE, OR, WIRE_DTYPE, Circuit

__all__ = ["Garbled", "GarbleEngine", "Streaming", "StreamEval", "GcbError", "Job", "set_devices", "get_devices",
           "host_alloc", "host_free"]


def set_devices(ids) -> None:
    """gcb_set_devices: the devices host-pointer calls fan out over (empty = back to the default)."""
    arr = (C.c_int * len(ids))(*ids)
    check(_lib.lib().gcb_set_devices(arr, len(ids)))


def get_devices() -> List[int]:
    arr = (C.c_int * 64)()
    n = _lib.lib().gcb_get_devices(arr, 64)
    return [int(arr[i]) for i in range(min(n, 64))]


def host_alloc(shape, dtype) -> np.ndarray:
    """Page-locked array from gcb_host_alloc (what the Go side's pooled slabs become); free with host_free."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = _lib.lib().gcb_host_alloc(max(n, 1))
    if not p:
        raise GcbError(_lib.E_CUDA, _lib.lib().gcb_last_error().decode())
    arr = np.frombuffer((C.c_uint8 * max(n, 1)).from_address(p), dtype=np.uint8)[:n].view(dtype).reshape(shape)
    _host_blocks[arr.ctypes.data] = p
    return arr


_host_blocks = {}


def host_free(arr: np.ndarray) -> None:
    p = _host_blocks.pop(arr.ctypes.data, None)
    if p:
        _lib.lib().gcb_host_free(p)


class HostCircuit:
    """The C++ circuit front end (gcb_circuit_*) over a parsed circuit: Circuit.Compute on wire bits
    (circuit/computer.go:15-91) for whole batches, on the host."""

    def __init__(self, circ: Circuit):
        self.circ = circ
        gates = np.ascontiguousarray(circ.gates)
        ins = np.asarray(circ.inputs, dtype=np.uint32)
        outs = np.asarray(circ.outputs, dtype=np.uint32)
        h = C.c_void_p()
        check(_lib.lib().gcb_circuit_from_gates(ptr(gates), circ.num_gates, circ.num_wires, ptr(ins), len(ins),
                                                ptr(outs), len(outs), C.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().gcb_circuit_destroy(h)
            except Exception:
                pass

    def compute_bits(self, in_bits: np.ndarray) -> np.ndarray:
        """uint8[batch, num_inputs] of 0/1 -> uint8[batch, num_outputs]."""
        b = np.ascontiguousarray(in_bits, dtype=np.uint8).reshape(-1, self.circ.num_inputs)
        out = np.zeros((len(b), self.circ.num_outputs), dtype=np.uint8)
        check(_lib.lib().gcb_circuit_compute(self._h, len(b), ptr(b), ptr(out)))
        return out


class Job:
    """A queued gcb_garble_begin / gcb_eval_begin call.  ``wait()`` blocks until the results are in the caller's
    buffers; the arrays handed to the begin call are kept alive until then."""

    def __init__(self, handle, keep):
        self._h, self._keep = handle, keep

    def done(self) -> bool:
        return self._h is None or bool(_lib.lib().gcb_job_done(self._h))

    def wait(self) -> None:
        h, self._h = self._h, None
        if h:
            try:
                check(_lib.lib().gcb_job_wait(h))
            finally:
                self._keep = None

    def __del__(self):
        try:
            self.wait()
        except Exception:
            pass


def _read(rand, n: int) -> bytes:
    """io.Reader.Read(n) on bytes-like or file-like sources."""
    if hasattr(rand, "read"):
        b = rand.read(n)
        if len(b) != n:
            raise EOFError("rand: short read")     # Garble returns the reader's error
        return b
    raise TypeError("rand must provide read(n)")


class _BytesReader:
    def __init__(self, b: bytes):
        self._b, self._pos = bytes(b), 0

    def read(self, n: int) -> bytes:
        out = self._b[self._pos:self._pos + n]
        self._pos += n
        return out


class Garbled:
    """circuit.Garbled (circuit/garble.go:162-168): R, Wires, Gates."""

    def __init__(self, r, wires, slab, row_off):
        self.R = r
        self.Wires = wires                     # WIRE_DTYPE[num_wires]
        self.slab = slab                       # LABEL_DTYPE[num_rows]
        self._row_off = row_off
        self._gates: Optional[List] = None

    @property
    def Gates(self) -> List[Optional[np.ndarray]]:
        """[][]ot.Label: slab sub-slices; None for XOR/XNOR (garble.go:292-298)."""
        if self._gates is None:
            ro = self._row_off
            self._gates = [self.slab[ro[i]:ro[i + 1]] if ro[i + 1] > ro[i] else None
                           for i in range(len(ro) - 1)]
        return self._gates


def _key_args(keys, batch: int):
    """bytes (one shared key) or uint8 [batch, keylen] (host) -> (array, keylen, stride)."""
    if isinstance(keys, (bytes, bytearray)):
        return _lib.u8(keys), len(keys), 0
    keys = np.ascontiguousarray(keys, dtype=np.uint8)
    if keys.ndim != 2 or keys.shape[0] != batch:
        raise ValueError("keys must be bytes or uint8[batch, keylen]")
    return keys, keys.shape[1], keys.shape[1]


FLAG_FANOUT = 1


class GarbleEngine:
    def __init__(self, circ: Circuit):
        self.circ = circ
        gates = np.ascontiguousarray(circ.gates)
        h = C.c_void_p()
        check(_lib.lib().gcb_plan_create(ptr(gates), circ.num_gates, circ.num_wires, circ.num_inputs,
                                         circ.num_outputs, C.byref(h)))
        self._h = h
        info = PlanInfo()
        check(_lib.lib().gcb_plan_get_info(self._h, C.byref(info)))
        self.info = info
        self.row_off = np.zeros(circ.num_gates + 1, dtype=np.uint32)
        check(_lib.lib().gcb_plan_row_offsets(self._h, ptr(self.row_off)))

    def info_for(self, batch: int) -> PlanInfo:
        """Geometry of the plan a call with ``batch`` instances per device runs on (gcb_plan_get_info_for_batch)."""
        info = PlanInfo()
        check(_lib.lib().gcb_plan_get_info_for_batch(self._h, int(batch), C.byref(info)))
        return info

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().gcb_plan_destroy(h)
            except Exception:
                pass

    @property
    def handle(self):
        return self._h

    # ---- the reference's single-instance API ----------------------------------
    def garble(self, rand, key: bytes) -> Garbled:
        """(*Circuit).Garble(rand, key).  Reads 16 bytes for R, then 16 per input
        wire, from ``rand`` in that order (garble.go:253-278)."""
        if isinstance(rand, (bytes, bytearray)):
            rand = _BytesReader(rand)
        c = self.circ
        r = np.frombuffer(_read(rand, 16), dtype=">u8").astype("<u8").view(LABEL_DTYPE).copy()
        l0 = np.zeros(c.num_inputs, dtype=LABEL_DTYPE)
        if c.num_inputs:
            raw = np.frombuffer(_read(rand, 16 * c.num_inputs), dtype=">u8").astype("<u8")
            l0 = raw.view(LABEL_DTYPE).copy()
        slab = np.zeros(max(c.num_rows, 1), dtype=LABEL_DTYPE)
        wires = np.zeros(c.num_wires, dtype=WIRE_DTYPE)
        k = _lib.u8(key)
        check(_lib.lib().gcb_garble(self._h, ptr(k), len(key), 0, 1, ptr(r), ptr(l0), ptr(slab), None,
                                    ptr(wires), 0))
        r["d0"] |= np.uint64(1) << np.uint64(63)          # r.SetS(true)
        return Garbled(r[0], wires, slab[: c.num_rows], self.row_off)

    def eval(self, key: bytes, wires: np.ndarray, garbled) -> None:
        """(*Circuit).Eval(key, wires, garbled): in place on ``wires``
        (LABEL_DTYPE[num_wires], inputs pre-filled).  ``garbled`` is the
        [][]ot.Label list (or a Garbled / a slab array).  Row-count errors are
        the reference's "corrupted circuit" errors (eval.go:55,87,103)."""
        c = self.circ
        if isinstance(garbled, Garbled):
            slab = garbled.slab
        elif isinstance(garbled, np.ndarray) and garbled.dtype == LABEL_DTYPE:
            slab = garbled
        else:
            need = {AND: 2, OR: 3, INV: 1}
            if len(garbled) != c.num_gates:
                raise GcbError(_lib.E_CORRUPT, "corrupted circuit: gate count")
            parts = []
            for i, op in enumerate(c.gates["op"].tolist()):
                rows = garbled[i]
                n = 0 if rows is None else len(rows)
                if op == AND and n != 2:
                    raise GcbError(_lib.E_CORRUPT, f"corrupted ciruit: AND row length: {n}")
                if op in (OR, INV) and n < need[op]:
                    raise GcbError(_lib.E_CORRUPT, f"corrupted circuit: index {need[op] - 1} >= row {n}")
                if op in need:
                    parts.append(np.asarray(rows[: need[op]], dtype=LABEL_DTYPE))
            slab = np.concatenate(parts) if parts else np.zeros(0, dtype=LABEL_DTYPE)
        if len(slab) < c.num_rows:
            raise GcbError(_lib.E_CORRUPT, "corrupted circuit: table too short")
        if wires.dtype != LABEL_DTYPE or len(wires) != c.num_wires or not wires.flags["C_CONTIGUOUS"]:
            raise ValueError("wires must be a contiguous LABEL_DTYPE[num_wires] array")
        slab = np.ascontiguousarray(slab if len(slab) else np.zeros(1, dtype=LABEL_DTYPE))
        inl = np.ascontiguousarray(wires[: c.num_inputs]) if c.num_inputs else np.zeros(1, LABEL_DTYPE)
        out = np.zeros(max(c.num_outputs, 1), dtype=LABEL_DTYPE)
        full = np.zeros(c.num_wires, dtype=LABEL_DTYPE)
        k = _lib.u8(key)
        check(_lib.lib().gcb_eval(self._h, ptr(k), len(key), 0, 1, ptr(slab), ptr(inl), ptr(out), ptr(full), 0))
        wires[:] = full

    # ---- batched, host buffers -------------------------------------------------
    def garble_batch(self, keys, r: np.ndarray, in_l0: np.ndarray, tables: Optional[np.ndarray] = None,
                     io_wires: Optional[np.ndarray] = None, want_io: bool = True):
        """r: LABEL[batch] raw R draws; in_l0: LABEL[batch, ninputs].
        Returns (tables LABEL[batch, rows], io_wires WIRE[batch, nin+nout])."""
        c = self.circ
        batch = len(r)
        ka, kl, ks = _key_args(keys, batch)
        r = np.ascontiguousarray(r, dtype=LABEL_DTYPE)
        in_l0 = np.ascontiguousarray(in_l0, dtype=LABEL_DTYPE)
        if tables is None:
            tables = np.zeros((batch, max(c.num_rows, 1)), dtype=LABEL_DTYPE)[:, : c.num_rows]
            tables = np.ascontiguousarray(tables) if c.num_rows else np.zeros((batch, 0), dtype=LABEL_DTYPE)
        if io_wires is None and want_io:
            io_wires = np.zeros((batch, c.num_inputs + c.num_outputs), dtype=WIRE_DTYPE)
        tp = ptr(tables) if c.num_rows else ptr(np.zeros(1, LABEL_DTYPE))
        check(_lib.lib().gcb_garble(self._h, ptr(ka), kl, ks, batch, ptr(r), ptr(in_l0), tp,
                                    ptr(io_wires) if io_wires is not None else None, None, 0))
        return tables, io_wires

    def eval_batch(self, keys, tables: np.ndarray, in_labels: np.ndarray, out_labels: Optional[np.ndarray] = None):
        c = self.circ
        batch = len(in_labels)
        ka, kl, ks = _key_args(keys, batch)
        in_labels = np.ascontiguousarray(in_labels, dtype=LABEL_DTYPE)
        if out_labels is None:
            out_labels = np.zeros((batch, c.num_outputs), dtype=LABEL_DTYPE)
        tp = ptr(np.ascontiguousarray(tables)) if c.num_rows else ptr(np.zeros(1, LABEL_DTYPE))
        check(_lib.lib().gcb_eval(self._h, ptr(ka), kl, ks, batch, tp, ptr(in_labels), ptr(out_labels), None, 0))
        return out_labels

    # ---- the same, split in two (one host thread keeps every selected device and both PCIe directions busy) ----
    def garble_begin(self, keys, r: np.ndarray, in_l0: np.ndarray, tables: np.ndarray,
                     io_wires: Optional[np.ndarray] = None) -> Job:
        """Queue a batched Garble; buffers must be C-contiguous and stay untouched until ``Job.wait()``."""
        c = self.circ
        batch = len(r)
        ka, kl, ks = _key_args(keys, batch)
        for a in (r, in_l0, tables, io_wires):
            assert a is None or a.flags["C_CONTIGUOUS"], "begin calls take contiguous arrays (no hidden copies)"
        h = C.c_void_p()
        check(_lib.lib().gcb_garble_begin(self._h, ptr(ka), kl, ks, batch, ptr(r), ptr(in_l0) if c.num_inputs else None,
                                          ptr(tables) if c.num_rows else ptr(np.zeros(1, LABEL_DTYPE)),
                                          ptr(io_wires) if io_wires is not None else None, None, 0, C.byref(h)))
        return Job(h, (ka, r, in_l0, tables, io_wires))

    def eval_begin(self, keys, tables: np.ndarray, in_labels: np.ndarray, out_labels: np.ndarray) -> Job:
        c = self.circ
        batch = len(in_labels)
        ka, kl, ks = _key_args(keys, batch)
        for a in (tables, in_labels, out_labels):
            assert a.flags["C_CONTIGUOUS"], "begin calls take contiguous arrays (no hidden copies)"
        h = C.c_void_p()
        check(_lib.lib().gcb_eval_begin(self._h, ptr(ka), kl, ks, batch,
                                        ptr(tables) if c.num_rows else ptr(np.zeros(1, LABEL_DTYPE)), ptr(in_labels),
                                        ptr(out_labels), None, 0, C.byref(h)))
        return Job(h, (ka, tables, in_labels, out_labels))

    # ---- the garbled tables as Garbler sends them (circuit/garbler.go:69-82) --------
    def tables_wire_size(self) -> int:
        import ctypes as C
        n = C.c_size_t()
        check(_lib.lib().gcb_tables_wire_size(self._h, C.byref(n)))
        return int(n.value)

    def tables_to_wire(self, tables: np.ndarray) -> np.ndarray:
        """[batch][rows] labels -> uint8[batch][wire_size]: u32 NumGates, then per gate u32 count + BE rows."""
        tables = np.ascontiguousarray(tables, dtype=LABEL_DTYPE).reshape(-1, max(self.circ.num_rows, 1))
        batch, n = len(tables), self.tables_wire_size()
        out = np.zeros((batch, n), dtype=np.uint8)
        check(_lib.lib().gcb_tables_to_wire(self._h, batch, ptr(tables), ptr(out), n))
        return out

    def tables_from_wire(self, wire: np.ndarray) -> np.ndarray:
        """The inverse (circuit/evaluator.go:40-66); raises GcbError on a count that does not fit the circuit."""
        wire = np.ascontiguousarray(wire, dtype=np.uint8)
        batch, stride = wire.shape
        tables = np.zeros((batch, max(self.circ.num_rows, 1)), dtype=LABEL_DTYPE)
        check(_lib.lib().gcb_tables_from_wire(self._h, batch, ptr(wire), stride, ptr(tables)))
        return tables[:, : self.circ.num_rows]

    # ---- batched, device-resident (torch tensors or raw device addresses) -------
    def garble_dev(self, keys_dev, keylen: int, key_stride: int, batch: int, r_dev, in_l0_dev, tables_dev,
                   io_wires_dev=None, wires_full_dev=None, stream: int = 0, flags: int = 0) -> None:
        """flags = FLAG_FANOUT: split the batch over the set_devices list (operands stay on this device)."""
        check(_lib.lib().gcb_garble_dev(self._h, ptr(keys_dev), keylen, key_stride, batch, ptr(r_dev),
                                        ptr(in_l0_dev), ptr(tables_dev), ptr(io_wires_dev),
                                        ptr(wires_full_dev), flags, stream))

    def eval_dev(self, keys_dev, keylen: int, key_stride: int, batch: int, tables_dev, in_labels_dev,
                 out_labels_dev, wires_full_dev=None, stream: int = 0, flags: int = 0) -> None:
        check(_lib.lib().gcb_eval_dev(self._h, ptr(keys_dev), keylen, key_stride, batch, ptr(tables_dev),
                                      ptr(in_labels_dev), ptr(out_labels_dev), ptr(wires_full_dev), flags, stream))


def select_labels_dev(wires_dev, wire_stride: int, bits_dev, out_dev, batch: int, n: int, stream: int = 0):
    """LabelForBit over a batch (circuit/helpers.go:10-16) on the device."""
    check(_lib.lib().gcb_select_labels_dev(ptr(wires_dev), wire_stride, ptr(bits_dev), ptr(out_dev), batch, n, stream))


def decode_bits_dev(wires_dev, wire_stride: int, labels_dev, bits_dev, batch: int, n: int, stream: int = 0):
    """BitFromLabel over a batch (circuit/helpers.go:18-27) on the device."""
    check(_lib.lib().gcb_decode_bits_dev(ptr(wires_dev), wire_stride, ptr(labels_dev), ptr(bits_dev), batch, n, stream))


def hash_half(key: bytes, x: np.ndarray, tweak0: int = 0) -> np.ndarray:
    """encryptHalf over n labels with tweaks tweak0+i (circuit/garble.go:104-136)."""
    x = np.ascontiguousarray(x, dtype=LABEL_DTYPE)
    out = np.zeros(len(x), dtype=LABEL_DTYPE)
    k = _lib.u8(key)
    check(_lib.lib().gcb_hash_half(ptr(k), len(key), ptr(x), tweak0, ptr(out), len(x)))
    return out


class Streaming:
    """circuit.Streaming (circuit/stream_garble.go:27-191) for ``batch``
    program instances garbled in lock step; batch = 1 is the Go object.

    ``r``: LABEL[batch] raw R draws (S forced inside, :44-50); ``input_ids``:
    permanent wire ids; ``in_l0``: LABEL[batch, len(input_ids)] (:56-73)."""

    def __init__(self, keys, r: np.ndarray, input_ids, in_l0: np.ndarray):
        r = np.ascontiguousarray(r, dtype=LABEL_DTYPE).reshape(-1)
        self.batch = len(r)
        ka, kl, ks = _key_args(keys, self.batch)
        ids = np.ascontiguousarray(input_ids, dtype=np.uint32)
        in_l0 = np.ascontiguousarray(in_l0, dtype=LABEL_DTYPE).reshape(self.batch, len(ids))
        h = C.c_void_p()
        check(_lib.lib().gcb_stream_create(ptr(ka), kl, ks, self.batch, ptr(r), ptr(ids) if len(ids) else None,
                                           len(ids), ptr(in_l0) if len(ids) else None, C.byref(h)))
        self._h = h

    @classmethod
    def new(cls, rand, key: bytes, inputs) -> "Streaming":
        """NewStreaming(cfg, key, inputs, conn): R then one L0 per input from ``rand``."""
        if isinstance(rand, (bytes, bytearray)):
            rand = _BytesReader(rand)
        ids = list(inputs)
        raw = _read(rand, 16 * (1 + len(ids)))
        lab = np.frombuffer(raw, dtype=">u8").astype("<u8").view(LABEL_DTYPE)
        return cls(key, lab[:1].copy(), ids, lab[1:].copy().reshape(1, len(ids)))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().gcb_stream_destroy(h)
                for slot in self.__dict__.get("_ring", []):
                    if slot[1]:
                        _lib.lib().gcb_host_free(slot[1])
                        slot[1] = None
            except Exception:
                pass

    #: page-locked output buffers used in turn (the Go side's conn.WriteBuf role): with 2 or more, the bytes of
    #: step k stay valid while step k+1 is garbled, so a consumer (socket writer, evaluator) can run concurrently
    stream_buffers = 1

    def _stream_buffer(self, n: int) -> np.ndarray:
        """Page-locked [batch, n] byte buffer from a ring of ``stream_buffers`` buffers kept across steps.
        The returned array is overwritten ``stream_buffers`` garble() calls later."""
        need = self.batch * n
        ring = self.__dict__.setdefault("_ring", [])
        turn = self.__dict__.get("_turn", 0)
        self._turn = turn + 1
        k = turn % max(1, int(self.stream_buffers))
        while len(ring) <= k:
            ring.append([0, None])                       # [capacity, address]
        cap, addr = ring[k]
        if cap < need:
            if addr:
                _lib.lib().gcb_host_free(addr)
            cap = need + need // 4
            addr = _lib.lib().gcb_host_alloc(cap)
            if not addr:
                raise GcbError(_lib.E_CUDA, _lib.lib().gcb_last_error().decode())
            ring[k] = [cap, addr]
        raw = (C.c_uint8 * need).from_address(addr)
        return np.frombuffer(raw, dtype=np.uint8).reshape(self.batch, n)

    def get_inputs(self, ids) -> np.ndarray:
        """GetInput / GetInputs (:117-128): WIRE[batch, n]."""
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        out = np.zeros((self.batch, len(ids)), dtype=WIRE_DTYPE)
        check(_lib.lib().gcb_stream_get_wires(self._h, ptr(ids) if len(ids) else None, len(ids),
                                              ptr(out) if len(ids) else None))
        return out

    def step_size(self, eng: GarbleEngine, in_ids, out_ids) -> int:
        i = np.ascontiguousarray(in_ids, dtype=np.uint32)
        o = np.ascontiguousarray(out_ids, dtype=np.uint32)
        n = C.c_size_t()
        check(_lib.lib().gcb_stream_step_size(self._h, eng.handle, ptr(i) if len(i) else None, len(i),
                                              ptr(o) if len(o) else None, len(o), C.byref(n)))
        return int(n.value)

    def garble(self, eng: GarbleEngine, in_ids, out_ids):
        """Streaming.Garble(c, in, out) (:161-191).  Returns (stream bytes
        uint8[batch, n], t_init_ns, t_garble_ns): per instance exactly what the
        reference writes into conn.WriteBuf for this sub-circuit."""
        i = np.ascontiguousarray(in_ids, dtype=np.uint32)
        o = np.ascontiguousarray(out_ids, dtype=np.uint32)
        # upper bound of the record stream (1 + 3 * 4 header bytes per gate, 16 per row): no second
        # pass over the gates just to size the buffer (step_size gives the exact figure when needed)
        n = 13 * eng.circ.num_gates + 16 * eng.circ.num_rows
        buf = self._stream_buffer(max(n, 1))
        w, t0, t1 = C.c_size_t(), C.c_uint64(), C.c_uint64()
        check(_lib.lib().gcb_stream_garble(self._h, eng.handle, ptr(i) if len(i) else None, len(i),
                                           ptr(o) if len(o) else None, len(o), ptr(buf), buf.shape[1],
                                           C.byref(w), C.byref(t0), C.byref(t1)))
        return buf[:, : int(w.value)], int(t0.value), int(t1.value)


    def garble_begin(self, eng: GarbleEngine, in_ids, out_ids) -> np.ndarray:
        """Queue one Streaming.Garble step and return its output buffer, whose bytes are valid only after a
        garble_wait() that covers it.  Set ``stream_buffers`` >= 2 (the buffers are handed out in turn)."""
        i = np.ascontiguousarray(in_ids, dtype=np.uint32)
        o = np.ascontiguousarray(out_ids, dtype=np.uint32)
        n = 13 * eng.circ.num_gates + 16 * eng.circ.num_rows
        buf = self._stream_buffer(max(n, 1))
        w = C.c_size_t()
        check(_lib.lib().gcb_stream_garble_begin(self._h, eng.handle, ptr(i) if len(i) else None, len(i),
                                                 ptr(o) if len(o) else None, len(o), ptr(buf), buf.shape[1], C.byref(w)))
        return buf[:, : int(w.value)]

    def garble_wait(self, leave_in_flight: int = 0) -> None:
        """0: every begun step's bytes are in place; 1: all but the step begun last."""
        check(_lib.lib().gcb_stream_garble_wait(self._h, leave_in_flight))


class StreamEval:
    """circuit.StreamEval + the OpCircuit gate loop of StreamEvaluator
    (circuit/stream_evaluator.go:29-96, 270-432) for ``batch`` instances."""

    def __init__(self, keys, batch: int = 1):
        ka, kl, ks = _key_args(keys, batch)
        self.batch = batch
        h = C.c_void_p()
        check(_lib.lib().gcb_seval_create(ptr(ka), kl, ks, batch, C.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().gcb_seval_destroy(h)
            except Exception:
                pass

    def set(self, ids, labels: np.ndarray) -> None:
        """Set / SetInputs: labels LABEL[batch, n]."""
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        labels = np.ascontiguousarray(labels, dtype=LABEL_DTYPE).reshape(self.batch, len(ids))
        check(_lib.lib().gcb_seval_set_wires(self._h, ptr(ids) if len(ids) else None, len(ids),
                                             ptr(labels) if len(ids) else None))

    def get(self, ids) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        out = np.zeros((self.batch, len(ids)), dtype=LABEL_DTYPE)
        check(_lib.lib().gcb_seval_get_wires(self._h, ptr(ids) if len(ids) else None, len(ids),
                                             ptr(out) if len(ids) else None))
        return out

    def circuit(self, stream: np.ndarray, ngates: int, ntmp: int, nwires: int) -> int:
        """Evaluate one OpCircuit body.  stream: uint8[batch, n] record bytes; returns bytes consumed."""
        if not (isinstance(stream, np.ndarray) and stream.dtype == np.uint8 and stream.ndim == 2 and
                len(stream) == self.batch and stream.strides[1] == 1):
            stream = np.ascontiguousarray(stream, dtype=np.uint8).reshape(self.batch, -1)
        used = C.c_size_t()               # rows of a wider buffer are passed by stride, not copied
        check(_lib.lib().gcb_seval_circuit(self._h, stream.ctypes.data, stream.strides[0], stream.shape[1], ngates, ntmp,
                                           nwires, C.byref(used)))
        return int(used.value)
