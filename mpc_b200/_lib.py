"""ctypes loader of libgcb200.so, the C-ABI library (include/gcb200.h).

There is no fallback: if the library cannot be built or loaded, importing the
compute modules raises; if no CUDA device is present every compute call raises
``GcbError`` with status GCB_E_CUDA.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

# gcb_status (include/gcb200.h)
OK, E_ARG, E_KEYLEN, E_BADOP, E_WIRE, E_CUDA, E_TOO_LARGE, E_BUFFER, E_CHUNK, E_CORRUPT = (
    0, -1, -2, -3, -4, -5, -6, -7, -8, -9)


class GcbError(RuntimeError):
    def __init__(self, rc: int, msg: str):
        super().__init__(msg or f"gcb200 error {rc}")
        self.rc = rc


class PlanInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "num_gates", "num_wires", "num_inputs", "num_outputs", "num_rows", "num_tweaks", "num_steps",
        "num_slots", "num_and", "num_or", "num_inv", "num_free", "teams_per_sm", "team_threads",
        "garble_hashes", "eval_hashes", "garble_passes", "eval_passes", "num_hot_slots")]


class Label(C.Structure):
    _fields_ = [("d0", C.c_uint64), ("d1", C.c_uint64)]


# every symbol include/gcb200.h declares: name -> (restype, argtypes)
_vp, _u32, _u64, _sz, _int = C.c_void_p, C.c_uint32, C.c_uint64, C.c_size_t, C.c_int
SYMBOLS = {
    "gcb_last_error": (C.c_char_p, []),
    "gcb_version": (C.c_char_p, []),
    "gcb_launch_count": (_u64, []),
    "gcb_set_device": (_int, [_int]),
    "gcb_set_devices": (_int, [_vp, _int]),
    "gcb_get_devices": (_int, [_vp, _int]),
    "gcb_device_count": (_int, []),
    "gcb_host_alloc": (_vp, [_sz]),
    "gcb_host_free": (None, [_vp]),
    "gcb_dev_alloc": (_vp, [_sz]),
    "gcb_dev_free": (None, [_vp]),
    "gcb_dev_upload": (_int, [_vp, _vp, _sz, _vp]),
    "gcb_dev_download": (_int, [_vp, _vp, _sz, _vp]),
    "gcb_dev_stream_create": (_int, [C.POINTER(_vp)]),
    "gcb_dev_stream_destroy": (None, [_vp]),
    "gcb_dev_sync": (_int, [_vp]),
    "gcb_plan_create": (_int, [_vp, _u32, _u32, _u32, _u32, C.POINTER(_vp)]),
    "gcb_plan_destroy": (None, [_vp]),
    "gcb_plan_get_info": (_int, [_vp, C.POINTER(PlanInfo)]),
    "gcb_plan_get_info_for_batch": (_int, [_vp, C.c_uint64, C.POINTER(PlanInfo)]),
    "gcb_plan_row_offsets": (_int, [_vp, _vp]),
    "gcb_garble": (_int, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _u32]),
    "gcb_eval": (_int, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _u32]),
    "gcb_garble_begin": (_int, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _u32, C.POINTER(_vp)]),
    "gcb_eval_begin": (_int, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _u32, C.POINTER(_vp)]),
    "gcb_job_wait": (_int, [_vp]),
    "gcb_job_done": (_int, [_vp]),
    "gcb_garble_dev": (_int, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _u32, _vp]),
    "gcb_eval_dev": (_int, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _u32, _vp]),
    "gcb_select_labels_dev": (_int, [_vp, _sz, _vp, _vp, _u32, _u32, _vp]),
    "gcb_decode_bits_dev": (_int, [_vp, _sz, _vp, _vp, _u32, _u32, _vp]),
    "gcb_hash_half": (_int, [_vp, _u32, _vp, _u32, _vp, _u64]),
    "gcb_hash_half_dev": (_int, [_vp, _u32, _vp, _u32, _vp, _u64, _vp]),
    "gcb_stream_create": (_int, [_vp, _u32, _u32, _u32, _vp, _vp, _u32, _vp, C.POINTER(_vp)]),
    "gcb_stream_destroy": (None, [_vp]),
    "gcb_stream_get_wires": (_int, [_vp, _vp, _u32, _vp]),
    "gcb_stream_step_size": (_int, [_vp, _vp, _vp, _u32, _vp, _u32, C.POINTER(_sz)]),
    "gcb_stream_garble": (_int, [_vp, _vp, _vp, _u32, _vp, _u32, _vp, _sz, C.POINTER(_sz),
                                 C.POINTER(_u64), C.POINTER(_u64)]),
    "gcb_stream_garble_begin": (_int, [_vp, _vp, _vp, _u32, _vp, _u32, _vp, _sz, C.POINTER(_sz)]),
    "gcb_stream_garble_wait": (_int, [_vp, _u32]),
    "gcb_seval_create": (_int, [_vp, _u32, _u32, _u32, C.POINTER(_vp)]),
    "gcb_seval_destroy": (None, [_vp]),
    "gcb_seval_set_wires": (_int, [_vp, _vp, _u32, _vp]),
    "gcb_seval_get_wires": (_int, [_vp, _vp, _u32, _vp]),
    "gcb_seval_circuit": (_int, [_vp, _vp, _sz, _sz, _u32, _u32, _u32, C.POINTER(_sz)]),
    "gcb_iknp_u_size": (_sz, [_u64]),
    "gcb_iknp_stream_advance": (_u64, [_u64]),
    "gcb_iknp_receiver_expand": (_int, [_vp, _vp, _u64, _vp, _u64, _vp, _vp]),
    "gcb_iknp_sender_expand": (_int, [_vp, _vp, _u64, _vp, _sz, _u64, _vp]),
    "gcb_iknp_receiver_expand_dev": (_int, [_vp, _vp, _u64, _vp, _u64, _vp, _vp, _vp]),
    "gcb_iknp_sender_expand_dev": (_int, [_vp, _vp, _u64, _vp, _sz, _u64, _vp, _vp]),
    "gcb_iknp_receiver_expand_bits": (_int, [_vp, _vp, _u64, _vp, _u64, _vp, _vp]),
    "gcb_iknp_sender_expand_bits": (_int, [_vp, _vp, _u64, _vp, _sz, _u64, _vp]),
    "gcb_iknp_receiver_expand_bits_dev": (_int, [_vp, _vp, _u64, _vp, _u64, _vp, _vp, _vp]),
    "gcb_iknp_sender_expand_bits_dev": (_int, [_vp, _vp, _u64, _vp, _sz, _u64, _vp, _vp]),
    "gcb_mitccrh_hash": (_int, [C.POINTER(Label), _u64, _vp, _u64, _u32]),
    "gcb_mitccrh_hash_dev": (_int, [C.POINTER(Label), _u64, _vp, _u64, _u32, _vp]),
    "gcb_cot_send": (_int, [C.POINTER(Label), C.POINTER(Label), _vp, _vp, _u64, _vp, _u32]),
    "gcb_cot_receive": (_int, [C.POINTER(Label), _vp, _vp, _vp, _u64, _vp, _u32]),
    "gcb_rot_send": (_int, [C.POINTER(Label), C.POINTER(Label), _vp, _u64, _vp]),
    "gcb_rot_receive": (_int, [C.POINTER(Label), _vp, _u64, _vp]),
    "gcb_cot_send_dev": (_int, [C.POINTER(Label), C.POINTER(Label), _vp, _vp, _u64, _vp, _u32, _vp]),
    "gcb_cot_receive_dev": (_int, [C.POINTER(Label), _vp, _vp, _vp, _u64, _vp, _u32, _vp]),
    "gcb_rot_send_dev": (_int, [C.POINTER(Label), C.POINTER(Label), _vp, _u64, _vp, _vp]),
    "gcb_rot_receive_dev": (_int, [C.POINTER(Label), _vp, _u64, _vp, _vp]),
    "gcb_circuit_parse": (_int, [_vp, _sz, _int, C.POINTER(_vp)]),
    "gcb_circuit_from_gates": (_int, [_vp, _u32, _u32, _vp, _u32, _vp, _u32, C.POINTER(_vp)]),
    "gcb_circuit_destroy": (None, [_vp]),
    "gcb_circuit_get_info": (_int, [_vp, _vp]),
    "gcb_circuit_get_gates": (_int, [_vp, _vp]),
    "gcb_circuit_get_io": (_int, [_vp, _vp, _vp]),
    "gcb_circuit_compute": (_int, [_vp, _u32, _vp, _vp]),
    "gcb_circuit_plan": (_int, [_vp, C.POINTER(_vp)]),
    "gcb_tables_wire_size": (_int, [_vp, C.POINTER(_sz)]),
    "gcb_tables_to_wire": (_int, [_vp, _u32, _vp, _vp, _sz]),
    "gcb_tables_from_wire": (_int, [_vp, _u32, _vp, _sz, _vp]),
    "gcb_tables_to_wire_dev": (_int, [_vp, _u32, _vp, _vp, _sz, _vp]),
    "gcb_tables_from_wire_dev": (_int, [_vp, _u32, _vp, _sz, _vp, _vp]),
    "gcb_iknp_check_sums": (_int, [C.POINTER(Label), _u64, _vp, _vp, _u64, _vp]),
    "gcb_iknp_check_sums_dev": (_int, [C.POINTER(Label), _u64, _vp, _vp, _u64, _vp, _vp]),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        so = _build.build()            # no-op when the in-tree .so is current
        L = C.CDLL(so)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)      # AttributeError = the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise GcbError(rc, lib().gcb_last_error().decode())


def ptr(a) -> int | None:
    """Host numpy array / torch tensor / int / None -> raw address."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous(), "tensor must be contiguous"
        return a.data_ptr()
    raise TypeError(type(a))


def u8(x) -> np.ndarray:
    return np.frombuffer(bytes(x), dtype=np.uint8).copy()
