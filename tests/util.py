"""Shared helpers of the parity tests (oracle = checker, never the product)."""
import numpy as np

from mpc_b200.circuit_io import LABEL_DTYPE, WIRE_DTYPE
from mpc_b200.drbg import DRBG, garble_inputs


def rand_to_labels(rand: np.ndarray, ninputs: int):
    """uint8[batch, 16*(1+nin)] reader bytes -> (R LABEL[batch], L0 LABEL[batch, nin]) (Label.SetData, BE)."""
    batch = rand.shape[0]
    lab = np.ascontiguousarray(rand).view(">u8").astype("<u8").reshape(batch, 1 + ninputs, 2)
    out = np.zeros((batch, 1 + ninputs), dtype=LABEL_DTYPE)
    out["d0"], out["d1"] = lab[..., 0], lab[..., 1]
    return np.ascontiguousarray(out[:, 0]), np.ascontiguousarray(out[:, 1:])


def select(io_wires: np.ndarray, bits: np.ndarray) -> np.ndarray:
    """LabelForBit over [batch, n] wires and bits."""
    return np.where(bits.astype(bool), io_wires["l1"], io_wires["l0"]).astype(LABEL_DTYPE)


def decode(out_wires: np.ndarray, labels: np.ndarray) -> np.ndarray:
    """BitFromLabel: 0/1, or 2 for an unknown label."""
    is0 = labels == out_wires["l0"]
    is1 = labels == out_wires["l1"]
    return np.where(is0, 0, np.where(is1, 1, 2)).astype(np.uint8)


def eq(a: np.ndarray, b: np.ndarray) -> bool:
    return a.shape == b.shape and a.tobytes() == b.tobytes()


def drbg_labels(tag: str, n: int) -> np.ndarray:
    raw = DRBG(tag).array(16 * n)
    return raw.view("<u8").reshape(n, 2).copy().view(LABEL_DTYPE).reshape(n)


__all__ = ["rand_to_labels", "select", "decode", "eq", "drbg_labels", "garble_inputs", "DRBG", "LABEL_DTYPE", "WIRE_DTYPE"]
