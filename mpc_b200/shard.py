"""How the hot path shards across the GPUs of one box.

Instances of a garble / eval batch are independent and IKNP rows are
independent per 512-row chunk (the CTR keystream is random access), so each
rank takes one contiguous block and there is no data-path collective.
"""
from __future__ import annotations

from typing import Tuple

IKNP_CHUNK_ROWS = 512          # ot/iknp.go:64-77


def instance_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of instances [lo, hi) of rank `rank`: sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def iknp_row_range(n: int, rank: int, world: int) -> Tuple[int, int, int]:
    """Rows [lo, hi) of rank `rank`, cut at chunk boundaries, and the keystream byte offset
    (relative to the call's stream position) at which the rank's first chunk starts."""
    chunks = (n + IKNP_CHUNK_ROWS - 1) // IKNP_CHUNK_ROWS
    clo, chi = instance_range(chunks, rank, world)
    lo, hi = min(n, clo * IKNP_CHUNK_ROWS), min(n, chi * IKNP_CHUNK_ROWS)
    return lo, hi, clo * (IKNP_CHUNK_ROWS // 8)
