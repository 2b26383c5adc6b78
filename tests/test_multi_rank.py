"""The N>1 path on CPU: two gloo ranks shard a batch the way bench.py does, run their
shards through the oracle, and the gathered result must equal the unsharded run; the
max-over-ranks reduction used for timing works over the process group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_circuit
from mpc_b200.shard import iknp_row_range, instance_range
from oracle import pyoracle as O
from util import drbg_labels, garble_inputs


def test_instance_ranges_partition_the_batch():
    for batch in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 3, 8):
            r = [instance_range(batch, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_iknp_ranges_are_chunk_aligned():
    for n in (1, 511, 512, 513, 5000, 1 << 24):
        for world in (1, 2, 8):
            r = [iknp_row_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            for (lo, hi, off), nxt in zip(r, r[1:] + [(n, n, 0)]):
                assert hi == nxt[0] and (lo % 512 == 0 or lo == n) and (off == lo // 8 or lo == n)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    circ = load_circuit("add64")
    batch = 9
    keys, rand = garble_inputs("mr", batch, circ.num_inputs, 16)
    lo, hi = instance_range(batch, rank, world)
    _, tables, io = O.garble_batch(circ, keys[lo:hi], rand[lo:hi])
    # gather shard table digests on rank 0 (no data-path collective is needed for the result itself)
    t = torch.from_numpy(tables.view(np.uint8).reshape(hi - lo, -1).astype(np.int64).sum(axis=1))
    pad = torch.zeros(batch, dtype=torch.int64)
    pad[lo:hi] = t
    dist.all_reduce(pad, op=dist.ReduceOp.SUM)
    # IKNP: each rank expands its own row range from the same keys
    k0, k1 = drbg_labels("mr/k0", 128), drbg_labels("mr/k1", 128)
    n = 3000
    b = (np.arange(n) % 3 == 0).astype(np.uint8)
    rlo, rhi, off = iknp_row_range(n, rank, world)
    _, lab, _ = O.iknp_receive(k0, k1, 17 + off, b[rlo:rhi])
    ms = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    q.put((rank, pad.numpy(), lab.tobytes(), float(ms)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_gloo_ranks_reproduce_the_unsharded_run():
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    circ = load_circuit("add64")
    keys, rand = garble_inputs("mr", 9, circ.num_inputs, 16)
    _, tables, _ = O.garble_batch(circ, keys, rand)
    want = tables.view(np.uint8).reshape(9, -1).astype(np.int64).sum(axis=1)
    assert np.array_equal(res[0][1], want) and np.array_equal(res[1][1], want)
    k0, k1 = drbg_labels("mr/k0", 128), drbg_labels("mr/k1", 128)
    b = (np.arange(3000) % 3 == 0).astype(np.uint8)
    _, lab, _ = O.iknp_receive(k0, k1, 17, b)
    assert res[0][2] + res[1][2] == lab.tobytes()
    assert res[0][3] == res[1][3] == 11.0
