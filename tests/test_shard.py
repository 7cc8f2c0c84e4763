"""Multi-rank host logic on CPU: world_size-2 gloo process group (no GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from smmregrid_b200.shard import HostGather, batch_shard, gather_to_host, shard_sizes


def test_batch_shard_partitions():
    for B in (0, 1, 7, 8, 1095, 8760):
        for world in (1, 2, 3, 4, 8):
            blocks = [batch_shard(B, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            assert sum(shard_sizes(B, world)) == B
            assert max(shard_sizes(B, world)) == -(-B // world)
    assert shard_sizes(8760, 8) == [1095] * 8          # BASELINE config C4 over one 8-GPU box
    with pytest.raises(ValueError):
        batch_shard(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, n_dst, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, stop = batch_shard(B, world, rank)
        # each rank "regrids" its block: rows carry their global batch index
        y_local = (torch.arange(start, stop, dtype=torch.float64)[:, None] * 1000
                   + torch.arange(n_dst, dtype=torch.float64)[None, :])
        full = gather_to_host(y_local, B, dst=0)
        everywhere = gather_to_host(y_local, B, dst=None)
        # shared-memory gather: every rank writes its own rows, no collective moves data
        hg = HostGather(B, (n_dst,), torch.float64, dst=0)
        shm = gather_to_host(y_local, B, dst=0, via="shm", out=hg)
        shm = None if shm is None else shm.numpy().copy()
        dist.barrier()                               # the result is consumed before the array is written again
        again = gather_to_host(y_local + 1.0, B, dst=0, via="shm", out=hg)       # the destination is reusable
        again = None if again is None else again.numpy().copy()
        hg.close()
        q.put((rank, None if full is None else full.numpy(), everywhere.numpy(), shm, again))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [5, 8])
def test_gather_world2_gloo(B):
    world, n_dst = 2, 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, n_dst, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        rank, full, everywhere, shm, again = q.get(timeout=120)
        res[rank] = (full, everywhere, shm, again)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.arange(B)[:, None] * 1000.0 + np.arange(n_dst)[None, :]
    assert res[1][0] is None
    assert np.array_equal(res[0][0], expect)
    assert np.array_equal(res[0][1], expect) and np.array_equal(res[1][1], expect)
    assert res[1][2] is None and np.array_equal(res[0][2], expect) and np.array_equal(res[0][3], expect + 1.0)
