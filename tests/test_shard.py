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


# ------------------------------------------------------------------ 3-D weights: levels over ranks

def test_isel_levels_matches_level_slices_of_the_whole():
    from oracle import oracle
    from smmregrid_b200 import synth
    lev = np.array([0.5, 10.0, 100.0, 1000.0, 5000.0])
    w = synth.ocean3d_weights(36, 18, 12, 6, n_levels=5, seed=3, level_values=lev)
    n_src, n_dst = 36 * 18, 12 * 6
    full = oracle.compute_weights_matrix3d_np(w["src_address"], w["dst_address"], w["remap_matrix"], w["link_length"],
                                              n_src, n_dst, builder=oracle.compute_weights_matrix_c)
    for sel in (slice(0, 2), slice(2, 5), slice(4, 5), [3, 1]):
        sub = w.isel_levels(sel)
        idx = np.arange(5)[sel]
        assert sub.n_levels == len(idx) and np.array_equal(sub.levels, lev[idx]) and sub.mask_dim == w.mask_dim
        assert sub["src_address"].shape[1] == int(w["link_length"][idx].max())        # padding trimmed
        assert np.array_equal(sub["src_grid_dims"], w["src_grid_dims"])              # shared variables untouched
        mats = oracle.compute_weights_matrix3d_np(sub["src_address"], sub["dst_address"], sub["remap_matrix"],
                                                  sub["link_length"], n_src, n_dst, builder=oracle.compute_weights_matrix_c)
        for k, l in enumerate(idx):
            for name in ("src", "dst", "w"):
                assert np.array_equal(getattr(mats[k], name), getattr(full[l], name)), (sel, l, name)
        assert np.array_equal(sub["dst_grid_frac"], w["dst_grid_frac"][idx])
        assert np.array_equal(sub["src_grid_imask"], w["src_grid_imask"][idx])
    with pytest.raises(ValueError):
        synth.conservative_latlon(36, 18, 12, 6).isel_levels(slice(0, 1))


class _OracleRegridder:
    """CPU stand-in for Regridder (the real one needs a GPU): same constructor keywords and
    regrid3d signature, computed by the oracle -- lets the gloo ranks exercise the level
    partition, the per-rank weights and the gather."""

    def __init__(self, weights=None, remap_area_min=0.5, **kw):
        from oracle import oracle
        w = weights
        self.w, self.amin, self.o = w, remap_area_min, oracle
        self.n_src, self.n_dst = w.sizes["src_grid_size"], w.sizes["dst_grid_size"]
        self.mats = oracle.compute_weights_matrix3d_np(w["src_address"], w["dst_address"], w["remap_matrix"],
                                                       w["link_length"], self.n_src, self.n_dst,
                                                       builder=oracle.compute_weights_matrix_c)
        self.im = np.stack([oracle.mask_tensordot_c(w["src_grid_imask"][l], self.mats[l])[0] for l in range(w.n_levels)])

    def regrid3d(self, x, level_axis=None, levels=None, transpose=None):
        x = np.asarray(x)
        la = x.ndim - 2 if level_axis is None else level_axis
        levels = self.w.levels if levels is None else levels
        return self.o.regrid3d_np(x, la, levels, self.w.levels, self.mats, self.im, self.w["dst_grid_frac"],
                                  self.o.check_mask_np(self.im), self.amin)


def _level_worker(rank, world, port, n_levels, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        from smmregrid_b200 import synth
        from smmregrid_b200.shard import LevelShardedRegridder
        # (the package attribute `regrid` is the function of that name, as in the reference; the module is in sys.modules)
        sys.modules["smmregrid_b200.regrid"].Regridder = _OracleRegridder
        w = synth.ocean3d_weights(36, 18, 12, 6, n_levels=n_levels, seed=3)
        x = synth.synthetic_field((5, n_levels, 36 * 18), np.float32, seed=2, nan_mode="random")
        sh = LevelShardedRegridder(w, remap_area_min=0.5)
        local = sh.regrid(x, gather=False)
        full = sh.regrid(x, dst=0)
        everywhere = sh.regrid(x, dst=None, via="nccl")
        q.put((rank, sh.by_level, (sh.start, sh.stop), None if sh.local is None else sh.local.w.n_levels,
               None if local is None else tuple(local.shape), None if full is None else full.numpy().copy(),
               everywhere.numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_levels", [(2, 5), (3, 4), (2, 1)])
def test_levels_over_ranks_gloo(world, n_levels):
    """Levels >= ranks: each rank holds the operators of its own levels only (rank 2 of 3 gets
    none of 4 levels under the ceil partition and still takes part in the gather); one level on
    two ranks: the time axis is split instead.  The gathered result equals the whole-operator
    oracle result."""
    from oracle import oracle  # noqa: F401  (the workers use it)
    from smmregrid_b200 import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_level_worker, args=(r, world, port, n_levels, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r = q.get(timeout=180)
        res[r[0]] = r[1:]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w = synth.ocean3d_weights(36, 18, 12, 6, n_levels=n_levels, seed=3)
    x = synth.synthetic_field((5, n_levels, 36 * 18), np.float32, seed=2, nan_mode="random")
    expect = _OracleRegridder(weights=w, remap_area_min=0.5).regrid3d(x)
    assert expect.shape == (5, n_levels, 72)
    by_level = n_levels >= world
    for r in range(world):
        assert res[r][0] == by_level
        assert np.array_equal(res[r][5], expect, equal_nan=True)                      # dst=None: everywhere
        assert (res[r][4] is not None) == (r == 0)
    assert np.array_equal(res[0][4], expect, equal_nan=True)
    if by_level:
        blocks = [res[r][1] for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n_levels
        for r in range(world):
            nl = blocks[r][1] - blocks[r][0]
            assert res[r][2] == (nl if nl else None)                                  # operators of its own levels only
            assert res[r][3] == ((5, nl, 72) if nl else None)
        if (world, n_levels) == (3, 4):
            assert blocks[2] == (4, 4)
    else:
        assert all(res[r][2] == n_levels for r in range(world))
