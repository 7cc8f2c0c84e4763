"""Batch-axis sharding over the GPUs of one box.

Every batch row (time step x level x any other kept dim) is an independent product with the
same operator (the reference contracts only ``n_src``, ``regrid.py:550``; dask splits exactly
this axis into chunks), so ranks take contiguous blocks of the flattened batch axis, the
operator is replicated per GPU and the hot path needs no collective.  The only communication
is the optional final gather of the (small) destination slabs.
"""
from __future__ import annotations

from typing import Optional, Tuple


def batch_shard(B: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block ``[start, stop)`` of ``B`` batch rows owned by ``rank``:
    ``ceil(B / world)`` rows per rank, the last ranks possibly short or empty."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    per = -(-B // world)
    start = min(B, rank * per)
    return start, min(B, start + per)


def shard_sizes(B: int, world: int):
    return [batch_shard(B, world, r)[1] - batch_shard(B, world, r)[0] for r in range(world)]


class HostGather:
    """Reusable destination of :func:`gather_to_host` with ``via="shm"``: one ``[B, ...]`` array in
    POSIX shared memory that every rank of the box maps; each rank's slice is registered with the
    CUDA driver once, so its block goes device -> host by DMA straight into place, all ranks in
    parallel over their own PCIe links (no collective, no staging copy on a single rank).

    Create it collectively (all ranks, same arguments); rank ``dst`` reads ``.array`` (a CPU torch
    tensor view of the shared block) and must be done with it before the ranks gather into it
    again.  ``close()`` collectively when done.
    """

    def __init__(self, B: int, tail, dtype, group=None, dst: int = 0, register: bool = True):
        import numpy as np
        import torch
        import torch.distributed as dist
        from multiprocessing import shared_memory

        self.B, self.tail, self.dtype, self.group, self.dst = int(B), tuple(tail), dtype, group, dst
        dist_on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if dist_on else 1
        self.rank = dist.get_rank(group) if dist_on else 0
        itemsize = torch.empty((), dtype=dtype).element_size()
        row = int(np.prod(self.tail, dtype=np.int64)) * itemsize
        nbytes = max(1, self.B * row)
        name = [None]
        if self.rank == dst:
            self._shm = shared_memory.SharedMemory(create=True, size=nbytes)
            name[0] = self._shm.name
        if dist_on:
            dist.broadcast_object_list(name, src=dst, group=group)
        if self.rank != dst:
            self._shm = shared_memory.SharedMemory(name=name[0])
            try:        # the creating rank owns the segment: keep this process's resource tracker out of it
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self._shm._name, "shared_memory")
            except Exception:
                pass
        flat = np.ndarray((nbytes,), dtype=np.uint8, buffer=self._shm.buf)
        self.array = torch.from_numpy(flat[: self.B * row]).view(dtype).reshape((self.B,) + self.tail)
        self.start, self.stop = batch_shard(self.B, self.world, self.rank)
        self.mine = self.array[self.start:self.stop]
        self._registered = None
        if register and torch.cuda.is_available() and self.mine.numel():
            rt = torch.cuda.cudart()
            ptr, size = self.mine.data_ptr(), self.mine.numel() * itemsize
            if int(rt.cudaHostRegister(ptr, size, 0)) == 0:      # else: pageable copy, still correct
                self._registered = ptr
        if dist_on:
            dist.barrier(group=group)

    def close(self):
        import torch
        import torch.distributed as dist
        if self._registered is not None:
            torch.cuda.cudart().cudaHostUnregister(self._registered)
            self._registered = None
        self.array = self.mine = None
        if dist.is_available() and dist.is_initialized():
            dist.barrier(group=self.group)
        try:
            self._shm.close()
            if self.rank == self.dst:
                self._shm.unlink()
        except Exception:
            pass


def gather_to_host(y_local, B: int, group=None, dst: Optional[int] = 0, via: str = "nccl", out=None):
    """Assemble the full ``[B, ...]`` result from per-rank blocks (order = rank order) -- the one
    exchange of the path, after the hot loop ("the final host gather").

    ``y_local``: this rank's ``[b_rank, ...]`` torch tensor (CUDA with NCCL, CPU with gloo).
    Returns a CPU tensor on rank ``dst``, else ``None``; ``dst=None`` returns it on every rank.

    * ``via="nccl"`` (any backend): with a destination rank the blocks travel once (``gather``: the
      device of rank ``dst`` receives the other ranks' blocks over NVLink and every block is copied
      device->host straight into its rows of a pinned result -- ``out`` if given); ``dst=None`` uses
      ``all_gather_into_tensor``.  The whole result crosses ONE PCIe link.
    * ``via="shm"`` (ranks on one box, ``out`` = a :class:`HostGather`): every rank copies its block
      device->host into its rows of the shared array over its OWN PCIe link, then a barrier; no
      collective moves data.
    Blocks shorter than ``ceil(B / world)`` rows (the last ranks) are padded for the collective only.
    """
    import torch
    import torch.distributed as dist

    dist_on = dist.is_available() and dist.is_initialized()
    if via == "shm":
        if not isinstance(out, HostGather):
            raise ValueError('via="shm" needs out=HostGather(...) created collectively beforehand')
        if out.mine.shape[0]:
            out.mine.copy_(y_local[: out.mine.shape[0]], non_blocking=True)
        if y_local.is_cuda:
            torch.cuda.current_stream(y_local.device).synchronize()
        if dist_on:
            dist.barrier(group=group)
        return out.array if (dst is None or out.rank == dst) else None
    if via != "nccl":
        raise ValueError('via must be "nccl" or "shm"')
    if not dist_on:
        if not y_local.is_cuda:
            return y_local
        host = out if out is not None else torch.empty(tuple(y_local.shape), dtype=y_local.dtype, pin_memory=True)
        host.copy_(y_local, non_blocking=True)
        torch.cuda.current_stream(y_local.device).synchronize()
        return host
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = -(-B // world)
    tail = tuple(y_local.shape[1:])
    block = y_local.contiguous()
    if block.shape[0] != per:
        block = torch.zeros((per,) + tail, dtype=y_local.dtype, device=y_local.device)
        block[: y_local.shape[0]] = y_local
    if dst is None:
        full = torch.empty((world * per,) + tail, dtype=y_local.dtype, device=y_local.device)
        dist.all_gather_into_tensor(full, block, group=group)
        return full[:B].cpu()
    if rank != dst:
        dist.gather(block, gather_list=None, dst=dst, group=group)
        return None
    bufs = [block if r == rank else torch.empty_like(block) for r in range(world)]
    dist.gather(block, gather_list=bufs, dst=dst, group=group)
    host = out if out is not None else torch.empty((B,) + tail, dtype=y_local.dtype, pin_memory=y_local.is_cuda)
    for r in range(world):
        a, b = batch_shard(B, world, r)
        if b > a:
            host[a:b].copy_(bufs[r][: b - a], non_blocking=True)
    if y_local.is_cuda:
        torch.cuda.current_stream(y_local.device).synchronize()
    return host


def regrid_sharded(regridder, x, group=None, gather: bool = True, dst: Optional[int] = 0, via: str = "nccl", out=None):
    """Regrid this rank's block of the leading (batch) axis of ``x`` and optionally gather.

    ``x`` is the FULL array on every rank (or a lazily sliceable object); only
    ``x[start:stop]`` of this rank is touched.  For 3-D weights this splits the leading (time)
    axis with the whole operator on every rank; :class:`LevelShardedRegridder` gives each rank
    only the operators of its own levels instead.
    """
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    B = x.shape[0]
    start, stop = batch_shard(B, world, rank)
    y_local = regridder.regrid(x[start:stop])
    if not gather:
        return y_local
    import torch
    if not isinstance(y_local, torch.Tensor):
        y_local = torch.from_numpy(y_local)
    return gather_to_host(y_local, B, group=group, dst=dst, via=via, out=out)


class LevelShardedRegridder:
    """3-D weights over the GPUs of one box ("partition (time x level) pairs so
    each GPU holds only the matrices of its levels when L >= G, else split time within level").

    * ``n_levels >= world``: rank r builds the operator of ITS contiguous block of levels only
      (``CdoWeights.isel_levels``: operator construction and device memory shrink by the number
      of ranks) and regrids those levels of every time step;
    * fewer levels than ranks: every rank holds the whole operator and takes a block of the
      leading (time) axis.

    Either way there is no data-path collective; ``regrid`` optionally gathers the result on the
    host (``gather_to_host``).  Construct it on every rank of the process group, after
    ``torch.cuda.set_device``; keyword arguments go to :class:`Regridder`.
    """

    def __init__(self, weights, group=None, **regridder_kw):
        import numpy as np
        import torch.distributed as dist
        from .regrid import Regridder
        from .weights import CdoWeights
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        else:
            self.world, self.rank = 1, 0
        w = CdoWeights.from_any(weights, mask_dim=regridder_kw.get("mask_dim"))
        if not w.is3d:
            raise ValueError("LevelShardedRegridder needs 3-D weights; use regrid_sharded for 2-D ones")
        self.levels = w.levels
        self.n_levels = w.n_levels
        self.n_src = w.sizes["src_grid_size"]
        self.src_grid_shape = tuple(int(n) for n in np.atleast_1d(w["src_grid_dims"])[::-1])
        self.by_level = self.n_levels >= self.world
        if self.by_level:
            self.start, self.stop = batch_shard(self.n_levels, self.world, self.rank)
            self.local = Regridder(weights=w.isel_levels(slice(self.start, self.stop)), **regridder_kw) \
                if self.stop > self.start else None       # (ceil partition: the last ranks may have no level)
        else:
            self.start, self.stop = 0, self.n_levels
            self.local = Regridder(weights=w, **regridder_kw)

    def _n_kept(self, shape) -> int:
        """Number of leading (non-horizontal) axes of the data (``Regridder._n_horizontal_axes``)."""
        shape = tuple(int(n) for n in shape)
        gs = self.src_grid_shape
        if len(shape) > len(gs) and shape[-len(gs):] == gs:
            return len(shape) - len(gs)
        if len(shape) > 1 and shape[-1] == self.n_src:
            return len(shape) - 1
        raise KeyError('Dimensions mismatch')

    def regrid(self, x, level_axis=None, gather: bool = True, dst: Optional[int] = 0, via: str = "nccl", out=None):
        """``x``: the FULL ``[time..., level, (lat, lon | cell)]`` array on every rank (numpy / torch /
        anything sliceable with basic indexing); only this rank's levels -- or time steps -- are
        touched.  ``gather=False``: this rank's block (``[..., L_local, tgt...]``, ``None`` on a rank
        without levels; or its time steps).  Otherwise the whole result as a host tensor on rank
        ``dst`` (``None`` elsewhere; ``dst=None``: on every rank)."""
        import numpy as np
        import torch
        if not self.by_level:
            start, stop = batch_shard(x.shape[0], self.world, self.rank)
            y_local = self.local.regrid3d(x[start:stop], level_axis=level_axis)
            if not isinstance(y_local, torch.Tensor):
                y_local = torch.from_numpy(np.asarray(y_local))
            return y_local if not gather else \
                gather_to_host(y_local, x.shape[0], group=self.group, dst=dst, via=via, out=out)
        ndim = len(x.shape)
        nkept = self._n_kept(x.shape)
        la = nkept - 1 if level_axis is None else level_axis % ndim
        if la >= nkept:
            raise ValueError("level_axis must be one of the non-horizontal axes")
        if x.shape[la] != self.n_levels:
            raise ValueError(f"data has {x.shape[la]} levels, weights {self.n_levels}")
        y_local = None
        if self.local is not None:
            sl = [slice(None)] * ndim
            sl[la] = slice(self.start, self.stop)
            y_local = self.local.regrid3d(x[tuple(sl)], level_axis=la, levels=self.levels[self.start:self.stop],
                                          transpose=True)      # [kept without level..., L_local, tgt...]
            if not isinstance(y_local, torch.Tensor):
                y_local = torch.from_numpy(np.asarray(y_local))
        if not gather:
            return y_local
        # the gather works on the leading axis: levels to the front, and back afterwards (a view)
        y_local = self._level_block(y_local, nkept)
        full = gather_to_host(y_local, self.n_levels, group=self.group, dst=dst, via=via, out=out)
        return None if full is None else torch.movedim(full, 0, nkept - 1)

    def _level_block(self, y_local, nkept):
        """``[L_local, ...]`` contiguous block for the gather.  A rank without levels still takes
        part: it learns the shape and dtype of one level's block from rank 0 (host-side metadata)."""
        import torch
        import torch.distributed as dist
        if y_local is not None:
            y_local = torch.movedim(y_local, nkept - 1, 0).contiguous()
        if batch_shard(self.n_levels, self.world, self.world - 1)[1] > batch_shard(self.n_levels, self.world, self.world - 1)[0]:
            return y_local                                 # every rank has levels: nothing to agree on
        meta = [(tuple(y_local.shape[1:]), y_local.dtype, y_local.device.type) if self.rank == 0 else None]
        dist.broadcast_object_list(meta, src=0, group=self.group)
        if y_local is None:
            shape, dtype, dev = meta[0]
            device = torch.device("cuda", torch.cuda.current_device()) if dev == "cuda" else torch.device("cpu")
            y_local = torch.empty((0,) + tuple(shape), dtype=dtype, device=device)
        return y_local


def bind_to_gpu_numa(device_index: int) -> Optional[int]:
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off, so that pinned
    host staging buffers (first-touch) and the H2D/D2H copies stay on the local socket.  On an
    8-GPU box all ranks otherwise tend to allocate on one node and share its memory controllers
    and inter-socket links.  Returns the node, or None when the topology cannot be read (then
    nothing is changed).  Call before allocating pinned memory."""
    import os
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None
