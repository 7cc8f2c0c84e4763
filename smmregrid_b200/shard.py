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


def gather_to_host(y_local, B: int, group=None, dst: Optional[int] = 0):
    """Assemble the full ``[B, ...]`` result from per-rank blocks (order = rank order) -- the one
    collective of the path, after the hot loop ("the final host gather").

    ``y_local``: this rank's ``[b_rank, ...]`` torch tensor (CUDA with NCCL, CPU with gloo).
    Returns a CPU tensor on rank ``dst`` (pinned when the blocks are CUDA tensors), else ``None``;
    ``dst=None`` returns it on every rank.  With a destination rank the blocks travel once
    (``gather``: the device of rank ``dst`` receives the other ranks' blocks over NVLink and
    every block is copied device->host straight into its rows of the result); ``dst=None`` uses
    ``all_gather_into_tensor``.  Blocks shorter than ``ceil(B / world)`` rows (the last ranks)
    are padded for the collective only.
    """
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        if not y_local.is_cuda:
            return y_local
        host = torch.empty(tuple(y_local.shape), dtype=y_local.dtype, pin_memory=True)
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
        out = torch.empty((world * per,) + tail, dtype=y_local.dtype, device=y_local.device)
        dist.all_gather_into_tensor(out, block, group=group)
        return out[:B].cpu()
    if rank != dst:
        dist.gather(block, gather_list=None, dst=dst, group=group)
        return None
    bufs = [block if r == rank else torch.empty_like(block) for r in range(world)]
    dist.gather(block, gather_list=bufs, dst=dst, group=group)
    host = torch.empty((B,) + tail, dtype=y_local.dtype, pin_memory=y_local.is_cuda)
    for r in range(world):
        a, b = batch_shard(B, world, r)
        if b > a:
            host[a:b].copy_(bufs[r][: b - a], non_blocking=True)
    if y_local.is_cuda:
        torch.cuda.current_stream(y_local.device).synchronize()
    return host


def regrid_sharded(regridder, x, group=None, gather: bool = True, dst: Optional[int] = 0):
    """Regrid this rank's block of the leading (batch) axis of ``x`` and optionally gather.

    ``x`` is the FULL array on every rank (or a lazily sliceable object); only
    ``x[start:stop]`` of this rank is touched.  2-D operators only (for 3-D weights shard the
    time axis yourself and call ``Regridder.regrid`` on the block).
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
    return gather_to_host(y_local, B, group=group, dst=dst)


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
