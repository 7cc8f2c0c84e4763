"""3-D weights over the GPUs of one box, levels partitioned over the ranks:
run under torchrun (NCCL), checks the gathered result against the CPU oracle on rank 0 and
prints operator build time / device bytes per rank next to the whole-operator figures.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29531 scripts/level_shard_2gpu.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smmregrid_b200 import Regridder, synth  # noqa: E402
from smmregrid_b200.shard import HostGather, LevelShardedRegridder  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        dist.init_process_group("nccl")
    Regridder(weights=synth.conservative_latlon(36, 18, 12, 6)).regrid(np.zeros((1, 18, 36), np.float32))   # context, library
    w = synth.config_weights("C3")                      # 75 levels, 362x292 -> r360x180, per-level masks
    L, n_src, n_dst = w.n_levels, w.sizes["src_grid_size"], w.sizes["dst_grid_size"]
    T = 8
    x = synth.synthetic_field((T, L, n_src), np.float32, seed=11, nan_mode="none")
    x[:, w["src_grid_imask"] == 0] = np.nan             # missing where the level's mask says so
    t0 = time.perf_counter()
    sh = LevelShardedRegridder(w, remap_area_min=0.5)
    t_build = time.perf_counter() - t0
    xd = torch.from_numpy(x).cuda()
    y = sh.regrid(xd, dst=0)                            # NCCL gather
    hg = HostGather(L, (T, 180, 360), torch.float64, dst=0)
    y_shm = sh.regrid(xd, dst=0, via="shm", out=hg)     # per-rank D2H into shared memory
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    ev0.record()
    for _ in range(5):
        sh.regrid(xd, gather=False)
    ev1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1) / 5], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    info = (rank, sh.start, sh.stop, round(t_build, 2),
            sum(sh.local.weights_matrix.info(l)["device_bytes"] for l in range(len(sh.local.weights_matrix))) if sh.local else 0)
    infos = [None] * world
    if world > 1:
        dist.all_gather_object(infos, info)
    else:
        infos = [info]
    if rank == 0:
        from oracle import oracle
        t0 = time.perf_counter()
        whole = Regridder(weights=w, remap_area_min=0.5)
        t_whole = time.perf_counter() - t0
        y_whole = whole.regrid(xd).cpu().numpy()
        yn = y.numpy()
        assert yn.shape == (T, L, 180, 360), yn.shape
        assert np.array_equal(yn, y_whole, equal_nan=True), "level-partitioned result differs from the whole operator"
        assert np.array_equal(y_shm.numpy(), y_whole, equal_nan=True), "shm route differs"
        # oracle on a few levels (the whole C3 on the CPU takes a while)
        mats = oracle.compute_weights_matrix3d_np(w["src_address"], w["dst_address"], w["remap_matrix"], w["link_length"],
                                                  n_src, n_dst, builder=oracle.compute_weights_matrix_c)
        for l in (0, L // 2 - 1, L // 2, L - 1):
            im, _ = oracle.mask_tensordot_c(w["src_grid_imask"][l], mats[l])
            ref = oracle.apply_weights_c(x[:, l], mats[l], im, w["dst_grid_frac"][l], 0.5, bool((im == 0).any()))
            got = yn[:, l].reshape(T, n_dst)
            assert np.array_equal(np.isnan(got), np.isnan(ref)), l
            ok = ~np.isnan(ref)
            assert np.max(np.abs(got[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1e-300)) <= 1e-12, l
        print(f"world {world}: C3 75 levels, {T} steps: gathered result == whole operator (bitwise) == oracle (1e-12) on 4 levels")
        print(f"apply of the local levels, max over ranks: {ms.item():.3f} ms")
        whole_bytes = sum(whole.weights_matrix.info(l)["device_bytes"] for l in range(L))
        print(f"whole operator on one GPU: built in {t_whole:.2f} s, {whole_bytes / 1e6:.0f} MB on the device")
        for r, a, b, tb, nbytes in infos:
            print(f"  rank {r}: levels [{a}, {b}) built in {tb:.2f} s, {nbytes / 1e6:.0f} MB on the device")
    hg.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
