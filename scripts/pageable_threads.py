"""Pageable (plain numpy) host input through Regridder.regrid on one GPU: throughput against the
number of host threads that fill the pinned bounce buffers (SMM_HOST_COPY_THREADS, default 12)
and the staging chunk size.  Usage: python scripts/pageable_threads.py > gpurun_out/<tag>_pageable_threads.txt"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smmregrid_b200 import Regridder, synth  # noqa: E402

w = synth.config_weights("C4")
rg = Regridder(weights=w, remap_area_min=0.5)
x = synth.synthetic_field((128, 1800, 3600), np.float32, seed=1)          # 3.3 GB, pageable
print(f"cpus available: {len(os.sched_getaffinity(0))}; input {x.nbytes / 1e9:.2f} GB pageable numpy, output pageable numpy")
for chunk in ("", "32", "128"):
    for nt in (4, 8, 12, 16, 20, 24, 32):
        os.environ["SMM_HOST_COPY_THREADS"] = str(nt)
        if chunk:
            os.environ["SMM_HOST_CHUNK_MB"] = chunk
        else:
            os.environ.pop("SMM_HOST_CHUNK_MB", None)
        rg.regrid(x)
        t = []
        for _ in range(4):
            t0 = time.perf_counter()
            rg.regrid(x)
            t.append(time.perf_counter() - t0)
        print(f"chunk {chunk or 'default (64)':>13s} MB  threads {nt:2d}: {x.nbytes / min(t) / 1e9:6.2f} GB/s best, {x.nbytes / np.median(t) / 1e9:6.2f} GB/s median")
