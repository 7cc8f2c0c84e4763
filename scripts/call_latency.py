"""Per-call latency of small regrids on one GPU (host numpy in, host numpy out, and device in/out):
what an interactive user sees -- bandwidth does not matter at these sizes, fixed costs do.
Usage: python scripts/call_latency.py > gpurun_out/<tag>_call_latency.txt"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smmregrid_b200 import Regridder, synth  # noqa: E402


def med(f, n=60):
    for _ in range(5):
        f()
    t = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f()
        torch.cuda.synchronize()
        t.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(t)), 1e3 * float(np.percentile(t, 90))


cases = [("HEALPix nside 32 -> r180x90 nn, 12 steps", synth.healpix_weights(32, 180, 90, 1), (12, 12288)),
         ("r360x180 -> r180x90 con, 1 step", synth.conservative_latlon(360, 180, 180, 90), (1, 180, 360)),
         ("r360x180 -> r180x90 con, 12 steps", synth.conservative_latlon(360, 180, 180, 90), (12, 180, 360)),
         ("1440x721 -> r360x180 con (C2), 1 step", synth.config_weights("C2"), (1, 721, 1440)),
         ("1440x721 -> r360x180 con (C2), 24 steps", synth.config_weights("C2"), (24, 721, 1440)),
         ("3600x1800 -> r360x180 con (C4), 1 step", synth.config_weights("C4"), (1, 1800, 3600))]
print(f"{'case':48s} {'host numpy in/out ms (median, p90)':>36s} {'device in/out ms':>20s} {'input MB':>9s}")
for name, w, shape in cases:
    t0 = time.perf_counter()
    rg = Regridder(weights=w, remap_area_min=0.5)
    build = time.perf_counter() - t0
    x = synth.synthetic_field(shape, np.float32, seed=1)
    xd = torch.from_numpy(x).cuda()
    h = med(lambda: rg.regrid(x))
    d = med(lambda: rg.regrid(xd))
    print(f"{name:48s} {h[0]:18.3f} {h[1]:17.3f} {d[0]:10.3f} {d[1]:9.3f} {x.nbytes / 1e6:9.1f}   (operator built in {build:.2f} s)")
