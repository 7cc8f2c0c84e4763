"""The CUDA path against the golden fixtures produced by the reference's own code
(tests/golden/make_golden.py), through the reference-shaped public API."""
import os

import numpy as np
import pytest

from helpers import assert_parity

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def weights_from(g):
    from smmregrid_b200 import CdoWeights
    n_dst = int(np.prod(g["in_dst_grid_dims"]))
    v = {k[3:]: g[k] for k in g if k.startswith("in_") and k != "in_x"}
    v["dst_grid_imask"] = np.ones_like(g["dst_grid_imask"])      # recomputed by mask_weights
    v["dst_grid_center_lat"] = np.zeros(n_dst)
    v["dst_grid_center_lon"] = np.zeros(n_dst)
    return CdoWeights(v, levels=g.get("levels"), mask_dim="lev" if "levels" in g else None)


@pytest.mark.parametrize("name", ["bil_f32", "con_masked_f32", "con_masked_f64", "unsorted_dups_f64"])
@pytest.mark.parametrize("kernel", [None, "gather"])
def test_golden_2d(smm_lib, cuda, name, kernel):
    from smmregrid_b200 import Regridder
    g = load(name)
    for i, am in enumerate(g["area_mins"]):
        rg = Regridder(weights=weights_from(g), remap_area_min=float(am), kernel=kernel)
        assert rg.masked == bool(g["masked"])
        assert np.array_equal(rg.weights["dst_grid_imask"].ravel(), g["dst_grid_imask"].ravel())
        y = rg.regrid(g["in_x"])
        ref = g[f"y_{i}"]
        assert y.shape == ref.shape and y.dtype == np.float64
        assert_parity(y, ref, 1e-12, f"{name} area_min={am}")
        # float32 output mode rounds the float64 result once
        rg32 = Regridder(weights=weights_from(g), remap_area_min=float(am), out_dtype=np.float32)
        y32 = rg32.regrid(g["in_x"])
        assert y32.dtype == np.float32
        with np.errstate(over="ignore"):
            ref32 = ref.astype(np.float32)
        fin = np.isfinite(ref32) | np.isnan(ref32)          # float64 leaks beyond float32 range -> inf
        assert_parity(np.where(fin, y32, 0), np.where(fin, ref32, 0), 1e-6, f"{name} f32")


def test_golden_3d(smm_lib, cuda):
    from smmregrid_b200 import Regridder
    g = load("ocean3d_f32")
    rg = Regridder(weights=weights_from(g), remap_area_min=float(g["area_min"]))
    assert np.array_equal(np.asarray(rg.masked), g["masked"])
    assert np.array_equal(rg.weights["dst_grid_imask"], g["dst_grid_imask"])
    y = rg.regrid(g["in_x"])
    assert y.shape == g["y"].shape
    assert_parity(y, g["y"], 1e-12, "ocean3d")
    # level subset by coordinate value (tests/levels_test.py of the reference)
    sel = [1, 2, 4]
    ys = rg.regrid(g["in_x"][:, sel], levels=g["levels"][sel])
    assert_parity(ys, g["y"][:, sel], 1e-12, "ocean3d subset")
