"""CDO weights ingestion, host side only: dict / .npz / netCDF-3 classic."""
import os

import numpy as np
import pytest

from smmregrid_b200 import CdoWeights, check_mask, synth


def test_from_mapping_and_sizes():
    w = synth.config_weights("C1")
    assert not w.is3d and w.n_levels == 1
    assert w.sizes == {"src_grid_size": 144 * 73, "dst_grid_size": 180 * 90}
    assert w["src_address"].min() == 1 and w["src_address"].max() <= 144 * 73      # 1-based (CDO)
    assert w["remap_matrix"].ndim == 2
    w2 = CdoWeights.from_any(dict(w.vars))
    assert w2.sizes == w.sizes
    w3 = w.assign(dst_grid_imask=np.zeros(180 * 90, np.int32))
    assert check_mask(w3) is True and check_mask(w) is False
    assert w["dst_grid_imask"].all()                                                  # assign copies
    with pytest.raises(TypeError):
        CdoWeights.from_any(42)


def test_npz_round_trip(tmp_path):
    w = synth.ocean3d_weights(36, 18, 12, 6, n_levels=4)
    path = os.path.join(tmp_path, "w3d.npz")
    w.save_npz(path)
    r = CdoWeights.from_file(path, mask_dim="lev")
    assert r.is3d and r.n_levels == 4 and np.array_equal(r.levels, w.levels)
    for k in ("src_address", "dst_address", "remap_matrix", "link_length", "src_grid_imask", "dst_grid_frac"):
        assert np.array_equal(r[k], w[k]), k
    assert r.attrs["normalization"] == "fracarea"
    # padded 3-D layout of cdogenerate.py:310-343: zeros beyond link_length
    for l, nl in enumerate(w["link_length"]):
        assert (w["src_address"][l, nl:] == 0).all() and (w["remap_matrix"][l, nl:] == 0).all()
    assert np.array_equal(check_mask(r, "lev"), np.zeros(4, bool))


def test_netcdf3_scrip_file(tmp_path):
    """A SCRIP-style weights file as CDO writes it (netCDF classic), read without xarray."""
    from scipy.io import netcdf_file
    w = synth.conservative_latlon(24, 12, 8, 4, num_wgts=3)
    path = os.path.join(tmp_path, "weights.nc")
    with netcdf_file(path, "w") as nc:
        nl = w["src_address"].size
        for name, n in (("src_grid_size", 288), ("dst_grid_size", 32), ("num_links", nl), ("num_wgts", 3),
                        ("src_grid_rank", 2), ("dst_grid_rank", 2)):
            nc.createDimension(name, n)
        spec = {"src_address": ("i", ("num_links",)), "dst_address": ("i", ("num_links",)),
                "remap_matrix": ("d", ("num_links", "num_wgts")), "src_grid_imask": ("i", ("src_grid_size",)),
                "dst_grid_imask": ("i", ("dst_grid_size",)), "dst_grid_frac": ("d", ("dst_grid_size",)),
                "src_grid_dims": ("i", ("src_grid_rank",)), "dst_grid_dims": ("i", ("dst_grid_rank",)),
                "dst_grid_center_lat": ("d", ("dst_grid_size",)), "dst_grid_center_lon": ("d", ("dst_grid_size",))}
        for name, (t, dims) in spec.items():
            v = nc.createVariable(name, t, dims)
            v[:] = w[name]
        nc.source_grid = "r24x12"
        nc.dest_grid = "r8x4"
    r = CdoWeights.from_file(path)
    assert r.sizes == {"src_grid_size": 288, "dst_grid_size": 32}
    assert r.attrs["source_grid"] == "r24x12"
    for k in ("src_address", "dst_address", "remap_matrix", "dst_grid_frac"):
        assert np.array_equal(r[k], w[k]), k
    assert r["remap_matrix"].shape[1] == 3


def test_synthetic_weights_are_cdo_shaped():
    """Conservative rows sum to 1 (fracarea), links sorted by destination then source."""
    for w in (synth.config_weights("C2", 8), synth.config_weights("C4", 10)):
        n_dst = w.sizes["dst_grid_size"]
        d, s = w["dst_address"].astype(np.int64), w["src_address"].astype(np.int64)
        assert (np.diff(d * 10**9 + s) > 0).all()
        rs = np.bincount(d - 1, weights=w["remap_matrix"][:, 0], minlength=n_dst)
        assert np.allclose(rs, 1.0, rtol=0, atol=1e-13)
        assert np.array_equal(w["dst_grid_frac"], np.ones(n_dst))
    # masked sources: dropped links, renormalised rows, frac = valid covered fraction
    mask = np.ones((18, 36), np.int32)
    mask[:, :18] = 0
    w = synth.conservative_latlon(36, 18, 12, 6, src_mask=mask)
    assert (mask.ravel()[w["src_address"] - 1] == 1).all()
    rs = np.bincount(w["dst_address"] - 1, weights=w["remap_matrix"][:, 0], minlength=72)
    assert set(np.round(rs, 12)) <= {0.0, 1.0}
    assert w["dst_grid_frac"].min() == 0.0 and w["dst_grid_frac"].max() == 1.0
