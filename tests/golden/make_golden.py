#!/usr/bin/env python
"""Generate the golden fixtures in this directory by EXECUTING THE REFERENCE'S OWN SOURCE
(/root/reference/smmregrid/weights.py and Regridder.apply_weights of regrid.py) under the
numpy-backed stand-ins of ``refshim.py`` (xarray/dask/sparse are not installable here).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The fixtures (*.npz) are committed; tests/test_oracle.py checks the oracle against them and
tests/test_gpu_parity.py checks the CUDA path against them.  See refshim.py for what is the
reference's real code and what is restated (the sparse.COO container and its matmul loop).
"""
import logging
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import refshim  # noqa: E402

xr = refshim.install()
sys.path.insert(0, "/root/reference")
from smmregrid import weights as ref_weights  # noqa: E402  (the reference's own module)
from smmregrid.regrid import Regridder as RefRegridder  # noqa: E402

from smmregrid_b200 import synth  # noqa: E402  (only to produce CDO-shaped INPUTS)


class _Self:
    """The two attributes Regridder.apply_weights reads from self."""
    def __init__(self, remap_area_min):
        self.loggy = logging.getLogger("golden")
        self.remap_area_min = remap_area_min


def to_dataset(w):
    """CdoWeights -> stand-in xarray.Dataset with CDO's dimension names."""
    three_d = "link_length" in w.vars
    lev = ("lev",) if three_d else ()
    rank_d = len(np.atleast_1d(w["dst_grid_dims"]))
    rank_s = len(np.atleast_1d(w["src_grid_dims"]))
    ds = xr.Dataset({
        "src_address": (lev + ("num_links",), w["src_address"]),
        "dst_address": (lev + ("num_links",), w["dst_address"]),
        "remap_matrix": (lev + ("num_links", "num_wgts"), w["remap_matrix"]),
        "src_grid_imask": (lev + ("src_grid_size",), w["src_grid_imask"]),
        "dst_grid_imask": (lev + ("dst_grid_size",), w["dst_grid_imask"]),
        "dst_grid_frac": (lev + ("dst_grid_size",), w["dst_grid_frac"]),
        "src_grid_dims": (("src_grid_rank",), np.atleast_1d(w["src_grid_dims"])),
        "dst_grid_dims": (("dst_grid_rank",), np.atleast_1d(w["dst_grid_dims"])),
        "src_grid_rank": (("src_grid_rank",), np.arange(rank_s)),      # xarray's virtual dim coordinate
        "dst_grid_rank": (("dst_grid_rank",), np.arange(rank_d)),
        "dst_grid_center_lat": (("dst_grid_size",), w["dst_grid_center_lat"]),
        "dst_grid_center_lon": (("dst_grid_size",), w["dst_grid_center_lon"]),
    }, attrs={"source_grid": w.attrs.get("source_grid", "src"), "dest_grid": w.attrs.get("dest_grid", "dst")})
    if three_d:
        ds["link_length"] = (("lev",), w["link_length"])
    return ds


def run_apply(ds, matrix, masked, x, dims, area_min):
    """Regridder.apply_weights (regrid.py:458-628), the reference's code, on numpy data."""
    da = xr.DataArray(x, dims=dims, coords={}, name="var", attrs={"units": "K"})
    out = RefRegridder.apply_weights(_Self(area_min), da, ds, weights_matrix=matrix, masked=masked,
                                     horizontal_dims=["lon", "lat", "cell"])
    return np.asarray(out.data)


def inputs_of(w):
    keys = ["src_address", "dst_address", "remap_matrix", "src_grid_imask", "dst_grid_frac",
            "src_grid_dims", "dst_grid_dims"]
    if "link_length" in w.vars:
        keys.append("link_length")
    return {"in_" + k: w[k] for k in keys}


def case_2d(name, w, x, dims, area_mins=(0.0, 0.5, 0.9)):
    ds = to_dataset(w)
    matrix = ref_weights.compute_weights_matrix(ds)                  # weights.py:25-44
    ds2 = ref_weights.mask_weights(ds, matrix)                        # weights.py:55-84
    masked = bool(ref_weights.check_mask(ds2))                        # weights.py:103-120
    out = inputs_of(w)
    out.update(in_x=x, coo_src=matrix.coords[0].astype(np.int32), coo_dst=matrix.coords[1].astype(np.int32),
               coo_w=matrix.data, dst_grid_imask=np.asarray(ds2["dst_grid_imask"].data).astype(np.int32),
               masked=np.asarray(masked), area_mins=np.asarray(area_mins))
    for i, am in enumerate(area_mins):
        y = run_apply(ds2, matrix, masked, x, dims, am)
        out[f"y_{i}"] = y
        # the flag the reference would pass when the mask happens to be full
        out[f"y_unmasked_{i}"] = run_apply(ds2, matrix, False, x, dims, am)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: nnz={matrix.data.size} masked={masked} y={out['y_0'].shape} "
          f"nan={[int(np.isnan(out[f'y_{i}']).sum()) for i in range(len(area_mins))]}")


def case_3d(name, w, x, area_min=0.5):
    """Level-varying mask: compute_weights_matrix3d + mask_weights/check_mask with mask_dim, then the
    level loop of regrid3d (regrid.py:387-410) calling the reference's apply_weights per level."""
    ds = to_dataset(w)
    mats = ref_weights.compute_weights_matrix3d(ds, mask_dim="lev")   # weights.py:7-23
    ds2 = ref_weights.mask_weights(ds, mats, mask_dim="lev")
    masked = np.asarray(ref_weights.check_mask(ds2, mask_dim="lev"))
    L = len(mats)
    ys = []
    for l in range(L):                                                # regrid.py:387-409
        wa = ds2.isel(lev=l)
        wa = wa.isel(num_links=slice(0, int(w["link_length"][l])))
        ys.append(run_apply(wa, mats[l], bool(masked[l]), x[:, l], ("time", "lat", "lon"), area_min))
    y = np.moveaxis(np.stack(ys, axis=0), 0, 1)                       # concat + transpose (:410-427)
    out = inputs_of(w)
    out.update(in_x=x, dst_grid_imask=np.asarray(ds2["dst_grid_imask"].data).astype(np.int32),
               masked=masked, y=y, area_min=np.asarray(area_min), levels=w.levels,
               coo_nnz=np.array([m.data.size for m in mats]))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: L={L} masked={masked.astype(int)} y={y.shape} nan={int(np.isnan(y).sum())}")


def main():
    rng = np.random.default_rng(2024)

    # bilinear, regular lon-lat, float32 with NaN blobs, an inf and an all-NaN step (basic_test.py:31-39)
    w = synth.bilinear_latlon(36, 19, 45, 22, src_poles=True)
    x = (280 + 20 * rng.standard_normal((4, 19, 36))).astype(np.float32)
    x[0, 3:6, 10:14] = np.nan
    x[1, 8, 8] = np.inf
    x[1, 9, 9] = -np.inf
    x[3] = np.nan
    case_2d("bil_f32", w, x, ("time", "lat", "lon"))

    # conservative with a source land mask: dst_grid_frac < 1 and a masked destination
    mask = (rng.random((18, 36)) > 0.25).astype(np.int32)
    mask[5:9, 6:13] = 0
    w = synth.conservative_latlon(36, 18, 12, 6, src_mask=mask)
    x = (280 + 20 * rng.standard_normal((3, 2, 18, 36))).astype(np.float32)
    x[..., mask == 0] = np.nan
    x[2, 1] = np.nan
    case_2d("con_masked_f32", w, x, ("time", "plev", "lat", "lon"))
    case_2d("con_masked_f64", w, x.astype(np.float64), ("time", "plev", "lat", "lon"))

    # unsorted links with duplicates, negative weights and extra remap_matrix columns (bicubic-like);
    # unstructured target rank 1
    n_src, n_dst = 300, 70
    nl = 900
    src = rng.integers(1, n_src + 1, nl).astype(np.int32)
    dst = rng.integers(1, n_dst + 1, nl).astype(np.int32)
    src[100:140], dst[100:140] = src[:40], dst[:40]
    rm = rng.random((nl, 4)) - 0.2
    w = synth.CdoWeights({
        "src_address": src, "dst_address": dst, "remap_matrix": rm,
        "src_grid_imask": (rng.random(n_src) > 0.3).astype(np.int32), "dst_grid_imask": np.ones(n_dst, np.int32),
        "dst_grid_frac": rng.random(n_dst), "src_grid_dims": np.array([n_src], np.int32),
        "dst_grid_dims": np.array([n_dst], np.int32),
        "dst_grid_center_lat": rng.random(n_dst), "dst_grid_center_lon": rng.random(n_dst)},
        attrs={"source_grid": "unstructured", "dest_grid": "unstructured"})
    x = rng.standard_normal((5, n_src)).astype(np.float64) * 50
    x[rng.random(x.shape) < 0.1] = np.nan
    x[4, 7] = 5e19                      # legit value above the threshold -> NaN (regrid.py:570)
    case_2d("unsorted_dups_f64", w, x, ("time", "cell"))

    # 3-D ocean-like, level-varying mask
    w = synth.ocean3d_weights(36, 18, 12, 6, n_levels=6, seed=3)
    x = (10 + rng.standard_normal((3, 6, 18, 36))).astype(np.float32)
    x[:, w["src_grid_imask"].reshape(6, 18, 36) == 0] = np.nan
    case_3d("ocean3d_f32", w, x)


def nan_variation_cases():
    """detect_nan_variation_dims (smmregrid/util.py:57-85), the reference's own function, on
    fields whose missing-value pattern varies along none / one / two of the extra dimensions."""
    from smmregrid.util import detect_nan_variation_dims as ref_detect
    rng = np.random.default_rng(77)
    dims = ("time", "lev", "member", "lat", "lon")
    out = {}
    for i, dt in enumerate((np.float32, np.float64)):
        base = rng.standard_normal((3, 5, 2, 6, 8)).astype(dt)
        fields = {}
        f = base.copy(); f[:, :, :, 2:4, 3:6] = np.nan                      # static land: no variation
        fields["static"] = f
        f = base.copy()
        for lev in range(5):                                                   # bathymetry: varies along lev only
            f[:, lev, :, :lev + 1, :] = np.nan
        fields["lev"] = f
        f = fields["lev"].copy(); f[:, :, 1, 5, 7] = np.nan                 # ... and one member differs
        fields["lev_member"] = f
        f = base.copy(); f[1:, 2, 0, 0, 0] = np.nan                          # only later time steps: first step clean
        fields["later_steps_only"] = f
        f = base.copy(); f[:, 3, :, 1, 1] = np.inf                           # inf is not null
        fields["inf_only"] = f
        for name, f in fields.items():
            da = xr.DataArray(f, dims=dims)
            got = ref_detect(da, time_dim=["time"], check_dims=["lev", "member"])
            out[f"{name}_{np.dtype(dt).name}"] = f
            out[f"{name}_{np.dtype(dt).name}_dims"] = np.array([dims.index(d) for d in got], np.int64)
            # no time dimension present: the whole field is inspected
            got2 = ref_detect(xr.DataArray(f[0], dims=dims[1:]), time_dim=["time"], check_dims=["lev", "member"])
            assert got2 == got
    np.savez_compressed(os.path.join(HERE, "nan_variation.npz"), **out)
    print("nan_variation.npz:", {k: v.tolist() for k, v in out.items() if k.endswith("_dims")})


def real_data_case():
    """The one data file of the reference's own test suite that is readable here (netCDF-3
    classic): tests/data/tas-healpix2.nc, 12 time steps of near-surface temperature on a HEALPix
    nside-32 NESTED grid (``cdo setgrid,hp32b_nest``), with the pixel centres.  Stored as a small
    fixture so that the GPU box can regrid real data (it cannot read /root/reference)."""
    from scipy.io import netcdf_file
    with netcdf_file("/root/reference/tests/data/tas-healpix2.nc", "r", mmap=False) as nc:
        tas = np.array(nc.variables["tas"][...]).astype(np.float32)
        lat = np.array(nc.variables["lat"][...]).astype(np.float64)
        lon = np.array(nc.variables["lon"][...]).astype(np.float64)
        time = np.array(nc.variables["time"][...]).astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "tas_healpix2.npz"), tas=tas, lat_rad=lat, lon_rad=lon, time=time)
    print("tas_healpix2.npz:", tas.shape, tas.dtype, float(tas.min()), float(tas.max()))


def netcdf4_cases():
    """netCDF-4 (HDF5) data files of the reference's test suite, for the package's own reader
    (smmregrid_b200/nc4.py; the image has neither netCDF4 nor h5py):

    * the two smallest files travel as they are (tests/golden/data/healpix_0.nc, regional.nc:
      written by the netCDF library through CDO / NCO -- superblock 2, version-2 object headers,
      dense attributes, chunked + deflated + shuffled variables);
    * ua-ipsl.nc (4 MB: eastward wind on 19 pressure levels, 143 x 144 lat-lon with points at the
      poles, 1e20 under the orography: REAL level-dependent masks) is cut to one time step and its
      six lowest levels, values untouched: tests/golden/ua_ipsl.npz;
    * regional.nc was cut by `cdo sellonlatbox` from r360x180.nc: the reader must return the
      same bits from both files (checked here, where both are readable)."""
    import shutil
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from smmregrid_b200 import nc4
    data = "/root/reference/tests/data/"
    os.makedirs(os.path.join(HERE, "data"), exist_ok=True)
    for name in ("healpix_0.nc", "regional.nc"):
        shutil.copyfile(data + name, os.path.join(HERE, "data", name))
        os.chmod(os.path.join(HERE, "data", name), 0o644)
    with nc4.File(data + "r360x180.nc") as g, nc4.File(data + "regional.nc") as r:
        la, lo = g.variables["lat"][...], g.variables["lon"][...]
        i0 = int(np.where(la == r.variables["lat"][0])[0][0])
        j0 = int(np.where(lo == r.variables["lon"][0])[0][0])
        assert np.array_equal(g.variables["pr"][...][0, i0:i0 + 90, j0:j0 + 61], r.variables["pr"][...][0])
    with nc4.File(data + "ua-ipsl.nc") as f:
        ua = f.variables["ua"]
        assert ua.dims == ("time", "plev", "lat", "lon")
        np.savez_compressed(os.path.join(HERE, "ua_ipsl.npz"), ua=ua[...][0, :6], plev=f.variables["plev"][:6],
                            lat=f.variables["lat"][...], lon=f.variables["lon"][...],
                            fill=np.float32(ua.attrs["_FillValue"]))
        print("ua_ipsl.npz:", ua[...][0, :6].shape, [(int((ua[...][0, l] == 1e20).sum())) for l in range(6)])


def _coord_record(da, key):
    c = da.coords[key]
    return {"dims": list(c.dims), "values": np.asarray(c.data), "attrs": dict(c.attrs)}


def dressing_cases():
    """The xarray side of Regridder.apply_weights (regrid.py:572-626), the reference's own code:
    output dims, kept coordinates, target lat/lon (remove_degenerate_axes, degrees, rounding),
    the i/j -> lat/lon swap for regular targets, coordinate and variable attributes -- for a
    regular, a curvilinear and an unstructured (rank-1) target."""
    import json
    rng = np.random.default_rng(5)
    out = {}

    def record(tag, w, x, dims, coords, attrs):
        ds = to_dataset(w)
        matrix = ref_weights.compute_weights_matrix(ds)
        ds2 = ref_weights.mask_weights(ds, matrix)
        masked = bool(ref_weights.check_mask(ds2))
        da = xr.DataArray(x, dims=dims, coords=coords, name="tos", attrs=dict(attrs))
        res = RefRegridder.apply_weights(_Self(0.5), da, ds2, weights_matrix=matrix, masked=masked,
                                         horizontal_dims=["lon", "lat", "cell", "i", "j"])
        meta = {"dims": list(res.dims), "name": res.name, "attrs": dict(res.attrs), "coords": {}}
        for k in res.coords:
            rec = _coord_record(res, k)
            out[f"{tag}_coord_{k}"] = rec["values"]
            meta["coords"][k] = {"dims": rec["dims"], "attrs": rec["attrs"]}
        out[f"{tag}_y"] = np.asarray(res.data)
        out[f"{tag}_x"] = x
        out[f"{tag}_meta"] = np.asarray(json.dumps(meta))
        for k, v in inputs_of(w).items():
            out[f"{tag}_{k}"] = v
        out[f"{tag}_in_dst_grid_center_lat"] = w["dst_grid_center_lat"]
        out[f"{tag}_in_dst_grid_center_lon"] = w["dst_grid_center_lon"]
        print(f"dressing {tag}: dims={meta['dims']} coords={ {k: v['dims'] for k, v in meta['coords'].items()} }")

    attrs = {"units": "K", "long_name": "temperature", "CDI_grid_type": "curvilinear"}
    # (a) regular target: lat/lon collapse to 1-D and replace i/j
    w = synth.conservative_latlon(36, 18, 12, 6)
    x = (280 + 20 * rng.standard_normal((3, 2, 18, 36))).astype(np.float32)
    record("regular", w, x, ("time", "plev", "lat", "lon"),
           {"time": np.array([10.0, 20.0, 30.0]), "plev": np.array([85000.0, 50000.0]),
            "lat": np.linspace(-85, 85, 18), "lon": np.arange(36) * 10.0}, attrs)
    # (b) curvilinear target: 2-D lat/lon stay on (i, j)
    v = dict(w.vars)
    jj, ii = np.meshgrid(np.arange(6), np.arange(12), indexing="ij")
    v["dst_grid_center_lat"] = np.deg2rad(-75 + 30.0 * jj + 1.5 * ii).ravel()
    v["dst_grid_center_lon"] = np.deg2rad(15 + 30.0 * ii + 2.0 * jj).ravel()
    record("curvilinear", synth.CdoWeights(v, attrs=w.attrs), x, ("time", "plev", "lat", "lon"),
           {"time": np.array([10.0, 20.0, 30.0]), "plev": np.array([85000.0, 50000.0])}, attrs)
    # (c) unstructured source and target (rank 1): lat/lon on 'cell'
    n_src, n_dst, nl = 300, 70, 600
    w1 = synth.CdoWeights({
        "src_address": rng.integers(1, n_src + 1, nl).astype(np.int32),
        "dst_address": rng.integers(1, n_dst + 1, nl).astype(np.int32), "remap_matrix": rng.random((nl, 1)),
        "src_grid_imask": np.ones(n_src, np.int32), "dst_grid_imask": np.ones(n_dst, np.int32),
        "dst_grid_frac": np.ones(n_dst), "src_grid_dims": np.array([n_src], np.int32),
        "dst_grid_dims": np.array([n_dst], np.int32),
        "dst_grid_center_lat": rng.uniform(-1.5, 1.5, n_dst), "dst_grid_center_lon": rng.uniform(0, 6.28, n_dst)},
        attrs={"source_grid": "unstructured", "dest_grid": "unstructured"})
    record("cell", w1, rng.standard_normal((4, n_src)), ("time", "cell"), {"time": np.arange(4.0)}, {"units": "1"})
    np.savez_compressed(os.path.join(HERE, "dressing.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "main"):
        main()
    if which in ("all", "nanvar"):
        nan_variation_cases()
    if which in ("all", "real"):
        real_data_case()
    if which in ("all", "dressing"):
        dressing_cases()
    if which in ("all", "netcdf4"):
        netcdf4_cases()
