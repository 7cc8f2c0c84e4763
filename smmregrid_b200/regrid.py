"""Regridder -- host-side mirror of ``smmregrid/regrid.py`` for the weight-application path.

Same class / method names, argument meaning and error behaviour as the reference for the
part of the API that sits on the hot path (``Regridder(weights=...)``, ``regrid``,
``regrid_array``, ``regrid2d``, ``regrid3d``, ``apply_weights``, module-level ``regrid``).
Weight *generation* (``Regridder(source_grid, target_grid)`` -> CDO subprocess) is out of
scope here and raises ``NotImplementedError``: generate the weights with CDO / the reference's
``CdoGenerate`` and pass them in.

Data may be a numpy array, a torch tensor (CPU or CUDA) or -- when xarray is installed -- a
``DataArray`` / ``Dataset``.  For bare arrays the horizontal axes are the trailing ones
(``[..., nlat, nlon]`` or ``[..., ncell]``) and, for 3-D weights, ``mask_dim`` is the axis
just before them unless ``level_axis`` says otherwise.  The numeric core always runs in the
sm_100a kernels through the C ABI; there is no CPU path.
"""
from __future__ import annotations

import ctypes
import logging
import math
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .weights import (CdoWeights, WeightsMatrix, check_mask, compute_weights_matrix,
                      compute_weights_matrix3d, mask_weights, _device_index)

DEFAULT_AREA_MIN = 0.5  # default minimum area for conservative remapping (regrid.py:49)
LEVEL_TOLERANCE = 1e-3  # regrid.py:390


def _logger(level):
    log = logging.getLogger("smmregrid.Regrid")
    log.setLevel(getattr(logging, str(level).upper(), logging.WARNING))
    return log


def _torch():
    import torch
    return torch


def _np_dtype_code(dt) -> int:
    dt = np.dtype(dt)
    if dt == np.float32:
        return _lib.SMM_F32
    if dt == np.float64:
        return _lib.SMM_F64
    raise TypeError(f"only float32/float64 data can be regridded, got {dt}")


def select_level(weight_levels, lev, name="lev") -> int:
    """Nearest weights level within 1e-3, else ValueError (``regrid.py:386-395``)."""
    wl = np.asarray(weight_levels, dtype=np.float64)
    d = np.abs(wl - float(lev))
    widx = int(np.argmin(d))
    if not d[widx] <= LEVEL_TOLERANCE:
        raise ValueError(f"{lev} not found in mask_dim {name}. Available levels: {list(wl)}")
    return widx


class Regridder(object):
    """Main regridding class (reference: ``smmregrid/regrid.py:52``), weights-initialised."""

    def __init__(self, source_grid=None, target_grid=None, weights=None,
                 method='con', remap_area_min=DEFAULT_AREA_MIN, transpose=True, mask_dim=None,
                 vertical_dim=None, horizontal_dims=None, cdo_extra=None, cdo_options=None,
                 check_nan=False, cdo='cdo', loglevel='WARNING',
                 device=None, out_dtype=None, renormalize=None):
        if (source_grid is None or target_grid is None) and (weights is None):
            raise ValueError("Either weights or source_grid/target_grid must be supplied")
        if vertical_dim is not None and mask_dim is None:      # deprecated alias (regrid.py:107)
            import warnings
            warnings.warn("vertical_dim is deprecated, use mask_dim", DeprecationWarning, stacklevel=2)
            mask_dim = vertical_dim
        self.loggy = _logger(loglevel)
        self.loglevel = loglevel
        self.transpose = transpose
        self.remap_area_min = float(remap_area_min)
        if self.remap_area_min < 0.0 or self.remap_area_min > 1.0:
            raise ValueError('The remap_area_min provided must be between 0.0 and 1.0')
        if weights is None:
            raise NotImplementedError(
                "smmregrid_b200 accelerates weight *application*; generate weights with CDO "
                "(reference CdoGenerate) and pass weights=")
        self.init_mode = 'weights'
        self.device = _device_index(device)
        self.out_dtype = None if out_dtype is None else np.dtype(out_dtype)

        w = CdoWeights.from_any(weights, mask_dim=mask_dim)
        self.mask_dim = w.mask_dim if w.is3d else None
        if w.is3d:
            self.weights_matrix = compute_weights_matrix3d(w, self.mask_dim, device=self.device)
        else:
            self.weights_matrix = compute_weights_matrix(w, device=self.device)
        # destination mask: precomputed flag or compute it now (regrid.py:196-203)
        if "dst_grid_masked" in w:
            self.masked = np.asarray(w["dst_grid_masked"]).astype(bool)
            if not w.is3d:
                self.masked = bool(self.masked)
        else:
            w = mask_weights(w, self.weights_matrix, self.mask_dim)
            self.masked = check_mask(w, self.mask_dim)
        self.weights = w
        # extension beyond the reference (off by default): time-varying missing values are
        # excluded and the remaining weights renormalised, NaN below `renormalize` valid fraction
        self.renormalize = renormalize
        n = w.sizes
        self.n_src, self.n_dst = n["src_grid_size"], n["dst_grid_size"]
        self.src_grid_shape = tuple(int(v) for v in np.atleast_1d(w["src_grid_dims"]))[::-1] \
            if "src_grid_dims" in w else (self.n_src,)
        dgd = tuple(int(v) for v in np.atleast_1d(w["dst_grid_dims"])) if "dst_grid_dims" in w else (self.n_dst,)
        rank = len(dgd)
        if rank == 2:                      # regrid.py:572-574
            self.tgt_shape = (dgd[1], dgd[0])
        elif rank == 1:
            self.tgt_shape = (dgd[0],)
        else:
            raise ValueError('Unknown dimensional target grid')

    # ------------------------------------------------------------------ public API

    def regrid(self, source_data, level_axis: Optional[int] = None, levels: Optional[Sequence[float]] = None):
        """Regrid an array / DataArray / Dataset (reference: ``regrid.py:233-271``)."""
        xr = _maybe_xarray()
        if xr is not None and isinstance(source_data, xr.Dataset):
            out = source_data.map(self.regrid_array, keep_attrs=True)
            degen = [v for v in out.data_vars if out[v].dims == ()]
            return out.drop_vars(degen)
        if xr is not None and isinstance(source_data, xr.DataArray):
            return self.regrid_array(source_data)
        if isinstance(source_data, np.ndarray) or _is_tensor(source_data):
            return self.regrid_array(source_data, level_axis=level_axis, levels=levels)
        if isinstance(source_data, dict):
            # plain-array counterpart of Dataset.map (regrid.py:262): one entry per variable;
            # bounds variables are dropped except time bounds (regrid.py:482-490)
            out = {}
            for name, arr in source_data.items():
                if any(sub in str(name) for sub in ("bnds", "bounds", "vertices")):
                    if 'time' in str(name):
                        out[name] = arr
                    continue
                out[name] = self.regrid_array(arr, level_axis=level_axis, levels=levels)
            return out
        raise TypeError('The object provided is not a Xarray object!')

    def regrid_array(self, source_data, level_axis=None, levels=None):
        """2-D or 3-D dispatch on the presence of ``mask_dim`` (reference: ``regrid.py:273-312``)."""
        xr = _maybe_xarray()
        if xr is not None and isinstance(source_data, xr.DataArray):
            return self._regrid_dataarray(source_data)
        if self.mask_dim:
            return self.regrid3d(source_data, level_axis=level_axis, levels=levels)
        return self.regrid2d(source_data)

    def regrid2d(self, source_data):
        """Single apply with the 2-D operator (reference: ``regrid.py:429-456``)."""
        return self.apply_weights(source_data, self.weights, weights_matrix=self.weights_matrix,
                                  masked=self.masked)

    def regrid3d(self, source_data, level_axis=None, levels=None):
        """Per-level apply for level-varying masks (reference: ``regrid.py:339-427``), issued as
        one grouped launch that writes every level at its final position."""
        torch = _torch()
        nh = self._n_horizontal_axes(source_data.shape)
        nkept = source_data.ndim - nh
        if nkept < 1:
            raise ValueError(f"3-D weights need a {self.mask_dim} axis in the data")
        la = nkept - 1 if level_axis is None else (level_axis % source_data.ndim)
        if la >= nkept:
            raise ValueError("level_axis must be one of the non-horizontal axes")
        Ld = source_data.shape[la]
        wl = self.weights.levels
        if levels is None:
            if Ld != len(wl):
                raise ValueError(f"data has {Ld} levels, weights {len(wl)}: pass levels= to select")
            levels = wl
        if len(levels) != Ld:
            raise ValueError("levels must have one value per data level")
        widx = np.array([select_level(wl, lev, self.mask_dim) for lev in levels], dtype=np.int32)
        self.weights_matrix.set_renormalize(self.renormalize)
        masked = np.ascontiguousarray(np.asarray(self.masked)[widx], dtype=np.uint8)

        was_numpy = isinstance(source_data, np.ndarray)
        x = torch.from_numpy(np.ascontiguousarray(source_data)) if was_numpy else source_data
        host_in = not x.is_cuda
        if host_in:
            return self._regrid3d_host(x, la, nkept, Ld, widx, masked, was_numpy)
        dev = torch.device("cuda", self.device)
        # canonical layout [T, L, n_src]: level axis last among the kept ones
        x = torch.movedim(x, la, nkept - 1)
        kept_shape = tuple(x.shape[:nkept - 1])
        T = int(np.prod(kept_shape)) if kept_shape else 1
        x = x.reshape(T, Ld, self.n_src).contiguous()
        ydt = self._out_torch_dtype(x.dtype)
        if self.transpose:
            y = torch.empty((T, Ld, self.n_dst), dtype=ydt, device=dev)
            ybs, yls = Ld * self.n_dst, self.n_dst
        else:
            y = torch.empty((Ld, T, self.n_dst), dtype=ydt, device=dev)
            ybs, yls = self.n_dst, T * self.n_dst
        lib = _lib.load()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.smm_apply_levels(
                self.weights_matrix.handle, Ld, widx.ctypes.data_as(ctypes.c_void_p),
                ctypes.c_void_p(x.data_ptr()), _np_dtype_code(_np_of(x.dtype)), T,
                Ld * self.n_src, self.n_src,
                ctypes.c_void_p(y.data_ptr()), _np_dtype_code(_np_of(y.dtype)), ybs, yls,
                masked.ctypes.data_as(ctypes.c_void_p), self.remap_area_min, ctypes.c_void_p(stream)))
        if self.transpose:      # [kept..., L, tgt...]  (regrid.py:420-427)
            y = y.reshape(kept_shape + (Ld,) + self.tgt_shape)
        else:                   # concat on a new leading axis (regrid.py:410)
            y = y.reshape((Ld,) + kept_shape + self.tgt_shape)
        return y

    def _regrid3d_host(self, x, la, nkept, Ld, widx, masked, was_numpy):
        """regrid3d for host data: streamed through the library in chunks of the leading kept
        axes (pinned bounce buffers for pageable arrays), never holding the whole field on the
        device."""
        torch = _torch()
        x = torch.movedim(x, la, nkept - 1)                    # canonical [T, L, n_src]
        kept_shape = tuple(x.shape[:nkept - 1])
        T = int(np.prod(kept_shape)) if kept_shape else 1
        x = x.reshape(T, Ld, self.n_src).contiguous()
        y = torch.empty((T, Ld, self.n_dst), dtype=self._out_torch_dtype(x.dtype), pin_memory=True)
        _lib.check(_lib.load().smm_apply_levels_host(
            self.weights_matrix.handle, Ld, widx.ctypes.data_as(ctypes.c_void_p),
            ctypes.c_void_p(x.data_ptr()), _np_dtype_code(_np_of(x.dtype)), T,
            ctypes.c_void_p(y.data_ptr()), _np_dtype_code(_np_of(y.dtype)),
            masked.ctypes.data_as(ctypes.c_void_p), self.remap_area_min, 0))
        if self.transpose:      # [kept..., L, tgt...]  (regrid.py:420-427)
            y = y.reshape(kept_shape + (Ld,) + self.tgt_shape)
        else:                   # level axis first, as xarray.concat leaves it (regrid.py:410)
            y = torch.movedim(y, 1, 0).reshape((Ld,) + kept_shape + self.tgt_shape)
        return y.numpy() if was_numpy else y

    def apply_weights(self, source_data, weights=None, weights_matrix=None, masked=True,
                      horizontal_dims=None, level: int = 0):
        """Numeric core of the reference's ``apply_weights`` (``regrid.py:536-584``):
        reshape to ``[kept, n_src]``, non-finite -> 1e20, sparse matmul, destination mask,
        ``dst_grid_frac < remap_area_min``, ``> 1e19 -> NaN``, reshape to the target grid."""
        torch = _torch()
        wm = weights_matrix if weights_matrix is not None else self.weights_matrix
        parent, lvl = (wm.parent, wm.level) if hasattr(wm, "parent") else (wm, level)
        nh = self._n_horizontal_axes(source_data.shape)
        kept_shape = tuple(source_data.shape[:source_data.ndim - nh])
        B = int(np.prod(kept_shape)) if kept_shape else 1
        lib = _lib.load()
        masked = 1 if bool(masked) else 0
        parent.set_renormalize(self.renormalize)

        if isinstance(source_data, np.ndarray) or not source_data.is_cuda:
            # host data: chunked H2D -> kernel -> D2H pipeline inside the library
            was_numpy = isinstance(source_data, np.ndarray)
            xt = torch.from_numpy(np.ascontiguousarray(source_data)) if was_numpy else source_data.contiguous()
            xt = xt.reshape(B, self.n_src)
            ydt = self._out_torch_dtype(xt.dtype)
            y = torch.empty((B, self.n_dst), dtype=ydt, pin_memory=True)
            _lib.check(lib.smm_apply_host(
                parent.handle, lvl, ctypes.c_void_p(xt.data_ptr()), _np_dtype_code(_np_of(xt.dtype)), B,
                self.n_src, ctypes.c_void_p(y.data_ptr()), _np_dtype_code(_np_of(y.dtype)), self.n_dst,
                masked, self.remap_area_min, 0))
            y = y.reshape(kept_shape + self.tgt_shape)
            return y.numpy() if was_numpy else y

        dev = source_data.device
        x = source_data.reshape(B, self.n_src)
        if not x.is_contiguous():
            x = x.contiguous()
        y = torch.empty((B, self.n_dst), dtype=self._out_torch_dtype(x.dtype), device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.smm_apply(
                parent.handle, lvl, ctypes.c_void_p(x.data_ptr()), _np_dtype_code(_np_of(x.dtype)), B,
                self.n_src, ctypes.c_void_p(y.data_ptr()), _np_dtype_code(_np_of(y.dtype)), self.n_dst,
                masked, self.remap_area_min, ctypes.c_void_p(stream)))
        return y.reshape(kept_shape + self.tgt_shape)

    # ------------------------------------------------------------------ helpers

    def _out_torch_dtype(self, x_dtype):
        """Reference output dtype is result_type(data, float64 weights) = float64
        (SURVEY §0.3); ``out_dtype=float32`` rounds once at the store."""
        torch = _torch()
        if self.out_dtype is None:
            return torch.float64
        return torch.float32 if self.out_dtype == np.float32 else torch.float64

    def _n_horizontal_axes(self, shape) -> int:
        shape = tuple(int(s) for s in shape)
        gs = self.src_grid_shape
        if len(shape) >= len(gs) and shape[-len(gs):] == gs:
            return len(gs)
        if len(shape) >= 1 and shape[-1] == self.n_src:
            return 1
        raise KeyError('Dimensions mismatch')          # regrid.py:524

    def _regrid_dataarray(self, da):
        """xarray front end: horizontal dims = the trailing dims matching the source grid."""
        xr = _maybe_xarray()
        if any(s in str(da.name) for s in ("bnds", "bounds", "vertices")):   # regrid.py:482-490
            return da if 'time' in str(da.name) else xr.DataArray(data=None)
        values = np.asarray(da.values)
        nh = self._n_horizontal_axes(values.shape)
        kept_dims = list(da.dims[:values.ndim - nh])
        if self.mask_dim and self.mask_dim in kept_dims:
            la = kept_dims.index(self.mask_dim)
            out = self.regrid3d(values, level_axis=la, levels=da[self.mask_dim].values)
            kept_dims = [d for d in kept_dims if d != self.mask_dim]
            kept_dims = kept_dims + [self.mask_dim] if self.transpose else [self.mask_dim] + kept_dims
        else:
            out = self.regrid2d(values)
        tgt_dims = ["lat", "lon"] if len(self.tgt_shape) == 2 else ["cell"]
        coords = {k: v for k, v in da.coords.items() if set(v.dims).issubset(kept_dims)}
        res = xr.DataArray(out, dims=kept_dims + tgt_dims, coords=coords, name=da.name, attrs=dict(da.attrs))
        res.attrs.pop('CDI_grid_type', None)
        w = self.weights
        if "dst_grid_center_lat" in w and len(self.tgt_shape) == 2:
            scale = 180.0 / math.pi
            lat = np.round(np.asarray(w["dst_grid_center_lat"]).reshape(self.tgt_shape)[:, 0] * scale, 10)
            lon = np.round(np.asarray(w["dst_grid_center_lon"]).reshape(self.tgt_shape)[0, :] * scale, 10)
            res = res.assign_coords(lat=("lat", lat), lon=("lon", lon))
        return res


def regrid(source_data, target_grid=None, weights=None, transpose=True, cdo='cdo', **kwargs):
    """One-shot helper (reference: ``regrid.py:656-677``)."""
    regridder = Regridder(source_data if weights is None else None, target_grid=target_grid,
                          weights=weights, cdo=cdo, transpose=transpose, **kwargs)
    return regridder.regrid(source_data)


def _maybe_xarray():
    try:
        import xarray
        return xarray
    except ImportError:
        return None


def _is_tensor(obj) -> bool:
    try:
        return isinstance(obj, _torch().Tensor)
    except ImportError:
        return False


def _np_of(torch_dtype):
    torch = _torch()
    if torch_dtype == torch.float32:
        return np.float32
    if torch_dtype == torch.float64:
        return np.float64
    raise TypeError(f"only float32/float64 data can be regridded, got {torch_dtype}")
