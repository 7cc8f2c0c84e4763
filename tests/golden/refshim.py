"""Stand-ins for the third-party packages the reference imports, so that the reference's OWN
source files (/root/reference/smmregrid/weights.py and regrid.py) can be executed in the
build container, where xarray / dask / sparse are not installable (no network).

TEST INFRASTRUCTURE, used only by ``make_golden.py`` (never on the GPU box, never by the
product).  What is real and what is restated:

* REAL: every line of ``smmregrid/weights.py`` and of ``Regridder.apply_weights``
  (``smmregrid/regrid.py:458-628``) -- addresses ``-1``, column 0 of remap_matrix, the
  matrix shape, reshape to ``[kept, n_src]``, ``fix_invalid``/``filled``, ``tensordot``,
  the three ``where`` passes and their order, the target reshape.
* numpy-backed stand-ins: ``dask.array.X`` -> ``numpy.X`` (dask documents these as blockwise
  numpy calls), ``dask.delayed``/``from_delayed`` -> eager evaluation, and a tiny
  ``xarray.DataArray``/``Dataset`` holding numpy arrays + dims.
* RESTATED (the part that stays "parity unpinned"): ``sparse.COO`` construction (sort by
  linear index, sum duplicates) and the ``ndarray . COO`` matmul loop of pydata/sparse
  (``for each batch row, for each stored element in order: out[b, col] += a[b, row] * v``).
"""
from __future__ import annotations

import sys
import types

import numpy as np


# ----------------------------------------------------------------------------- sparse

class COO:
    def __init__(self, coords, data, shape=None, **_kw):
        coords = np.asarray(coords).astype(np.int64)
        data = np.asarray(data)
        self.shape = tuple(int(s) for s in shape)
        lin = coords[0] * self.shape[1] + coords[1]
        order = np.argsort(lin, kind="stable")
        lin, data = lin[order], data[order]
        if lin.size:
            first = np.concatenate(([True], lin[1:] != lin[:-1]))
            starts = np.flatnonzero(first)
            data = np.add.reduceat(data, starts)
            lin = lin[starts]
        self.coords = np.stack([lin // self.shape[1], lin % self.shape[1]])
        self.data = data
        self.dtype = data.dtype
        self.ndim = 2

    @property
    def nnz(self):
        return self.data.size


def _dot_ndarray_coo(a, coo):
    """pydata/sparse `_dot_ndarray_coo`: loop over stored elements in COO order per row of a."""
    import numba

    @numba.njit(cache=False)
    def kern(a2, rows, cols, data, out):
        for i in range(a2.shape[0]):
            for k in range(data.shape[0]):
                out[i, cols[k]] += a2[i, rows[k]] * data[k]

    a2 = np.ascontiguousarray(a).reshape(-1, a.shape[-1])
    out = np.zeros((a2.shape[0], coo.shape[1]), dtype=np.result_type(a2.dtype, coo.dtype))
    kern(a2, coo.coords[0], coo.coords[1], coo.data, out)
    return out.reshape(a.shape[:-1] + (coo.shape[1],))


# ------------------------------------------------------------------------------- dask

class _DaskArray:        # only used in isinstance checks (regrid.py:538)
    pass


def _tensordot(a, b, axes=1):
    assert axes == 1
    a = np.asarray(a)
    if isinstance(b, COO):
        return _dot_ndarray_coo(a, b)
    return np.tensordot(a, b, axes=1)


def _delayed(fn):
    return fn


def _from_delayed(value, shape=None, dtype=None, **_kw):
    return value


# ----------------------------------------------------------------------------- xarray

class DataArray:
    def __init__(self, data=None, dims=None, coords=None, name=None, attrs=None):
        self.data = None if data is None else data if isinstance(data, (COO, FakeDaskArray)) else np.asarray(data)
        self.dims = tuple(dims) if dims is not None else ()
        self.name = name
        self.attrs = dict(attrs or {})
        self.coords = _Coords(self)
        for k, v in (coords or {}).items():
            self.coords[k] = v

    # numpy-ish surface
    @property
    def values(self): return self.data
    @property
    def shape(self): return self.data.shape
    @property
    def ndim(self): return self.data.ndim
    @property
    def dtype(self): return self.data.dtype
    @property
    def size(self): return self.data.size
    @property
    def sizes(self): return dict(zip(self.dims, self.data.shape))
    @property
    def chunks(self): return getattr(self.data, "chunks", None)      # None unless dask-backed, like xarray
    def __array__(self, dtype=None, copy=None): return np.asarray(self.data, dtype=dtype)
    def __len__(self): return len(self.data)
    def _new(self, data, dims=None): return DataArray(data, self.dims if dims is None else dims, name=self.name, attrs=self.attrs)
    def __sub__(self, o): return self._new(self.data - np.asarray(o))
    def __mul__(self, o): return self._new(self.data * np.asarray(o))
    def __eq__(self, o): return self._new(self.data == np.asarray(o))
    def __invert__(self): return self._new(~self.data)
    def __lt__(self, o): return self._new(self.data < np.asarray(o))
    def round(self, decimals=0, out=None): return self._new(np.round(self.data, decimals))
    def compute(self): return self.data if self.data.ndim else self.data[()]

    def __getitem__(self, key):
        key = key if isinstance(key, tuple) else (key,)
        dims = [d for d, k in zip(self.dims, key + (slice(None),) * (len(self.dims) - len(key)))
                if not isinstance(k, (int, np.integer))]
        return self._new(self.data[key], dims)

    def __getattr__(self, name):
        coords = self.__dict__.get("coords")
        if coords is not None and name in coords:
            return coords[name]
        raise AttributeError(name)

    def _reduce(self, fn, dim):
        if dim is None:
            return self._new(fn(self.data), ())
        dim = [dim] if isinstance(dim, str) else list(dim)
        ax = tuple(self.dims.index(d) for d in dim)
        return self._new(fn(self.data, axis=ax), [d for d in self.dims if d not in dim])
    def max(self, dim=None): return self._reduce(np.max, dim)
    def min(self, dim=None): return self._reduce(np.min, dim)
    def mean(self, dim=None): return self._reduce(np.mean, dim)
    def all(self, dim=None): return self._reduce(np.all, dim)

    # the surface util.py:detect_nan_variation_dims uses (util.py:69-80)
    def isnull(self): return self._new(np.isnan(self.data))
    def astype(self, dt): return self._new(self.data.astype(dt))
    def diff(self, dim): return self._new(np.diff(self.data, axis=self.dims.index(dim)))
    def any(self, dim=None): return self._reduce(np.any, dim)
    def sum(self, dim=None): return self._reduce(np.sum, dim)

    def isel(self, indexers=None, **kw):
        kw.pop("drop", None)
        indexers = dict(indexers or {}, **kw)
        key = tuple(indexers.get(d, slice(None)) for d in self.dims)
        dims = [d for d in self.dims if not isinstance(indexers.get(d, slice(None)), (int, np.integer))]
        out = self._new(self.data[key], dims)
        for k, v in self.coords.items():            # coordinates follow the selection
            if set(v.dims) <= set(dims):
                ck = tuple(indexers.get(d, slice(None)) for d in v.dims)
                out.coords[k] = DataArray(v.data[ck], v.dims, attrs=v.attrs)
        return out

    def transpose(self, *dims):
        order = [self.dims.index(d) for d in dims]
        out = DataArray(np.transpose(self.data, order), dims, name=self.name, attrs=self.attrs)
        for k, v in self.coords.items():
            out.coords[k] = v
        return out

    def swap_dims(self, mapping):
        out = DataArray(self.data, [mapping.get(d, d) for d in self.dims], name=self.name, attrs=self.attrs)
        for k, v in self.coords.items():
            out.coords[k] = DataArray(v.data, [mapping.get(d, d) for d in v.dims], attrs=v.attrs)
        return out


class _Coords(dict):
    def __init__(self, owner):
        super().__init__()
        self._owner = owner

    def __setitem__(self, k, v):
        if isinstance(v, tuple):                    # xarray's (dims, data[, attrs]) form
            dims = (v[0],) if isinstance(v[0], str) else tuple(v[0])
            v = DataArray(np.asarray(v[1]), dims, attrs=v[2] if len(v) > 2 else None)
        elif not isinstance(v, DataArray):
            v = DataArray(np.asarray(v), (k,) if np.ndim(v) == 1 else ())
        super().__setitem__(k, v)


class Dataset:
    def __init__(self, data_vars=None, attrs=None):
        self.variables = {}
        self.attrs = dict(attrs or {})
        for k, v in (data_vars or {}).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, tuple):
            v = DataArray(v[1], v[0], name=k)
        if v.name is None:
            v.name = k
        self.variables[k] = v

    def __getitem__(self, k): return self.variables[k]
    def __contains__(self, k): return k in self.variables
    def __getattr__(self, k):
        v = self.__dict__.get("variables", {})
        if k in v:
            return v[k]
        raise AttributeError(k)

    @property
    def sizes(self):
        out = {}
        for v in self.variables.values():
            out.update(v.sizes)
        return out

    @property
    def dims(self): return self.sizes

    @property
    def data_vars(self): return self.variables

    def map(self, func, keep_attrs=False):
        out = Dataset(attrs=self.attrs if keep_attrs else None)
        for k, v in self.variables.items():
            r = func(v)
            r.name = k
            out.variables[k] = r
        return out

    def drop_vars(self, names):
        return Dataset({k: v for k, v in self.variables.items() if k not in set(names)}, self.attrs)

    def isel(self, indexers=None, **kw):
        indexers = dict(indexers or {}, **kw)
        return Dataset({k: v.isel({d: i for d, i in indexers.items() if d in v.dims})
                        for k, v in self.variables.items()}, self.attrs)

    def assign(self, **kw):
        out = Dataset(dict(self.variables), self.attrs)
        for k, v in kw.items():
            out[k] = v
        return out


# ------------------------------------------------------------- lazy stand-ins (front-end tests)

class LazyDataArray(DataArray):
    """A DataArray whose backing store is read block by block (like xarray's lazily indexed
    netCDF variables): every ``.values`` / ``np.asarray`` materialisation is recorded in
    ``reads`` (bytes) so that a test can bound the peak host memory of the front end."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.reads = []

    @property
    def values(self):
        self.reads.append(self.data.nbytes)
        return self.data

    def _new(self, data, dims=None):
        out = LazyDataArray(data, self.dims if dims is None else dims, name=self.name, attrs=self.attrs)
        out.reads = self.reads                       # children report to the root array
        return out

    def transpose(self, *dims):
        out = super().transpose(*dims)
        lazy = LazyDataArray(out.data, out.dims, name=out.name, attrs=out.attrs)
        for k, v in out.coords.items():
            lazy.coords[k] = v
        lazy.reads = self.reads
        return lazy


class FakeDaskArray:
    """Minimal stand-in for ``dask.array.Array``: numpy data + a chunk structure, evaluated only by
    ``compute()``.  ``map_blocks`` below records one task per block without running it."""

    def __init__(self, data=None, chunks=None, tasks=None, shape=None, dtype=None):
        self._data, self.chunks, self._tasks = data, tuple(tuple(c) for c in chunks), tasks
        self.shape = tuple(sum(c) for c in self.chunks) if shape is None else shape
        self.dtype = np.dtype(dtype) if dtype is not None else data.dtype
        self.ndim = len(self.chunks)

    @property
    def nbytes(self): return int(np.prod(self.shape)) * self.dtype.itemsize

    def rechunk(self, spec):
        chunks = list(self.chunks)
        for ax, c in spec.items():
            chunks[ax] = (self.shape[ax],) if c == -1 else tuple(c)
        return FakeDaskArray(self._data, chunks, self._tasks, dtype=self.dtype)

    def blocks(self):
        import itertools
        starts = [np.concatenate(([0], np.cumsum(c)))[:-1] for c in self.chunks]
        for idx in itertools.product(*[range(len(c)) for c in self.chunks]):
            yield idx, tuple(slice(int(starts[a][i]), int(starts[a][i] + self.chunks[a][i])) for a, i in enumerate(idx))

    def compute(self):
        if self._tasks is None:
            return self._data
        out = np.empty(self.shape, self.dtype)
        for sl, thunk in self._tasks:
            out[sl] = thunk()
        return out

    def __array__(self, dtype=None, copy=None):
        raise AssertionError("a dask-backed field must never be materialised by the front end")


FakeDaskArray.__module__ = "dask.array.core"


def fake_dask_map_blocks(func, x, dtype=None, chunks=None, drop_axis=(), new_axis=(), meta=None):
    """``dask.array.map_blocks`` for one input whose dropped axes are single-chunk: builds the
    output block structure from ``chunks`` and defers ``func`` to ``compute()``."""
    assert all(len(x.chunks[a]) == 1 for a in drop_axis), "dropped axes must be one chunk"
    keep = [a for a in range(x.ndim) if a not in set(drop_axis)]
    out = FakeDaskArray(None, chunks, tasks=[], dtype=dtype)
    calls = []
    out_starts = [np.concatenate(([0], np.cumsum(c)))[:-1] for c in out.chunks]
    for idx, sl in x.blocks():
        osl, k = [], 0
        for a in range(out.ndim):
            if a in set(new_axis):
                osl.append(slice(0, out.shape[a]))
            else:
                i = idx[keep[k]]; k += 1
                osl.append(slice(int(out_starts[a][i]), int(out_starts[a][i] + out.chunks[a][i])))
        def thunk(sl=sl):
            calls.append(sl)
            return func(np.asarray(x.compute()[sl]) if x._tasks is not None else x._data[sl])
        out._tasks.append((tuple(osl), thunk))
    out.calls = calls
    return out


def install_fake_dask():
    """Register the stand-in as ``dask`` / ``dask.array`` (front-end tests only)."""
    dask = types.ModuleType("dask")
    da = types.ModuleType("dask.array")
    da.Array = FakeDaskArray
    da.map_blocks = fake_dask_map_blocks
    da.moveaxis = lambda a, s, d: FakeDaskArray(np.moveaxis(a.compute(), s, d), [(n,) for n in np.moveaxis(np.empty(a.shape, bool), s, d).shape])
    dask.array = da
    sys.modules.update({"dask": dask, "dask.array": da})
    return da


# -------------------------------------------------------------------------- installer

def install():
    """Register the stand-ins as ``xarray``, ``dask``, ``dask.array`` and ``sparse``."""
    xr = types.ModuleType("xarray")
    xr.DataArray, xr.Dataset = DataArray, Dataset
    xr.open_dataset = xr.open_mfdataset = xr.merge = xr.concat = None
    dask = types.ModuleType("dask")
    da = types.ModuleType("dask.array")
    ma = types.ModuleType("dask.array.ma")
    ma.set_fill_value, ma.fix_invalid, ma.filled = np.ma.set_fill_value, np.ma.fix_invalid, np.ma.filled
    da.ma = ma
    da.Array = _DaskArray
    da.reshape, da.where, da.broadcast_to, da.stack = np.reshape, np.where, np.broadcast_to, np.stack
    da.tensordot, da.from_delayed = _tensordot, _from_delayed
    dask.array, dask.delayed = da, _delayed
    sp = types.ModuleType("sparse")
    sp.COO = COO
    sys.modules.update({"xarray": xr, "dask": dask, "dask.array": da, "dask.array.ma": ma, "sparse": sp})
    return xr
