"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np


def random_links(rng, n_src, n_dst, nnz_per_row, dup_frac=0.0, sort=True, negative=False):
    """Random CDO-style link arrays (1-based).  Rows get 0..2*nnz_per_row links."""
    counts = rng.integers(0, 2 * nnz_per_row + 1, size=n_dst)
    dst = np.repeat(np.arange(n_dst), counts)
    src = rng.integers(0, n_src, size=dst.size)
    w = rng.random(dst.size)
    if negative:
        w -= 0.3
    if dup_frac > 0 and dst.size:
        k = int(dup_frac * dst.size)
        pick = rng.integers(0, dst.size, size=k)
        dst = np.concatenate([dst, dst[pick]])
        src = np.concatenate([src, src[pick]])
        w = np.concatenate([w, rng.random(k)])
    if sort:
        o = np.lexsort((src, dst))
    else:
        o = rng.permutation(dst.size)
    return (src[o] + 1).astype(np.int32), (dst[o] + 1).astype(np.int32), w[o].reshape(-1, 1)


def assert_parity(y, y_ref, rtol, what=""):
    """Bit-exact NaN pattern; values within rtol relative (floor on tiny magnitudes)."""
    y = np.asarray(y, dtype=np.float64)
    y_ref = np.asarray(y_ref, dtype=np.float64)
    assert y.shape == y_ref.shape, (what, y.shape, y_ref.shape)
    nan, nan_ref = np.isnan(y), np.isnan(y_ref)
    assert np.array_equal(nan, nan_ref), \
        f"{what}: NaN pattern differs at {int((nan != nan_ref).sum())} of {y.size} cells"
    ok = ~nan
    if ok.any():
        scale = np.maximum(np.abs(y_ref[ok]), 1e-300)
        err = np.abs(y[ok] - y_ref[ok]) / scale
        assert err.max() <= rtol, f"{what}: max rel err {err.max():.3e} > {rtol:.1e}"
