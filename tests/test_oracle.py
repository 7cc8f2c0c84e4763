"""The CPU oracle against the golden fixtures and against itself (no GPU).

tests/golden/*.npz were produced by executing the reference's own weights.py /
Regridder.apply_weights under numpy-backed stand-ins for dask/sparse/xarray
(tests/golden/make_golden.py).  Both oracle restatements (numpy and the C port) must
reproduce them: COO contents and masks bit-exactly, applied fields bit-exactly for the
loop-order faithful C port and to rounding for the scipy-ordered numpy one.
"""
import glob
import os

import numpy as np
import pytest

from helpers import assert_parity, random_links

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES_2D = ["bil_f32", "con_masked_f32", "con_masked_f64", "unsorted_dups_f64"]


def load(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def test_fixtures_present():
    names = {os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))}
    assert set(CASES_2D + ["ocean3d_f32"]) <= names


@pytest.mark.parametrize("name", CASES_2D)
@pytest.mark.parametrize("builder", ["np", "c"])
def test_weights_matrix_golden(oracle, name, builder):
    g = load(name)
    n_src, n_dst = int(np.prod(g["in_src_grid_dims"])), int(np.prod(g["in_dst_grid_dims"]))
    build = oracle.compute_weights_matrix_np if builder == "np" else oracle.compute_weights_matrix_c
    mat = build(g["in_src_address"], g["in_dst_address"], g["in_remap_matrix"], n_src, n_dst)
    assert np.array_equal(mat.src, g["coo_src"]) and np.array_equal(mat.dst, g["coo_dst"])
    assert np.array_equal(mat.w, g["coo_w"])
    for fn in (oracle.mask_tensordot_np, oracle.mask_tensordot_c):
        imask, _ = fn(g["in_src_grid_imask"].ravel(), mat)
        assert np.array_equal(imask, g["dst_grid_imask"].ravel())
    assert oracle.check_mask_np(g["dst_grid_imask"]) == bool(g["masked"])


@pytest.mark.parametrize("name", CASES_2D)
def test_apply_golden(oracle, name):
    g = load(name)
    n_src, n_dst = int(np.prod(g["in_src_grid_dims"])), int(np.prod(g["in_dst_grid_dims"]))
    mat = oracle.compute_weights_matrix_c(g["in_src_address"], g["in_dst_address"], g["in_remap_matrix"], n_src, n_dst)
    x = g["in_x"]
    nh = len(g["in_src_grid_dims"])
    xf = x.reshape(x.shape[:x.ndim - nh] + (n_src,))
    imask, frac = g["dst_grid_imask"].ravel(), g["in_dst_grid_frac"].ravel()
    for i, am in enumerate(g["area_mins"]):
        for key, masked in ((f"y_{i}", bool(g["masked"])), (f"y_unmasked_{i}", False)):
            ref = g[key].reshape(xf.shape[:-1] + (n_dst,))
            assert ref.dtype == np.float64                      # result_type(data, float64 weights)
            y_c = oracle.apply_weights_c(xf, mat, imask, frac, float(am), masked)
            assert np.array_equal(np.isnan(y_c), np.isnan(ref))
            assert np.array_equal(y_c[~np.isnan(ref)], ref[~np.isnan(ref)]), "C port must be bit-exact"
            y_np = oracle.apply_weights_np(xf, mat, imask, frac, float(am), masked)
            assert_parity(y_np, ref, 1e-12, name)


def test_ocean3d_golden(oracle):
    g = load("ocean3d_f32")
    n_src, n_dst = int(np.prod(g["in_src_grid_dims"])), int(np.prod(g["in_dst_grid_dims"]))
    L = g["in_link_length"].size
    mats = oracle.compute_weights_matrix3d_np(g["in_src_address"], g["in_dst_address"], g["in_remap_matrix"],
                                              g["in_link_length"], n_src, n_dst, builder=oracle.compute_weights_matrix_c)
    assert [m.nnz for m in mats] == list(g["coo_nnz"])
    imask = np.stack([oracle.mask_tensordot_c(g["in_src_grid_imask"][l], mats[l])[0] for l in range(L)])
    assert np.array_equal(imask, g["dst_grid_imask"])
    masked = oracle.check_mask_np(imask)
    assert np.array_equal(masked, g["masked"])
    x = g["in_x"]
    xf = x.reshape(x.shape[:2] + (n_src,))
    y = oracle.regrid3d_np(xf, 1, g["levels"], g["levels"], mats, imask, g["in_dst_grid_frac"], masked,
                           float(g["area_min"]))
    ref = g["y"].reshape(y.shape)
    assert np.array_equal(np.isnan(y), np.isnan(ref))
    assert np.array_equal(y[~np.isnan(ref)], ref[~np.isnan(ref)])


def test_c_vs_numpy_random(oracle):
    rng = np.random.default_rng(0)
    for trial in range(6):
        n_src, n_dst = int(rng.integers(5, 400)), int(rng.integers(1, 200))
        src, dst, w = random_links(rng, n_src, n_dst, int(rng.integers(1, 12)), dup_frac=0.1, sort=False,
                                   negative=trial % 2 == 1)
        a = oracle.compute_weights_matrix_np(src, dst, w, n_src, n_dst)
        b = oracle.compute_weights_matrix_c(src, dst, w, n_src, n_dst)
        assert np.array_equal(a.src, b.src) and np.array_equal(a.dst, b.dst) and np.array_equal(a.w, b.w)
        x = rng.standard_normal((7, n_src)).astype(np.float32 if trial % 3 else np.float64) * 100
        x[rng.random(x.shape) < 0.1] = np.nan
        imask = (rng.random(n_dst) > 0.2).astype(np.int32)
        frac = rng.random(n_dst)
        ya = oracle.apply_weights_np(x, a, imask, frac, 0.4, True)
        yb = oracle.apply_weights_c(x, b, imask, frac, 0.4, True, nthreads=3)
        assert np.array_equal(np.isnan(ya), np.isnan(yb))
        ok = ~np.isnan(ya)
        scale = np.abs(x[np.isfinite(x)]).max() * np.abs(w).sum() / n_dst + 1e-300
        assert np.all(np.abs(ya[ok] - yb[ok]) <= 1e-12 * np.maximum(np.abs(yb[ok]), scale))


def test_reference_quirks(oracle):
    """Known answers derivable without CDO."""
    n = 8
    src = np.arange(1, n + 1, dtype=np.int32)
    mat = oracle.compute_weights_matrix_c(src, src, np.ones((n, 1)), n, n)
    x = np.arange(n, dtype=np.float32).reshape(1, n) + 1
    assert np.array_equal(oracle.apply_weights_c(x, mat, None, None, 0.0, False), x.astype(np.float64))
    # all-NaN row -> all-NaN
    assert np.isnan(oracle.apply_weights_c(np.full((1, n), np.nan, np.float32), mat, None, None, 0.0, False)).all()
    # 0.05-weight NaN leaks 5e18 instead of NaN; float32 fill promotes to 1.00000002004e20
    w = np.ones((n, 1)); w[0] = 0.05
    mat = oracle.compute_weights_matrix_c(src, src, w, n, n)
    xn = x.copy(); xn[0, 0] = np.nan
    y = oracle.apply_weights_c(xn, mat, None, None, 0.0, False)
    assert y[0, 0] == 0.05 * np.float64(np.float32(1e20)) and np.isfinite(y[0, 0])
    y64 = oracle.apply_weights_c(xn.astype(np.float64), mat, None, None, 0.0, False)
    assert y64[0, 0] == 0.05 * 1e20
    # +-inf are filled too; legit values above 1e19 become NaN
    xi = x.astype(np.float64).copy(); xi[0, 1] = np.inf; xi[0, 2] = -np.inf; xi[0, 3] = 2e19
    y = oracle.apply_weights_c(xi, mat, None, None, 0.0, False)
    assert np.isnan(y[0, 1]) and np.isnan(y[0, 2]) and np.isnan(y[0, 3])
    # remap_area_min monotonicity of the NaN count (remapareamin_test.py:15-29)
    frac = np.linspace(0, 1, n)
    counts = [np.isnan(oracle.apply_weights_c(x, mat, None, frac, am, False)).sum() for am in (0.0, 0.5, 0.9)]
    assert counts == sorted(counts) and counts[0] == 0 and counts[-1] > counts[1] > 0


def test_level_selection(oracle):
    lv = [0.5, 10.0, 100.0, 1000.0]
    assert oracle.select_level_np(lv, 10.0004) == 1
    assert oracle.select_level_np(lv, 1000.0) == 3
    with pytest.raises(ValueError):
        oracle.select_level_np(lv, 10.01)
    from smmregrid_b200.regrid import select_level
    assert select_level(lv, 99.9995) == 2
    with pytest.raises(ValueError, match="not found in mask_dim"):
        select_level(lv, 55.0)


def test_out_of_range_addresses(oracle):
    for fn in (oracle.compute_weights_matrix_np, oracle.compute_weights_matrix_c):
        with pytest.raises(ValueError):
            fn(np.array([0]), np.array([1]), np.ones((1, 1)), 4, 4)
        with pytest.raises(ValueError):
            fn(np.array([1]), np.array([5]), np.ones((1, 1)), 4, 4)


def test_nan_variation_oracle_matches_reference_fixture(oracle):
    """detect_nan_variation_dims: fixture produced by the reference's own util.py function."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "nan_variation.npz"))
    names = [k for k in g.files if not k.endswith("_dims")]
    assert len(names) == 10
    for k in names:
        want = g[k + "_dims"].tolist()
        assert oracle.detect_nan_variation_dims_np(g[k], 0, [1, 2]) == want, k
        assert oracle.detect_nan_variation_dims_np(g[k][0], None, [0, 1]) == [d - 1 for d in want], k
