"""Property tests on the GPU (hypothesis): random link sets, NaN/inf patterns, masks, dtypes and
strides against the oracle -- all kernel families, every lane configuration."""
import ctypes
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from helpers import assert_parity

pytestmark = pytest.mark.gpu


@st.composite
def cases(draw):
    rng = np.random.default_rng(draw(st.integers(0, 2**31 - 1)))
    n_src = draw(st.integers(1, 5000))
    n_dst = draw(st.integers(1, 900))
    max_links = draw(st.sampled_from([0, 1, 2, 4, 7, 9, 16, 20, 40, 70, 130, 260, 600]))
    local = draw(st.booleans())
    B = draw(st.integers(1, 67))
    xdt = draw(st.sampled_from([np.float32, np.float64]))
    ydt = draw(st.sampled_from([np.float32, np.float64]))
    counts = rng.integers(0, max_links + 1, size=n_dst)
    if draw(st.booleans()):         # mostly short rows with a few long ones: split plans (packed + row list)
        counts = rng.integers(0, 9, size=n_dst)
        nlong = min(n_dst, draw(st.integers(0, 12)))
        counts[rng.choice(n_dst, size=nlong, replace=False)] = rng.integers(17, max(18, max_links + 1), size=nlong)
    dst = np.repeat(np.arange(n_dst), counts)
    if local:
        spread = draw(st.sampled_from([8, 64, 400]))
        src = np.clip((dst * n_src) // n_dst + rng.integers(-spread, spread + 1, size=dst.size), 0, n_src - 1)
    else:
        src = rng.integers(0, n_src, size=dst.size)
    w = rng.random(dst.size) - draw(st.sampled_from([0.0, 0.0, 0.3]))
    order = rng.permutation(dst.size) if draw(st.booleans()) else np.lexsort((src, dst))
    x = (draw(st.sampled_from([1.0, 280.0, 1e-30, 1e18])) * (1 + rng.standard_normal((B, n_src)))).astype(xdt)
    nanfrac = draw(st.sampled_from([0.0, 0.0, 0.01, 0.3, 1.0]))
    x[rng.random(x.shape) < nanfrac] = np.nan
    if draw(st.booleans()) and x.size:
        x.flat[rng.integers(0, x.size)] = np.inf
        x.flat[rng.integers(0, x.size)] = -np.inf
        x.flat[rng.integers(0, x.size)] = 0.0
    masked = draw(st.booleans())
    amin = draw(st.sampled_from([0.0, 0.5, 1.0]))
    return dict(n_src=n_src, n_dst=n_dst, src=(src[order] + 1).astype(np.int32), dst=(dst[order] + 1).astype(np.int32),
                w=w[order].reshape(-1, 1), x=x, ydt=ydt, masked=masked, amin=amin,
                imask=(rng.random(n_dst) > 0.2).astype(np.int32), frac=rng.random(n_dst),
                pad=draw(st.sampled_from([0, 4, 1])))


@settings(max_examples=int(os.environ.get("SMM_HYP_EXAMPLES", "60")), deadline=None,
          suppress_health_check=list(HealthCheck), derandomize=os.environ.get("SMM_HYP_RANDOM") is None)
@given(c=cases())
def test_random_operator_matches_oracle(smm_lib, oracle, cuda, c):
    import torch
    from smmregrid_b200 import _lib
    mat = oracle.compute_weights_matrix_c(c["src"], c["dst"], c["w"], c["n_src"], c["n_dst"])
    y_ref = oracle.apply_weights_c(c["x"], mat, c["imask"], c["frac"], c["amin"], c["masked"])
    h = ctypes.c_void_p()
    _lib.check(smm_lib.smm_create(c["n_src"], c["n_dst"], c["src"].size, c["src"].ctypes.data, c["dst"].ctypes.data,
                                  c["w"].ctypes.data, 1, 1, 0, None, ctypes.byref(h)))
    try:
        _lib.check(smm_lib.smm_set_dst_mask(h, 0, c["imask"].ctypes.data, c["frac"].ctypes.data))
        B, n_src, n_dst = c["x"].shape[0], c["n_src"], c["n_dst"]
        ldx = n_src + c["pad"]
        xd = torch.zeros((B, ldx), dtype=torch.from_numpy(c["x"]).dtype, device="cuda")
        xd[:, :n_src] = torch.from_numpy(c["x"]).cuda()
        tol = 1e-12 if c["ydt"] == np.float64 else 1e-6
        with np.errstate(over="ignore"):
            ref = y_ref.astype(c["ydt"])
        for kernel in (0, 2, 3):           # automatic, direct gathers, two-pass compact path for gather-family levels
            y = torch.full((B, n_dst), 3.0, dtype=torch.float64 if c["ydt"] == np.float64 else torch.float32, device="cuda")
            _lib.check(smm_lib.smm_apply(h, 0, xd.data_ptr(), 0 if c["x"].dtype == np.float32 else 1, B, ldx,
                                         y.data_ptr(), 1 if c["ydt"] == np.float64 else 0, n_dst,
                                         int(c["masked"]), c["amin"], _lib.apply_opts(kernel), None))
            torch.cuda.synchronize()
            got = y.cpu().numpy()
            assert np.array_equal(np.isnan(got), np.isnan(ref))
            ok = ~np.isnan(ref)
            # backward-error bound: rows with negative weights may cancel, so scale by sum |w||x|
            xf = np.where(np.isfinite(c["x"]), c["x"], 1e20).astype(np.float64)
            scale = np.zeros((B, n_dst))
            np.add.at(scale, (slice(None), mat.dst), np.abs(xf[:, mat.src]) * np.abs(mat.w))
            err = np.abs(got.astype(np.float64) - ref.astype(np.float64))
            fin = ok & np.isfinite(ref)
            assert np.all(err[fin] <= tol * np.maximum(scale[fin], 1e-300)), (kernel, err[fin].max())
            assert np.array_equal(np.isinf(got), np.isinf(ref))
            if (mat.w < 0).any() and c["ydt"] == np.float64:
                # weights of both signs: summed in the reference's order on every path -> bit-identical
                assert np.array_equal(got[ok], ref[ok]), (kernel, int((got[ok] != ref[ok]).sum()))
    finally:
        smm_lib.smm_destroy(h)
