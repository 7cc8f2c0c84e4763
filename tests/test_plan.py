"""Host logic of the operator construction (no GPU): CSR build + staged tile plan.

The plan arrays are pulled through the host-only introspection ABI and the staged kernel's
data flow is EMULATED here in numpy (segments -> staged footprint -> per-lane links ->
lane-group reduction); the emulation must reproduce the oracle.  This validates exactly what
smm_create uploads to the device.
"""
import ctypes

import numpy as np
import pytest

from helpers import assert_parity, random_links


class HostPlan:
    def __init__(self, lib, src, dst, rm, n_src, n_dst, summation=0, cache_dir=None):
        from smmregrid_b200 import _lib
        self.lib = lib
        src = np.ascontiguousarray(src, np.int32)
        dst = np.ascontiguousarray(dst, np.int32)
        rm = np.ascontiguousarray(rm, np.float64)
        h = ctypes.c_void_p()
        _lib.check(lib.smm_host_plan_build(n_src, n_dst, src.size, src.ctypes.data, dst.ctypes.data,
                                           rm.ctypes.data, rm.shape[1], 1, _lib.create_opts(summation, cache_dir), ctypes.byref(h)))
        inf = _lib.SmmInfo()
        ns = ctypes.c_int64()
        _lib.check(lib.smm_host_plan_info(h, ctypes.byref(inf), ctypes.byref(ns)))
        self.info = inf.asdict()
        nnz, nt, kpl = inf.nnz, inf.n_tiles, inf.links_per_lane
        self.nct = nct = inf.consumer_threads or 256
        self.rowptr = np.empty(n_dst + 1, np.int32)
        self.col = np.empty(nnz, np.int32)
        self.val = np.empty(nnz, np.float64)
        self.tiles = np.empty((nt, 8), np.int32)
        self.segs = np.empty((ns.value, 4), np.uint32)
        self.wplan = np.empty((nt, kpl, nct), np.float64)
        self.iplan = np.empty((nt, kpl, nct), np.uint16)
        _lib.check(lib.smm_host_plan_copy(h, self.rowptr.ctypes.data, self.col.ctypes.data, self.val.ctypes.data,
                                          self.tiles.ctypes.data, self.segs.ctypes.data,
                                          self.wplan.ctypes.data, self.iplan.ctypes.data))
        self.rowmap = None
        if inf.rows_reordered and not inf.packed_rows:      # packed plans carry the rows in rowslot
            self.rowmap = np.empty(n_dst, np.int32)
            _lib.check(lib.smm_host_plan_rowmap(h, self.rowmap.ctypes.data))
        self.gather_rows = np.empty(inf.gather_rows, np.int32)
        if inf.gather_rows:
            _lib.check(lib.smm_host_plan_gather_rows(h, self.gather_rows.ctypes.data))
        self.rowslot = None
        if inf.packed_rows:
            self.rowslot = np.empty((nt, 4, nct), np.int32)
            _lib.check(lib.smm_host_plan_rowslot(h, self.rowslot.ctypes.data))
        lib.smm_host_plan_free(h)

    def emulate_reference_order(self, x):
        """What the ORD staged kernels compute for filled input x [B, n_src]: one chain per row
        in (lane, slot) order -- lane l of a row's group holds links l*kpl .. of its ascending-source
        list -- with separate multiply and add (numpy never contracts to FMA)."""
        assert self.info["summation"] == 2
        B = x.shape[0]
        lpr, kpl = self.info["lanes_per_row"], self.info["links_per_lane"]
        y = np.zeros((B, self.info["n_dst"]))
        for t, (row0, nrows, seg0, nseg, elems, *_r) in enumerate(self.tiles):
            stage = np.zeros((B, max(elems, 1)))
            for s, d, ln, _p in self.segs[seg0:seg0 + nseg]:
                stage[:, d:d + ln] = x[:, s:s + ln]
            prod = stage[:, self.iplan[t]] * self.wplan[t][None]                 # [B, kpl, nct], each rounded once
            if self.rowslot is not None:
                rs = self.rowslot[t]
                chain = np.zeros((B, self.nct))
                last = np.zeros((B, 4, self.nct))
                for u in range(4):
                    chain = np.where(rs[u] == -2, chain, 0.0)                    # a continuation keeps the chain
                    for j in range(4):
                        chain = chain + prod[:, 4 * u + j]
                    last[:, u] = chain
                val = last.copy()
                for u in (2, 1, 0):                                              # a row's value = chain after its last sub-row
                    val[:, u] = np.where(rs[u + 1] == -2, val[:, u + 1], last[:, u])
                own = rs >= 0
                assert own.sum() == nrows
                y[:, rs[own]] = val[:, own]
                continue
            rows = np.zeros((B, self.nct // lpr))
            p4 = prod.reshape(B, kpl, self.nct // lpr, lpr)
            for l in range(lpr):
                for k in range(kpl):
                    rows = rows + p4[:, k, :, l]
            if self.rowmap is None:
                y[:, row0:row0 + nrows] = rows[:, :nrows]
            else:
                y[:, self.rowmap[row0:row0 + nrows]] = rows[:, :nrows]
        self._gather_rows_into(y, x)                   # (ascending source, separate multiply and add: reference order)
        return y

    def _gather_rows_into(self, y, x):
        """split plans: the long rows are served by the gather kernel straight from the CSR"""
        for r in self.gather_rows:
            a, b = self.rowptr[r], self.rowptr[r + 1]
            acc = np.zeros(x.shape[0])
            for j in range(a, b):
                acc = acc + x[:, self.col[j]] * self.val[j]
            y[:, r] = acc

    def emulate(self, x):
        """What staged_kernel computes for filled input x [B, n_src] (float64 math)."""
        B = x.shape[0]
        lpr = self.info["lanes_per_row"]
        y = np.zeros((B, self.info["n_dst"]))
        for t, (row0, nrows, seg0, nseg, elems, *_r) in enumerate(self.tiles):
            stage = np.zeros((B, max(elems, 1)))
            for s, d, ln, _p in self.segs[seg0:seg0 + nseg]:
                stage[:, d:d + ln] = x[:, s:s + ln]
            if self.rowslot is not None:
                # packed rows: thread = 4 sub-rows of 4 links; a row = its first sub-row + continuations
                sub = (stage[:, self.iplan[t]] * self.wplan[t][None]).reshape(B, 4, 4, self.nct).sum(axis=2)
                rs = self.rowslot[t]
                for u in (2, 1, 0):
                    sub[:, u] += np.where(rs[u + 1] == -2, sub[:, u + 1], 0.0)
                own = rs >= 0
                assert own.sum() == nrows
                y[:, rs[own]] = sub[:, own]
                continue
            lane = (stage[:, self.iplan[t]] * self.wplan[t][None]).sum(axis=1)      # [B, nct]
            rows = lane.reshape(B, self.nct // lpr, lpr).sum(axis=2)
            if self.rowmap is None:
                y[:, row0:row0 + nrows] = rows[:, :nrows]
            else:                                      # row0 = first tile slot in the re-ordered sequence
                y[:, self.rowmap[row0:row0 + nrows]] = rows[:, :nrows]
        self._gather_rows_into(y, x)
        return y


def _check_plan_invariants(p, n_src):
    t, s = p.tiles, p.segs
    assert (t[:, 1] > 0).all() and t[:, 1].sum() == p.info["n_dst"]
    assert (np.diff(t[:, 0]) == t[:-1, 1]).all()
    for row0, nrows, seg0, nseg, elems, *_ in t:
        sg = s[seg0:seg0 + nseg]
        if nseg:
            assert (sg[:, 0] % 8 == 0).all() and (sg[:, 1] % 8 == 0).all()      # 16-byte aligned TMA
            assert (sg[:, 0] + sg[:, 2] <= n_src).all()
            assert (np.diff(sg[:, 0].astype(np.int64)) > 0).all()
            assert sg[:, 2].sum() == elems
            assert (sg[1:, 1] == np.cumsum(sg[:-1, 2])).all()
    assert p.iplan.max(initial=0) < max(1, p.info["max_tile_elems"])


@pytest.mark.parametrize("nnz_per_row", [1, 2, 5, 10, 20, 50, 100, 240])
def test_plan_emulation_matches_oracle(smm_lib, oracle, nnz_per_row):
    rng = np.random.default_rng(nnz_per_row)
    n_src, n_dst, B = 3000, 611, 5
    counts = rng.integers(0, nnz_per_row + 1, size=n_dst)
    counts[7] = nnz_per_row
    dst = np.repeat(np.arange(n_dst), counts)
    src = np.clip((dst * n_src) // n_dst + rng.integers(-200, 200, size=dst.size), 0, n_src - 1)
    w = rng.random(dst.size)
    o = rng.permutation(dst.size)                        # unsorted input, duplicates included
    src, dst, w = src[o] + 1, dst[o] + 1, w[o].reshape(-1, 1)
    p = HostPlan(smm_lib, src, dst, w, n_src, n_dst)
    assert p.info["kernel_name"] == "staged", p.info
    assert p.info["lanes_per_row"] * p.info["links_per_lane"] >= p.info["max_row_nnz"]
    _check_plan_invariants(p, n_src)
    mat = oracle.compute_weights_matrix_c(src, dst, w, n_src, n_dst)
    # CSR rows hold the oracle's links in ascending-src order with duplicates summed
    o2 = np.lexsort((mat.src, mat.dst))
    assert np.array_equal(p.col, mat.src[o2])
    assert np.array_equal(np.repeat(np.arange(n_dst), np.diff(p.rowptr)), mat.dst[o2])
    assert np.allclose(p.val, mat.w[o2], rtol=1e-15, atol=0)
    x = rng.standard_normal((B, n_src))
    y_ref = oracle.apply_weights_c(x, mat, None, None, 0.0, False)
    assert_parity(np.where(np.abs(y_ref) > 0, p.emulate(x), 0.0), y_ref, 1e-11)


def test_plan_named_configs(smm_lib, oracle):
    from smmregrid_b200 import synth
    for cfg, scale, lanes in (("C1", 1, (1, 16)), ("C2", 4, (2, 14)), ("C4", 5, (8, 14))):
        w = synth.config_weights(cfg, scale)
        n_src, n_dst = w.sizes["src_grid_size"], w.sizes["dst_grid_size"]
        p = HostPlan(smm_lib, w["src_address"], w["dst_address"], w["remap_matrix"], n_src, n_dst)
        assert p.info["kernel_name"] == "staged"
        assert (p.info["lanes_per_row"], p.info["links_per_lane"]) == lanes
        assert p.info["packed_rows"] == (cfg == "C1")        # rows of <= 16 links pack 4 to a thread
        assert p.info["touched_src"] == n_src
        if cfg != "C1":                                      # down-sampling: little over-read of the slab
            assert p.info["sum_tile_elems"] < 1.3 * n_src
        _check_plan_invariants(p, n_src)
        x = np.random.default_rng(1).standard_normal((2, n_src)) + 10
        mat = oracle.compute_weights_matrix_c(w["src_address"], w["dst_address"], w["remap_matrix"], n_src, n_dst)
        assert_parity(p.emulate(x), oracle.apply_weights_c(x, mat, None, None, 0.0, False), 1e-12, cfg)


def test_scattered_sources_use_gather(smm_lib):
    from smmregrid_b200 import synth
    w = synth.config_weights("C5dis", 8)
    p = HostPlan(smm_lib, w["src_address"], w["dst_address"], w["remap_matrix"],
                 w.sizes["src_grid_size"], w.sizes["dst_grid_size"])
    assert p.info["kernel_name"] == "gather"
    assert p.info["nnz"] == 4 * w.sizes["dst_grid_size"]


def test_long_rows_use_gather(smm_lib):
    n_src, n_dst = 2000, 3
    dst = np.repeat(np.arange(n_dst), 600)
    src = np.tile(np.arange(600), n_dst)
    p = HostPlan(smm_lib, src + 1, dst + 1, np.ones((dst.size, 1)), n_src, n_dst)
    assert p.info["kernel_name"] == "gather" and p.info["max_row_nnz"] == 600


def test_empty_and_ragged(smm_lib, oracle):
    # no links at all
    p = HostPlan(smm_lib, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 1)), 10, 4)
    assert p.info["nnz"] == 0 and np.array_equal(p.rowptr, np.zeros(5))
    # links reaching the last source cell: the aligned segment is clipped at n_src (13 = odd tail)
    p = HostPlan(smm_lib, np.array([13]), np.array([4]), np.array([[2.0]]), 13, 4)
    assert p.info["nnz"] == 1 and p.col[0] == 12 and p.info["kernel_name"] == "gather"   # too short to stage
    n_src = 45
    src = np.arange(20, 46)
    p = HostPlan(smm_lib, src, np.full(src.size, 4), np.ones((src.size, 1)), n_src, 4)
    assert p.info["kernel_name"] == "staged"
    _check_plan_invariants(p, n_src)
    assert p.segs[-1, 0] == 16 and p.segs[-1, 2] == n_src - 16
    x = np.arange(float(n_src)).reshape(1, n_src)
    assert p.emulate(x)[0, 3] == x[0, 19:].sum() and (p.emulate(x)[0, :3] == 0).all()


def test_address_range_errors(smm_lib):
    from smmregrid_b200 import _lib
    h = ctypes.c_void_p()
    for src, dst in (([0], [1]), ([1], [0]), ([11], [1]), ([1], [5])):
        s, d = np.array(src, np.int32), np.array(dst, np.int32)
        w = np.ones((1, 1))
        rc = smm_lib.smm_host_plan_build(10, 4, 1, s.ctypes.data, d.ctypes.data, w.ctypes.data, 1, 1, None, ctypes.byref(h))
        assert rc == _lib.SMM_ERR_RANGE
        assert b"outside the grids" in smm_lib.smm_last_error()


@pytest.mark.parametrize("packed", ["0", None])
def test_healpix_nested_source_gets_reordered_plan(smm_lib, oracle, monkeypatch, packed):
    """A HEALPix-nested source is scattered along destination rows but compact in 2-D: the
    natural lane-per-row plan is rejected, the plan on rows re-ordered by mean source address is
    staged.  (The default packed layout tiles 4x the rows and may get by in natural order.)"""
    from smmregrid_b200 import synth
    if packed is not None:
        monkeypatch.setenv("SMM_PACKED", packed)
    for k in (1, 4):
        w = synth.healpix_weights(64, 180, 90, k)
        n_src, n_dst = w.sizes["src_grid_size"], w.sizes["dst_grid_size"]
        # nearest-neighbour sanity: every pixel index valid, poles map to the polar faces
        assert w["src_address"].min() >= 1 and w["src_address"].max() <= n_src
        p = HostPlan(smm_lib, w["src_address"], w["dst_address"], w["remap_matrix"], n_src, n_dst)
        assert p.info["kernel_name"] == "staged", p.info
        assert p.info["packed_rows"] == (packed is None)
        assert p.info["rows_reordered"] == 1 or packed is None, p.info
        rows = p.rowslot[p.rowslot >= 0] if p.info["packed_rows"] else p.rowmap
        assert sorted(rows.tolist()) == list(range(n_dst))
        t, sg = p.tiles, p.segs
        assert t[:, 1].sum() == n_dst and (sg[:, 0] % 8 == 0).all()
        x = np.random.default_rng(k).standard_normal((3, n_src)) + 5
        mat = oracle.compute_weights_matrix_c(w["src_address"], w["dst_address"], w["remap_matrix"], n_src, n_dst)
        assert_parity(p.emulate(x), oracle.apply_weights_c(x, mat, None, None, 0.0, False), 1e-12, f"hp{k}")


def test_ang2pix_nest_known_values():
    from smmregrid_b200.synth import ang2pix_nest
    # nside = 1: the 12 base pixels; centres at lat +-41.81 (caps) and 0 (belt)
    lat = np.array([41.81, 41.81, 41.81, 41.81, 0, 0, 0, 0, -41.81, -41.81, -41.81, -41.81])
    lon = np.array([45, 135, 225, 315, 0, 90, 180, 270, 45, 135, 225, 315.0])
    assert ang2pix_nest(1, lat, lon).tolist() == list(range(12))
    # nside = 2: every pixel hit exactly by a dense sampling, indices within range
    la, lo = np.meshgrid(np.linspace(-89.9, 89.9, 400), np.linspace(0.1, 359.9, 800), indexing="ij")
    pix = ang2pix_nest(2, la, lo)
    assert pix.min() == 0 and pix.max() == 47 and np.unique(pix).size == 48


@pytest.mark.parametrize("reorder", [False, True])
def test_packed_rows_plan(smm_lib, oracle, reorder):
    """Short-row operators (max <= 16 links, mostly <= 4): a thread owns up to four rows."""
    rng = np.random.default_rng(11)
    n_src, n_dst, B = 6000, 5003, 3
    counts = rng.choice([0, 1, 2, 3, 4, 4, 4, 6, 9, 13, 16], size=n_dst)
    dst = np.repeat(np.arange(n_dst), counts)
    centre = (dst * n_src) // n_dst
    if reorder:                                           # rows permuted: natural tiles would be scattered
        perm = rng.permutation(n_dst)
        centre = (perm[dst] * n_src) // n_dst
    src = np.clip(centre + rng.integers(-12, 12, size=dst.size), 0, n_src - 1)
    w = rng.random((dst.size, 1))
    p = HostPlan(smm_lib, src + 1, dst + 1, w, n_src, n_dst)
    assert p.info["kernel_name"] == "staged" and p.info["packed_rows"] == 1, p.info
    assert (p.info["lanes_per_row"], p.info["links_per_lane"]) == (1, 16)
    assert bool(p.info["rows_reordered"]) == reorder
    _check_plan_invariants(p, n_src)
    rs = p.rowslot
    assert np.array_equal(np.sort(rs[rs >= 0]), np.arange(n_dst))        # every row exactly once
    assert not (rs[:, 0] == -2).any()                                     # a continuation follows a row
    need = np.maximum(1, -(-np.diff(p.rowptr) // 4))                      # sub-rows per row
    assert (rs == -2).sum() == (need - 1).sum()
    # padding slots carry weight 0; links sit in their row's sub-rows in ascending-src order
    assert np.count_nonzero(p.wplan) == np.count_nonzero(p.val)
    mat = oracle.compute_weights_matrix_c(src + 1, dst + 1, w, n_src, n_dst)
    x = rng.standard_normal((B, n_src)) + 3
    assert_parity(p.emulate(x), oracle.apply_weights_c(x, mat, None, None, 0.0, False), 1e-12)


def test_packed_rows_fall_back_when_footprint_too_large(smm_lib, oracle):
    """A packed tile holds 4x the rows; when its source footprint does not fit the stages the
    lane-per-row layout (smaller tiles) is used instead."""
    rng = np.random.default_rng(3)
    n_dst, k = 2048, 4
    n_src = n_dst * 40                                   # 40 source columns per destination row
    dst = np.repeat(np.arange(n_dst), k)
    src = dst * 40 + rng.integers(0, 40, size=dst.size)   # every tile row touches its own 40-column block
    w = rng.random((dst.size, 1))
    p = HostPlan(smm_lib, src + 1, dst + 1, w, n_src, n_dst)
    assert p.info["kernel_name"] == "staged" and p.info["packed_rows"] == 0, p.info
    assert p.info["max_tile_elems"] <= 25600
    mat = oracle.compute_weights_matrix_c(src + 1, dst + 1, w, n_src, n_dst)
    x = rng.standard_normal((2, n_src))
    assert_parity(p.emulate(x), oracle.apply_weights_c(x, mat, None, None, 0.0, False), 1e-12)


def test_compact_plan_of_scattered_operator(smm_lib):
    """Two-pass compact plan (uploaded for gather-family levels): touched columns ascending, block
    ranges over 256-column blocks, link columns replaced by their rank."""
    from smmregrid_b200 import _lib
    rng = np.random.default_rng(4)
    n_src, n_dst = 100_003, 700
    src, dst, w = random_links(rng, n_src, n_dst, 3, dup_frac=0.1, sort=False)
    s32, d32 = np.ascontiguousarray(src, np.int32), np.ascontiguousarray(dst, np.int32)
    h = ctypes.c_void_p()
    _lib.check(smm_lib.smm_host_plan_build(n_src, n_dst, s32.size, s32.ctypes.data, d32.ctypes.data,
                                           w.ctypes.data, 1, 1, None, ctypes.byref(h)))
    try:
        inf = _lib.SmmInfo()
        _lib.check(smm_lib.smm_host_plan_info(h, ctypes.byref(inf), None))
        rowptr = np.empty(n_dst + 1, np.int32); col = np.empty(inf.nnz, np.int32); val = np.empty(inf.nnz)
        _lib.check(smm_lib.smm_host_plan_copy(h, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data, None, None, None, None))
        nt, nb = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(smm_lib.smm_host_plan_compact(h, ctypes.byref(nt), ctypes.byref(nb), None, None, None))
        assert nt.value == inf.touched_src == np.unique(col).size and nb.value == -(-n_src // 256)
        tcols = np.empty(nt.value, np.int32); blk = np.empty(nb.value + 1, np.int32); rcol = np.empty(inf.nnz, np.int32)
        _lib.check(smm_lib.smm_host_plan_compact(h, None, None, tcols.ctypes.data, blk.ctypes.data, rcol.ctypes.data))
        assert np.array_equal(tcols, np.unique(col))
        assert np.array_equal(tcols[rcol], col)                                  # rank -> column round trip
        assert blk[0] == 0 and blk[-1] == nt.value and (np.diff(blk) >= 0).all()
        for i in (0, 7, nb.value - 1):                                           # block i holds columns [256 i, 256 (i + 1))
            c = tcols[blk[i]:blk[i + 1]]
            assert ((c >= 256 * i) & (c < 256 * (i + 1))).all()
            assert c.size == ((np.unique(col) // 256) == i).sum()
    finally:
        smm_lib.smm_host_plan_free(h)


@pytest.mark.parametrize("pattern", ["all16", "all0_but_one", "alt_1_13", "ramp", "five_everywhere"])
def test_packed_rows_corner_cases(smm_lib, oracle, pattern):
    """Packing corner cases: one row per thread (16 links), almost empty operators, rows whose
    sub-row counts do not divide a thread's four sub-rows, more rows than one tile holds."""
    rng = np.random.default_rng(len(pattern))
    n_dst = 2600                                       # > 1024 rows: several tiles even with 4 rows per thread
    counts = {"all16": np.full(n_dst, 16), "all0_but_one": np.zeros(n_dst, np.int64),
              "alt_1_13": np.where(np.arange(n_dst) % 2 == 0, 1, 13), "ramp": np.arange(n_dst) % 17,
              "five_everywhere": np.full(n_dst, 5)}[pattern]
    if pattern == "all0_but_one":
        counts[1234] = 3
    n_src = 40 * 1024
    dst = np.repeat(np.arange(n_dst), counts)
    within = np.concatenate([np.arange(c) for c in counts]) if dst.size else np.zeros(0, np.int64)
    src = np.clip((dst * n_src) // n_dst + 2 * within + rng.integers(0, 2, size=dst.size), 0, n_src - 1)
    w = rng.random((dst.size, 1)) + 0.1
    p = HostPlan(smm_lib, src + 1, dst + 1, w, n_src, n_dst)
    if pattern == "all0_but_one":          # three links in all: not worth staging, the gather kernel serves it
        assert p.info["kernel_name"] == "gather" and p.info["nnz"] == 3
        return
    assert p.info["kernel_name"] == "staged" and p.info["packed_rows"] == 1, p.info
    _check_plan_invariants(p, n_src)
    rs = p.rowslot
    assert np.array_equal(np.sort(rs[rs >= 0]), np.arange(n_dst))
    need = np.maximum(1, -(-np.diff(p.rowptr) // 4))
    assert (rs == -2).sum() == (need - 1).sum()
    # a continuation always follows its row inside the same thread
    for u in (1, 2, 3):
        cont = rs[:, u] == -2
        assert ((rs[:, u - 1] >= 0) | (rs[:, u - 1] == -2))[cont].all()
    if pattern == "all16":
        assert p.info["n_tiles"] == -(-n_dst // 256)     # one row per thread
    if pattern == "five_everywhere":
        assert p.info["n_tiles"] == -(-n_dst // 512)     # two rows (2 + 2 sub-rows) per thread
    mat = oracle.compute_weights_matrix_c(src + 1, dst + 1, w, n_src, n_dst)
    x = rng.standard_normal((2, n_src)) + 2
    assert_parity(p.emulate(x), oracle.apply_weights_c(x, mat, None, None, 0.0, False), 1e-12, pattern)


@pytest.mark.parametrize("nnz_per_row", [2, 7, 12, 30, 60, 120, 200])
def test_reference_order_plan_is_bit_identical_to_the_oracle(smm_lib, oracle, nnz_per_row):
    """Operators with negative weights are planned for reference-order summation (SMM_SUM_AUTO):
    emulating the ORD kernels' chain over the plan arrays reproduces the oracle BIT FOR BIT,
    for lane-per-row and packed layouts; a positive operator keeps the fast placement unless
    SMM_SUM_REFERENCE is asked for."""
    rng = np.random.default_rng(40 + nnz_per_row)
    n_src, n_dst, B = 3000, 500, 5
    counts = rng.integers(0, nnz_per_row + 1, size=n_dst)
    counts[3] = nnz_per_row
    dst = np.repeat(np.arange(n_dst), counts)
    src = np.clip((dst * n_src) // n_dst + rng.integers(-200, 200, size=dst.size), 0, n_src - 1)
    w = rng.standard_normal(dst.size)                      # both signs: sums cancel
    o = np.lexsort((src, dst))
    src, dst, w = (src[o] + 1).astype(np.int32), (dst[o] + 1).astype(np.int32), w[o].reshape(-1, 1)
    x = rng.standard_normal((B, n_src)) * 50
    plan = HostPlan(smm_lib, src, dst, w, n_src, n_dst)
    assert plan.info["summation"] == 2 and plan.info["kernel_name"] == "staged"
    if nnz_per_row > 16:
        # rows beyond the packed layout: a thread per row, its links in shared memory (ordered_kernel):
        # one "lane" per row, as many link slots as the longest row, as many rows per tile as fit
        assert plan.info["lanes_per_row"] == 1 and plan.info["links_per_lane"] == plan.info["max_row_nnz"]
        assert plan.info["consumer_threads"] in (256, 128, 64, 32)
        assert plan.info["links_per_lane"] * plan.info["consumer_threads"] * 12 <= 100 * 1024
    else:
        assert plan.info["consumer_threads"] == 256 and plan.info["packed_rows"] == 1
    mat = oracle.compute_weights_matrix_c(src, dst, w, n_src, n_dst)
    y_ref = oracle.apply_weights_c(x, mat, None, None, 0.0, False)
    y = plan.emulate_reference_order(x)
    assert np.array_equal(y, y_ref), float(np.abs(y - y_ref).max())
    # the same operator forced to the fast placement, and a positive one forced to reference order
    assert HostPlan(smm_lib, src, dst, w, n_src, n_dst, summation=1).info["summation"] == 1
    wp = np.abs(w)
    assert HostPlan(smm_lib, src, dst, wp, n_src, n_dst).info["summation"] == 1
    planp = HostPlan(smm_lib, src, dst, wp, n_src, n_dst, summation=2)
    matp = oracle.compute_weights_matrix_c(src, dst, wp, n_src, n_dst)
    assert np.array_equal(planp.emulate_reference_order(x), oracle.apply_weights_c(x, matp, None, None, 0.0, False))


def test_plan_cache_round_trip(smm_lib, tmp_path):
    """On-disk cache of the operator construction: a second build of the same links is a hit and
    returns byte-identical CSR + plan arrays; different links, a different summation order, a
    truncated or foreign file are misses that rebuild (and repair) silently."""
    import os
    from smmregrid_b200 import synth
    w = synth.config_weights("C4", 10)
    n_src, n_dst = w.sizes["src_grid_size"], w.sizes["dst_grid_size"]
    args = (smm_lib, w["src_address"], w["dst_address"], w["remap_matrix"], n_src, n_dst)
    d = str(tmp_path / "plans")
    a = HostPlan(*args, cache_dir=d)
    assert a.info["plan_cache_hit"] == 0
    files = os.listdir(d)
    assert len(files) == 1 and files[0].endswith(".plan")
    b = HostPlan(*args, cache_dir=d)
    assert b.info["plan_cache_hit"] == 1
    for name in ("rowptr", "col", "val", "tiles", "segs", "wplan", "iplan"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    assert {k: v for k, v in a.info.items() if k != "plan_cache_hit"} == \
        {k: v for k, v in b.info.items() if k != "plan_cache_hit"}
    # other summation order -> other key; other weights -> other key
    assert HostPlan(*args, summation=2, cache_dir=d).info["plan_cache_hit"] == 0
    rm2 = np.array(w["remap_matrix"]); rm2[0, 0] *= 0.5
    assert HostPlan(smm_lib, w["src_address"], w["dst_address"], rm2, n_src, n_dst, cache_dir=d).info["plan_cache_hit"] == 0
    assert len(os.listdir(d)) == 3
    # a damaged file is a miss and is rewritten
    path = os.path.join(d, files[0])
    size = os.path.getsize(path)
    with open(path, "r+b") as f:
        f.truncate(size // 2)
    c = HostPlan(*args, cache_dir=d)
    assert c.info["plan_cache_hit"] == 0 and np.array_equal(c.wplan, a.wplan)
    assert os.path.getsize(path) == size
    assert HostPlan(*args, cache_dir=d).info["plan_cache_hit"] == 1
    # an unwritable directory only disables the cache
    assert HostPlan(*args, cache_dir="/proc/definitely/not/writable").info["plan_cache_hit"] == 0


def test_split_plan_for_mostly_short_rows(smm_lib, oracle):
    """A tripolar (north-fold) ocean grid: nearly all destination rows have a handful of links, the
    cells next to the grid poles collect dozens.  The plan packs the short rows (a thread owns up
    to four of them) and lists the long ones for the gather kernel instead of laying EVERY row out
    for the longest one; emulating both parts reproduces the oracle.  Operators whose rows are all
    long (remapcon 0.1 -> 1 deg) or all short are planned as before."""
    from smmregrid_b200 import synth
    w = synth.tripolar_weights(181, 146, 180, 90)
    n_src, n_dst = w.sizes["src_grid_size"], w.sizes["dst_grid_size"]
    plan = HostPlan(smm_lib, w["src_address"], w["dst_address"], w["remap_matrix"], n_src, n_dst)
    info = plan.info
    counts = np.bincount(w["dst_address"] - 1, minlength=n_dst)
    assert info["max_row_nnz"] > 16 and info["kernel_name"] == "staged" and info["packed_rows"] == 1
    assert info["gather_rows"] == int((counts > 16).sum()) > 0
    assert np.array_equal(plan.gather_rows, np.flatnonzero(counts > 16))
    listed = plan.rowslot[plan.rowslot >= 0]
    assert np.array_equal(np.sort(listed), np.flatnonzero(counts <= 16))        # every short row exactly once
    x = np.random.default_rng(0).standard_normal((3, n_src))
    mat = oracle.compute_weights_matrix_c(w["src_address"], w["dst_address"], w["remap_matrix"], n_src, n_dst)
    y_ref = oracle.apply_weights_c(x, mat, None, None, 0.0, False)
    assert_parity(plan.emulate(x), y_ref, 1e-12, "split plan")
    # unchanged decisions elsewhere
    for cfg, scale, packed in (("C4", 10, 0), ("C2", 8, 0), ("C1", 1, 1)):
        wc = synth.config_weights(cfg, scale)
        pc = HostPlan(smm_lib, wc["src_address"], wc["dst_address"], wc["remap_matrix"],
                      wc.sizes["src_grid_size"], wc.sizes["dst_grid_size"])
        assert pc.info["gather_rows"] == 0 and pc.info["packed_rows"] == packed, cfg
