"""Host-side pieces of the front end that need no GPU."""
import numpy as np
import torch


def test_pinned_result_pool_recycles_only_dropped_results(monkeypatch):
    """Results are handed out in recycled page-locked buffers (here: ordinary memory, no GPU): only
    for a size that was asked for before (a loop), a buffer is reused only after every view /
    numpy array of the previous result is gone, a result still held gets the second buffer of its
    size, a third concurrent result falls back to pageable memory (None)."""
    import sys
    import smmregrid_b200  # noqa: F401
    mod = sys.modules["smmregrid_b200.regrid"]
    pool = mod._PinnedResults()
    monkeypatch.setattr(pool, "_alloc", lambda nbytes: torch.empty(nbytes, dtype=torch.uint8))
    pool.min_bytes = 1024
    assert pool._users(torch.empty(8, dtype=torch.uint8)) >= 1
    assert pool.take((4, 4), np.float64) is None                       # too small
    assert pool.take((4, 1024), np.float64) is None                    # first request of this size: not a loop yet
    a = pool.take((4, 1024), np.float64)
    assert a.shape == (4, 1024) and a.dtype == torch.float64 and len(pool._bufs) == 1
    ptr_a = a.data_ptr()
    an = a.numpy()
    del a
    b = pool.take((4, 1024), np.float64)                                # `an` still refers to the first buffer
    assert b.data_ptr() != ptr_a and len(pool._bufs) == 2
    view = an.reshape(-1)[10:]
    del an
    assert pool.take((4, 1024), np.float64) is None                     # a derived numpy view still holds the first, b the second
    del view
    d = pool.take((4, 1024), np.float64)                                # first buffer is free again
    assert d.data_ptr() == ptr_a and len(pool._bufs) == 2
    del b, d
    pool.take((3, 1024), np.float64)
    e = pool.take((3, 1024), np.float64)                                # a slightly smaller result reuses a buffer
    assert e.data_ptr() == ptr_a and e.shape == (3, 1024) and len(pool._bufs) == 2
    pool.cap_bytes = 4 * 1024 * 8 * 2                                   # the cap stops a new size class
    pool.take((64, 1024), np.float64)
    assert pool.take((64, 1024), np.float64) is None


def test_new_host_result_prefault_keeps_contents():
    """The helper thread only asks the kernel to populate the pages: data written meanwhile stays."""
    import sys
    import time
    import smmregrid_b200  # noqa: F401
    mod = sys.modules["smmregrid_b200.regrid"]
    a = mod._new_host_result((64, 64800), np.float64)                   # 33 MB: above the threshold
    a[:] = 7.0
    time.sleep(0.2)
    assert a.shape == (64, 64800) and np.all(a == 7.0)
