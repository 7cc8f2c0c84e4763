"""C-ABI library: loads, exports every symbol the header declares, fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "smmregrid_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(smm_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported(smm_lib):
    names = _declared_symbols()
    assert len(names) >= 15
    raw = ctypes.CDLL(os.path.join(ROOT, "smmregrid_b200", "lib", "libsmmregrid_b200.so"))
    for n in names:
        assert hasattr(raw, n), f"{n} declared in the header but not exported"


def test_binding_covers_header(smm_lib):
    from smmregrid_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_version_and_error_string(smm_lib):
    assert b"sm_100a" in smm_lib.smm_version()
    assert isinstance(smm_lib.smm_last_error(), bytes)
    assert smm_lib.smm_launch_count() >= 0


def test_library_is_sm100a_only():
    """The shipped binary holds sm_100a code (and nothing older)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    so = os.path.join(ROOT, "smmregrid_b200", "lib", "libsmmregrid_b200.so")
    out = subprocess.run([cuobjdump, "--list-elf", so], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback(smm_lib):
    """Without a CUDA device every compute entry point refuses; nothing routes to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from smmregrid_b200 import _lib, Regridder, synth
    src = np.array([1, 2], np.int32)
    dst = np.array([1, 1], np.int32)
    w = np.array([[0.5], [0.5]])
    h = ctypes.c_void_p()
    rc = smm_lib.smm_create(2, 1, 2, src.ctypes.data, dst.ctypes.data, w.ctypes.data, 1, 1, 0, None, ctypes.byref(h))
    assert rc == _lib.SMM_ERR_CUDA and not h.value
    assert b"no CPU fallback" in smm_lib.smm_last_error() or b"CUDA" in smm_lib.smm_last_error()
    with pytest.raises(_lib.SmmError):
        Regridder(weights=synth.config_weights("C1"))


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: the package must not reference it."""
    pkg = os.path.join(ROOT, "smmregrid_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("oracle-free", ""), os.path.join(dirpath, f)


def test_argument_validation_without_device(smm_lib):
    from smmregrid_b200 import _lib
    h = ctypes.c_void_p()
    ll = np.array([5], np.int64)
    rc = smm_lib.smm_create_levels(1, ll.ctypes.data, 3, 4, 4, None, None, None, 1, 1, 0, None, ctypes.byref(h))
    assert rc == _lib.SMM_ERR_INVALID          # link_length > nl_max
    rc = smm_lib.smm_create_levels(0, ll.ctypes.data, 3, 4, 4, None, None, None, 1, 1, 0, None, ctypes.byref(h))
    assert rc == _lib.SMM_ERR_INVALID
    assert smm_lib.smm_destroy(None) == 0
    rc = smm_lib.smm_apply(None, 0, None, 0, 1, 1, None, 0, 1, 0, 0.5, None, None)
    assert rc == _lib.SMM_ERR_INVALID
