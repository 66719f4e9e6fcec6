"""The C-ABI library loads without a GPU and exports every symbol include/deepsc_b200.h declares."""
import ctypes
import os
import re

import deepsc_gan_b200  # noqa: F401
from deepsc_gan_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "deepsc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dsc_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"{n} declared in deepsc_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), "ctypes prototypes out of sync with the header"


def test_version_and_error_channel():
    lib = _lib.load()
    assert lib.dsc_version() >= 100
    assert isinstance(lib.dsc_last_error(), bytes)
    # argument validation happens before any CUDA call, so it can be exercised without a GPU
    rc = lib.dsc_linear(None, 128, None, 128, None, None, 128, 4, 128, 128, 0, 0, 0, 0, None)
    assert rc == -1 and b"null pointer" in lib.dsc_last_error()
    rc = lib.dsc_channel(16, None, 1.0, None, 0, 0, None, None, 1.0, None, None, 16, 7, 16, None, 1, 64, None)
    assert rc == -1 and lib.dsc_last_error() == b"detector must in LS and MMSE"


def test_new_entry_points_validate_before_touching_the_gpu():
    """dsc_gemm_nt_tc / dsc_transpose / the star-cycle flags reject bad arguments with DSC_ERR_BAD_ARG (-1) and a message;
    none of this needs a device."""
    lib = _lib.load()
    assert lib.dsc_gemm_nt_tc(None, 128, None, 128, None, 128, 4, 4, 128, 0, None) == -1
    assert b"null pointer" in lib.dsc_last_error()
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.dsc_gemm_nt_tc(p, 127, p, 128, p, 128, 4, 4, 126, 0, None) == -1          # odd leading dimension
    assert b"8-byte aligned" in lib.dsc_last_error()
    assert lib.dsc_gemm_nt_tc(p, 128, p, 128, p, 2, 4, 4, 128, 0, None) == -1            # ldc < N
    assert lib.dsc_transpose(p, 2, p, 8, 8, 8, None) == -1                               # ld_src < cols
    # DSC_STAR_FIRST_SAT_DONE needs at least two cycles (the cached half is cycle 0's)
    big = (ctypes.c_float * 4096)()
    q = ctypes.cast(ctypes.addressof(big) + (-ctypes.addressof(big)) % 128, ctypes.c_void_p)
    rc = lib.dsc_star_cycles_tc(q, q, q, q, None, 0, q, q, q, q, q, q, q, q, 4, 1, 1 | _lib.STAR_FIRST_SAT_DONE, None)
    assert rc == -1 and b"n_cycles >= 2" in lib.dsc_last_error()
    rc = lib.dsc_star_cycles_tc(q, q, q, q, None, 0, q, q, q, q, q, q, q, q, 6, 8, 1, None)
    assert rc == -1 and b"multiple of 4" in lib.dsc_last_error()
    # the two kernel-form flags exclude each other
    rc = lib.dsc_star_cycles_tc(q, q, q, q, None, 0, q, q, q, q, q, q, q, q, 8, 2,
                                1 | _lib.STAR_FORM_ONE_TILE | _lib.STAR_FORM_TWO_TILE, None)
    assert rc == -1 and b"at most one kernel form" in lib.dsc_last_error()
    # the product library has one kernel form: the experimental two-tile kernel is in the debug-tools library only
    rc = lib.dsc_star_cycles_tc(q, q, q, q, None, 0, q, q, q, q, q, q, q, q, 8, 2, 1 | _lib.STAR_FORM_TWO_TILE, None)
    assert rc == -1 and b"libdeepsc_b200_debug.so" in lib.dsc_last_error()


def test_product_has_no_cpu_fallback():
    import pytest
    import torch
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.unit_sumsq(torch.zeros(64), 1)
