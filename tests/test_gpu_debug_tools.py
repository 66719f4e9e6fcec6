"""Developer tools that live in libdeepsc_b200_debug.so (built by __graft_entry__.build() next to the product library): they
run in a subprocess because a process binds one of the two libraries."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEBUG_LIB = os.path.join(ROOT, "deepsc-gan_b200", "csrc", "libdeepsc_b200_debug.so")

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not os.path.exists(DEBUG_LIB), reason="debug-tools library not built")
def test_two_tile_star_kernel_is_bit_identical_to_the_product_kernel():
    """tools/pp_check.py: the experimental two-tile form of the fused star layer (csrc/debug/dsc_star_pp.cu) against the
    product's one-tile kernel - which the other tests pin on the oracle - for every flag combination (first satellite half
    cached, no final relay), 1 / 2 / 3 / 8 cycles, 0 / 17 target keys and 2 .. 593 tiles (odd counts, one CTA, 2 / 4 / 5
    tiles per CTA), bf16x3 and bf16: results must be EQUAL."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "pp_check.py")], capture_output=True, text=True, timeout=600)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-2000:]
    assert "MISMATCH" not in out and "bit-identical" in out, out[-2000:]
    assert out.count("identical: True") == 2, out[-2000:]


def test_product_library_rejects_the_two_tile_form():
    import torch
    import deepsc_gan_b200  # noqa: F401
    from deepsc_gan_b200 import _lib
    lib = _lib.load()
    q = torch.zeros(1 << 16, device="cuda")
    p = q.data_ptr()
    rc = lib.dsc_star_cycles_tc(p, p, p, p, None, 0, p, p, p, p, p, p, p, p, 8, 2, 1 | _lib.STAR_FORM_TWO_TILE, None)
    assert rc == -1 and b"debug" in lib.dsc_last_error()
