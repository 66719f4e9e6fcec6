import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import deepsc_gan_b200  # noqa: F401
    from deepsc_gan_b200 import _lib
    _lib.load()            # fail loudly if the extension is missing: no fallback
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _default_precision():
    """Every test starts and ends in the package's default arithmetic (prec 1, tcgen05 bf16x3); a test that wants the fp32
    debug mode or the single-pass bf16 mode selects it explicitly."""
    import deepsc_gan_b200.models.modules as Mod
    Mod.set_precision(1)
    yield
    Mod.set_precision(1)
