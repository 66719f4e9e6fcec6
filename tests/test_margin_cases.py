"""The margin-enforced greedy cases (tests/golden/margin_*.npz) are what the strict token-id tests on the GPU compare
against.  Here, on the CPU: the fixtures are reproducible from the oracle, and every decoded step of every sentence has an
fp64 top-2 logit margin of at least TEST_MARGIN x max|logit| - twenty times the error the tensor-core kernels are allowed
(1e-3 relative is the contract, 2e-4 is asserted at prec 1) - so a differing id on the GPU is a defect, never a tie."""
import numpy as np
import pytest
import torch

import _cases

TEST_MARGIN = 2e-3


@pytest.mark.parametrize("name", list(_cases.MARGIN_CASES))
def test_margin_case_is_reproducible_and_margin_clean(name):
    fx = np.load(_cases.margin_path(name))
    inp = torch.from_numpy(fx["inp"])
    assert inp.shape == (64, 31) and fx["ids"].shape == (64, 31) and fx["counts"].shape == (64, 10)
    ids64, margin = _cases.margin_oracle(name, inp, fx["seeds"], torch.float64)
    assert np.array_equal(ids64.numpy(), fx["ids"])
    assert float(margin.min()) >= TEST_MARGIN, float(margin.min())
    assert np.allclose(margin.numpy(), fx["margin"], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("name", ["Transeiver_AWGN", "Transeiver_Star_Rayleigh"])
def test_fp32_oracle_agrees_with_fp64_on_margin_cases(name):
    fx = np.load(_cases.margin_path(name))
    ids32, _ = _cases.margin_oracle(name, torch.from_numpy(fx["inp"]), fx["seeds"], torch.float32)
    assert np.array_equal(ids32.numpy(), fx["ids"])


@pytest.mark.parametrize("channel", ["AWGN", "Rayleigh"])
def test_sweep_margin_case_is_reproducible_and_margin_clean(channel):
    """The 2 units x 3 SNR points case of tests/test_gpu_sweep.py: every (SNR point, unit) item against the fp64 oracle."""
    fx = np.load(_cases.sweep_margin_path(channel))
    units = torch.from_numpy(fx["units"])
    items = [(s, u) for s in range(len(_cases.SWEEP_SNRS)) for u in range(_cases.SWEEP_UNITS)]
    P = _cases.params("Transeiver_Star")
    for i, (s, u) in enumerate(items):
        ids, margin = _cases.greedy_with_margin("Transeiver_Star", P, units[64 * u:64 * u + 64], channel, _cases.SWEEP_SNRS[s],
                                                None, fx["seeds"][i], torch.float64)
        assert np.array_equal(ids.numpy(), fx["ids"][64 * i:64 * i + 64])
        assert float(margin.min()) >= TEST_MARGIN
