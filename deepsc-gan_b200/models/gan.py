"""Perturbation generator: mirror of DeepSC-GAN/models/gan.py.

Only ``G`` is on the path (``Transeiver_GAN.generator``, models/transceiver.py:261).  ``D``, ``D_CNN``
and ``G_CNN`` are defined by the reference but never instantiated; they are out of scope.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import _lib
from .modules import Dense


class G(nn.Module):
    """models/gan.py:4-16: Dense(256, relu) -> Dense(16) -> x / sqrt(2 * mean(x^2)) over the whole
    tensor, i.e. the perturbation carries half the signal power."""

    def __init__(self, size1: int = 256, size2: int = 16):
        super().__init__()
        self.fc0 = Dense(size2, size1, activation="relu")
        self.fc1 = Dense(size1, size2)

    def raw(self, inputs: torch.Tensor) -> torch.Tensor:
        """fc1(fc0(x)) before the power budget (the fused channel kernel can apply the budget itself)."""
        return self.fc1(self.fc0(inputs))

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        from . import modules as _M
        from .. import autograd as AG
        if _M.is_differentiable():
            return AG.PowerNormalize.apply(self.raw(inputs), 1, 2.0)
        g = self.raw(inputs).contiguous()
        return _lib.power_normalize(g, 1, factor=2.0)

    call = forward
