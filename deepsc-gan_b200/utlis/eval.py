"""Evaluation steps: mirror of DeepSC-GAN/utlis/eval.py.

``greedy_decode_noattack`` keeps the reference's signature and result (ids [bs, 31] int32).  The
reference treats the whole batch as one unit (batch-global power norm, one fading coefficient), and so
does this function; the multi-unit form used by the SNR sweep is ``engine.greedy_units``.

The FGM/PGD evaluators (greedy_decode, greedy_decode_gan, eval_step_normal, eval_step_star,
eval_step_FGM, eval_step_normal_pgd) need the gradient of the loss with respect to the channel symbols,
i.e. the backward kernels, which are not in this revision: they raise
NotImplementedError rather than fall back to autograd on another backend.
"""
from __future__ import annotations

import math

import torch

from .. import _lib, engine
from ..models.modules import create_look_ahead_mask, create_masks, create_padding_mask, loss_function
from .tools import BleuScore, SeqtoText, SNR_to_noise


def greedy_decode_noattack(args, inp, net, PNR_dB, channel='AWGN', n_std=0.1, epsilon=1, *, noise=None, h=None,
                           seed: int = 0, decoder=None):
    """utlis/eval.py:78-117.  ``noise`` = injected unit-normal tensor [bs,31,16]; ``h`` = the two
    unit-normal draws of the fading coefficient (z1, z2).  AWGN is the inline form without sqrt(size)
    (:90-93); p is zero so PNR_dB has no effect, as in the reference."""
    dev = inp.device
    ns = torch.full((1,), float(n_std), device=dev, dtype=torch.float32)
    hh = None
    if channel != 'AWGN':
        K = 1 if channel == 'Rician' else 0
        mean, std = math.sqrt(K / (2 * (K + 1))), math.sqrt(1 / (2 * (K + 1)))
        z = torch.randn(2).tolist() if h is None else [float(h[0]), float(h[1])]
        hh = torch.tensor([[mean + std * z[0], mean + std * z[1]]], device=dev, dtype=torch.float32)
    out = engine.greedy_units(net, inp, 1, ns, channel=channel, noise=None if noise is None else noise.contiguous(),
                              seed=seed, h=hh, max_length=args.max_length, start_idx=args.start_idx, decoder=decoder)
    return out.clone()


def _needs_backward(name):
    def fn(*a, **k):
        raise NotImplementedError(f"{name} needs d(loss)/d(symbols), i.e. the backward kernels (SURVEY.md K17); "
                                  "not implemented in this revision, and no autograd fallback is provided")
    fn.__name__ = name
    return fn


greedy_decode = _needs_backward("greedy_decode")
greedy_decode_gan = _needs_backward("greedy_decode_gan")
eval_step_normal = _needs_backward("eval_step_normal")
eval_step_normal_pgd = _needs_backward("eval_step_normal_pgd")
eval_step_star = _needs_backward("eval_step_star")
eval_step_FGM = _needs_backward("eval_step_FGM")
