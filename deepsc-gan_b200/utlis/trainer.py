"""Training steps: mirror of DeepSC-GAN/utlis/trainer.py (train_step_noattack :12-27, train_attack_step :30-64).

``optim_net`` is an ``optim.Adam`` over the model's ``optim.FlatParams`` (see ``make_optimizer``).  Forward and backward
run on the libdeepsc_b200.so kernels through ``models.modules.differentiable()``; with a process group initialised the
flat gradient bucket is all-reduced over NCCL before the update (data parallel, one 64-sentence unit per rank).
Keyword-only additions: ``noise`` / ``noise2`` (injected unit-normal channel draws), ``h``.
"""
from __future__ import annotations

import torch

from .. import _lib
from .. import optim as O
from ..models.modules import create_masks, differentiable, loss_function, precision
from .eval import GRADIENT_PREC, fgm_perturbation


def make_optimizer(net, learning_rate: float = 5e-4, n_grad_buffers: int = 2, **adam_kw) -> O.Adam:
    """Flatten ``net``'s parameters and build the Adam the reference's (missing) driver would pass as ``optim_net``."""
    return O.Adam(O.FlatParams(net, n_grad_buffers=n_grad_buffers), learning_rate=learning_rate, **adam_kw)


def _all(_name: str) -> bool:
    return True


def _forward(net, inp, tar_inp, p, PNR_dB, channel, n_std, masks, **kw):
    enc_padding_mask, combined_mask, dec_padding_mask = masks
    return net(inp, tar_inp, p, PNR_dB, channel=channel, n_std=n_std, training=True, enc_padding_mask=enc_padding_mask,
               combined_mask=combined_mask, dec_padding_mask=dec_padding_mask, **kw)


def train_step_noattack(inp, tar, p, net, optim_net, channel='AWGN', n_std=0.1, train_with_mine=False, epsilon=1, *,
                        noise=None, h=None):
    """utlis/trainer.py:12-27: masks, forward at PNR_dB = 0 with training=True, CE, gradients on every variable, apply."""
    tar_inp, tar_real = tar[:, :-1], tar[:, 1:]
    masks = create_masks(inp, tar_inp)
    fp = optim_net.fp
    fp.grad_bucket[0].zero_()
    fp.point_grads(0)
    with differentiable():
        outs = _forward(net, inp, tar_inp, p, 0, channel, n_std, masks, noise=noise, h=h)
        loss = loss_function(tar_real, outs[0])
        loss.backward()
    scale = O.all_reduce_mean_scale(fp.grad_bucket[0])
    optim_net.apply(fp.ranges(_all), fp.grad_bucket[0], scale)
    return loss.detach()


def train_attack_step(inp, tar, p, PNR_dB, net, optim_net, channel='AWGN', n_std=0.1, train_with_mine=False, epsilon=1,
                      *, noise=None, noise2=None, h=None):
    """utlis/trainer.py:30-64 (star models: the loss target is the full ``tar``, :32): forward, d loss/d received
    symbols, FGM normalisation (:45-53), second forward with the perturbation, backward, apply."""
    tar_inp, tar_real = tar[:, :-1], tar
    masks = create_masks(inp, tar_inp)
    fp = optim_net.fp
    with differentiable(), precision(GRADIENT_PREC):       # the FGM direction is taken in fp32 (utlis/eval.py GRADIENT_PREC)
        outs = _forward(net, inp, tar_inp, p, PNR_dB, channel, n_std, masks, noise=noise, h=h)
        loss = loss_function(tar_real, outs[0])
        (g,) = torch.autograd.grad(loss, outs[3])
    r_list = fgm_perturbation(g, epsilon)
    fp.grad_bucket[0].zero_()
    fp.point_grads(0)
    with differentiable():
        outs = _forward(net, inp, tar_inp, r_list, PNR_dB, channel, n_std, masks, noise=noise2, h=h)
        loss_m = loss_function(tar_real, outs[0])
        loss_m.backward()
    scale = O.all_reduce_mean_scale(fp.grad_bucket[0])
    optim_net.apply(fp.ranges(_all), fp.grad_bucket[0], scale)
    return loss.detach(), loss_m.detach()
