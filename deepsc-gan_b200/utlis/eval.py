"""Evaluation steps: mirror of DeepSC-GAN/utlis/eval.py (same names, positional arguments and return arity).

The reference treats the whole batch as one unit (batch-global power norm, one fading coefficient, FGM global
norm) and so do these functions; the multi-unit forms used by the SNR sweep are ``engine.greedy_units`` and
``sweep.SweepRunner``.

Gradients of the loss with respect to the channel symbols come from the backward kernels through
``models.modules.differentiable()``; the per-sample / global FGM normalisation is one kernel (dsc_fgm_normalize)
with no host synchronisation, where the reference loops over 64 samples in Python (utlis/eval.py:36-44).

Keyword-only additions (injected randomness): ``noise`` / ``noise2`` = unit-normal tensors [bs,31,16] for the
first (clean) and the second (attacked) channel pass; ``h`` = the two unit-normal draws of the fading coefficient.
Reference defects repaired by intent: decoders return logits only; ``zeros([64,31,16])``
is ``None`` (no perturbation); ``eval_step_FGM`` returns its results (the reference falls off the end).
"""
from __future__ import annotations

import contextlib
import math

import torch

from .. import _lib, engine
from ..models import modules as M
from ..models.modules import create_look_ahead_mask, create_masks, create_padding_mask, differentiable, loss_function
from .tools import BleuScore, SeqtoText, SNR_to_noise


# Arithmetic of the taped pass that yields the FGM direction (see models.modules.precision): fp32.  Every other forward of
# the evaluators (the attacked pass, the bisection forwards, the greedy loops) runs in the package default (tcgen05 bf16x3).
GRADIENT_PREC = 0


@contextlib.contextmanager
def _frozen(net):
    """The evaluators differentiate with respect to the channel symbols only: parameters do not require grad."""
    params = [p for p in net.parameters() if p.requires_grad]
    for p in params:
        p.requires_grad_(False)
    try:
        yield
    finally:
        for p in params:
            p.requires_grad_(True)


def _fading_h(channel, h, dev):
    if channel == 'AWGN':
        return None
    K = 1 if channel == 'Rician' else 0
    mean, std = math.sqrt(K / (2 * (K + 1))), math.sqrt(1 / (2 * (K + 1)))
    z = torch.randn(2).tolist() if h is None else [float(h[0]), float(h[1])]
    return torch.tensor([[mean + std * z[0], mean + std * z[1]]], device=dev, dtype=torch.float32)


def fgm_perturbation(gradients: torch.Tensor, epsilon=1) -> torch.Tensor:
    """utlis/eval.py:36-44: r_b = eps*g_b/||g_b|| per sample, pertutation = r/||r||_F (whole batch = one unit)."""
    return _lib.fgm_normalize(gradients.contiguous(), 1, float(epsilon))


def _call(net, inp, tar_inp, p, PNR_dB, channel, n_std, masks, **kw):
    enc_padding_mask, combined_mask, dec_padding_mask = masks
    return net(inp, tar_inp, p, PNR_dB, channel=channel, n_std=n_std, training=False, enc_padding_mask=enc_padding_mask,
               combined_mask=combined_mask, dec_padding_mask=dec_padding_mask, **kw)


def _symbol_gradient(net, inp, tar_inp, tar_real, PNR_dB, channel, n_std, masks, noise, h, gan=False, wrt="symbols",
                     reforward_awgn=True):
    """First (clean) forward under the tape and d(loss)/d(symbols or received symbols).  ``reforward_awgn``: over a fading
    channel the eval_step_* functions take the direction from a second, AWGN forward (utlis/eval.py:204-211); the greedy
    decoders differentiate the forward of the given channel itself (:25-33, :137-144).
    Returns (loss, predictions, gradient, outs)."""
    with _frozen(net), differentiable(), M.precision(GRADIENT_PREC):
        kw = dict(noise_r=noise, h_r=h, traingan=False) if gan else dict(noise=noise, h=h)
        outs = _call(net, inp, tar_inp, None, PNR_dB, channel, n_std, masks, **kw)
        pred = outs[1] if gan else outs[0]
        loss = loss_function(tar_real, pred)
        if channel != 'AWGN' and reforward_awgn:
            # the attack direction is taken through an AWGN forward (utlis/eval.py:204-211, 336-344, 384-390)
            kw2 = dict(noise_r=noise, traingan=False) if gan else dict(noise=noise)
            outs2 = _call(net, inp, tar_inp, None, PNR_dB, 'AWGN', n_std, masks, **kw2)
            loss_temp = loss_function(tar_real, outs2[1] if gan else outs2[0])
            (g,) = torch.autograd.grad(loss_temp, outs2[2] if gan else outs2[1])
        else:
            target = (outs[3] if wrt == "received" else outs[2]) if gan else (outs[3] if wrt == "received" else outs[1])
            (g,) = torch.autograd.grad(loss, target)
    return loss.detach(), pred.detach(), g, tuple(o.detach() for o in outs)


def eval_step_normal(inp, tar, net, PNR_dB, channel='AWGN', n_std=0.1, epsilon=1, *, noise=None, noise2=None, h=None):
    """utlis/eval.py:189-232 (baseline ``Transeiver``): clean forward, FGM direction from d loss/d symbols,
    attacked forward.  Returns (loss, loss_m, predictions, predictions2)."""
    tar_inp, tar_real = tar[:, :-1], tar[:, 1:]
    masks = create_masks(inp, tar_inp)
    loss, predictions, g, _ = _symbol_gradient(net, inp, tar_inp, tar_real, PNR_dB, channel, n_std, masks, noise, h)
    pertutation = fgm_perturbation(g, epsilon)
    outs = _call(net, inp, tar_inp, pertutation, PNR_dB, channel, n_std, masks, noise=noise2, h=h)
    loss_m = loss_function(tar_real, outs[0])
    return loss, loss_m, predictions, outs[0]


def eval_step_star(inp, tar, net, PNR_dB, channel='AWGN', n_std=0.1, epsilon=1, *, noise=None, noise2=None, h=None):
    """utlis/eval.py:321-365 (star models): as eval_step_normal with the full ``tar`` as the loss target,
    because the star decoder emits 31 positions (:334)."""
    tar_inp = tar[:, :-1]
    masks = create_masks(inp, tar_inp)
    loss, predictions, g, _ = _symbol_gradient(net, inp, tar_inp, tar, PNR_dB, channel, n_std, masks, noise, h)
    pertutation = fgm_perturbation(g, epsilon)
    outs = _call(net, inp, tar_inp, pertutation, PNR_dB, channel, n_std, masks, noise=noise2, h=h)
    loss_m = loss_function(tar, outs[0])
    return loss, loss_m, predictions, outs[0]


def eval_step_FGM(inp, tar, net, PNR_dB, channel='AWGN', n_std=0.1, epsilon=1, *, noise=None, noise2=None,
                  noise2_r=None, h=None):
    """utlis/eval.py:367-408 (``Transeiver_GAN``, traingan=False): the clean branch gives the loss and the
    FGM direction (w.r.t. the received symbols y_r, :391), the attacked pass is scored on the perturbed branch.
    Returns (loss, loss_m, predictions_r, predictions_p_m)."""
    tar_inp, tar_real = tar[:, :-1], tar[:, 1:]
    masks = create_masks(inp, tar_inp)
    loss, predictions_r, g, _ = _symbol_gradient(net, inp, tar_inp, tar_real, PNR_dB, channel, n_std, masks, noise, h,
                                                 gan=True, wrt="received")
    pertutation = fgm_perturbation(g, epsilon)
    outs = _call(net, inp, tar_inp, pertutation, PNR_dB, channel, n_std, masks, traingan=False, noise=noise2, h=h,
                 noise_r=noise2_r, h_r=h)
    loss_m = loss_function(tar_real, outs[0])
    return loss, loss_m, predictions_r, outs[0]


def eval_step_normal_pgd(inp, tar, net, PNR_dB, channel='AWGN', n_std=0.1, epsilon=1, *, noise=None, noises=None,
                         h=None, verbose=True):
    """utlis/eval.py:235-318: FGM direction w.r.t. the received symbols, then 10 bisection steps on the attack
    strength eps in [0, 1].  The bisection state lives on the device (the reference compares ``loss_m - loss``
    on the host every step, :293); the only synchronisation is the final ``print('epsilon=', ...)``.
    ``noises`` = optional list of 10 unit-normal tensors for the bisection forwards.
    Returns (loss_ori, loss_m, predictions, predictions2)."""
    tar_inp, tar_real = tar[:, :-1], tar[:, 1:]
    masks = create_masks(inp, tar_inp)
    dev = inp.device
    loss, predictions, g, _ = _pgd_first(net, inp, tar_inp, tar_real, PNR_dB, channel, n_std, masks, noise, h)
    direction = fgm_perturbation(g, epsilon)                        # r_list / power
    enc = net.semantic_encoder.call(inp, False, masks[0])
    x = net.channel_encoder.call(enc)
    size = float(x.numel())
    PNR = 10 ** (PNR_dB / 10)
    lo = torch.zeros((), device=dev)
    hi = torch.ones((), device=dev)
    eps = (lo + hi) / 2
    found = torch.zeros((), device=dev, dtype=torch.bool)
    last_eps = torch.ones((), device=dev)
    ns = torch.full((1,), float(n_std), device=dev, dtype=torch.float32)
    hh = _fading_h(channel, h, dev)
    predictions2 = None
    loss_m = loss
    for i in range(10):
        z = None if noises is None else noises[i].contiguous()
        if channel == 'AWGN':
            # p = sqrt(size) * eps * r/power; y = x + n + n_std*sqrt(PNR)*p   (:277-280): eps rides in p_scale on the device
            ps = (eps * (float(n_std) * math.sqrt(PNR) * math.sqrt(size))).reshape(1).to(torch.float32)
            y, _ = _lib.channel(x.contiguous(), 1, ns, noise=z, seed=net.channel_layer.seed,
                                offset=net.channel_layer._next_offset(), p=direction, p_scale=ps)
        else:
            y, _ = _lib.channel(x.contiguous(), 1, ns, noise=z, seed=net.channel_layer.seed,
                                offset=net.channel_layer._next_offset(), h=hh, detector=0)
        mem = net.channel_decoder.call(y)
        predictions2 = net.semantic_decoder.call(tar_inp, mem, False, masks[1], masks[2])
        loss_m = loss_function(tar_real, predictions2)
        weaker = (loss_m - loss) < 0                               # :293: attack too weak -> raise the lower bound
        lo = torch.where(weaker, eps, lo)
        hi = torch.where(weaker, hi, eps)
        last_eps = torch.where(weaker, last_eps, eps)              # att.append([eps, loss]) (:299)
        found = found | ~weaker
        eps = (lo + hi) / 2
    loss_m = torch.where(found, loss, loss_m)                       # att[-1][1] is the ORIGINAL loss (:299, :312)
    if verbose:
        print('epsilon=', float(torch.where(found, last_eps, torch.ones_like(last_eps))))
    return loss, loss_m, predictions, predictions2


def _pgd_first(net, inp, tar_inp, tar_real, PNR_dB, channel, n_std, masks, noise, h):
    """First pass of eval_step_normal_pgd: the gradient is w.r.t. the received symbols y of the given channel (:250)."""
    with _frozen(net), differentiable(), M.precision(GRADIENT_PREC):
        outs = _call(net, inp, tar_inp, None, PNR_dB, channel, n_std, masks, noise=noise, h=h)
        loss = loss_function(tar_real, outs[0])
        (g,) = torch.autograd.grad(loss, outs[3])
    return loss.detach(), outs[0].detach(), g, None


def _draw(net, seed):
    """(seed, offset) of the on-device Philox stream for one channel call: every call advances the channel layer's
    counter, so successive batches and SNR points see fresh noise, as tf.random.normal gives the reference
    (utlis/eval.py:90-93).  ``seed`` overrides the layer's seed when given."""
    layer = net.channel_layer
    return (layer.seed if seed is None else seed), layer._next_offset()


def greedy_decode_noattack(args, inp, net, PNR_dB, channel='AWGN', n_std=0.1, epsilon=1, *, noise=None, h=None,
                           seed=None, decoder=None):
    """utlis/eval.py:78-117.  ``noise`` = injected unit-normal tensor [bs,31,16]; ``h`` = the two
    unit-normal draws of the fading coefficient (z1, z2).  AWGN is the inline form without sqrt(size)
    (:90-93); p is zero so PNR_dB has no effect, as in the reference.  Without ``noise`` every call draws a fresh
    Philox sample (``net.channel_layer`` holds the seed and the call counter)."""
    dev = inp.device
    ns = torch.full((1,), float(n_std), device=dev, dtype=torch.float32)
    sd, off = _draw(net, seed)
    out = engine.greedy_units(net, inp, 1, ns, channel=channel, noise=None if noise is None else noise.contiguous(),
                              seed=sd, offset=off, h=_fading_h(channel, h, dev), max_length=args.max_length,
                              start_idx=args.start_idx, decoder=decoder)
    return out.clone()


def _attacked_greedy(args, inp, net, PNR_dB, channel, n_std, pertutation, noise, h, seed, decoder):
    """Transmit with the FGM perturbation (inline AWGN without sqrt(size), utlis/eval.py:47-55) and greedy-decode."""
    dev = inp.device
    ns = torch.full((1,), float(n_std), device=dev, dtype=torch.float32)
    PNR = 10 ** (PNR_dB / 10)
    ps = torch.full((1,), float(n_std) * math.sqrt(PNR), device=dev, dtype=torch.float32)
    inp32 = inp.to(torch.int32).contiguous()
    sd, off = _draw(net, seed)
    x, y = engine.transmit(net, inp32, 1, ns, channel=channel, noise=None if noise is None else noise.contiguous(),
                           seed=sd, offset=off, h=_fading_h(channel, h, dev), p=pertutation if channel == 'AWGN' else None,
                           p_scale=ps if channel == 'AWGN' else None, want_x_norm=True)
    if decoder is None:
        decoder = engine.make_decoder(net, inp32.shape[0], args.max_length, graph=False)
    if isinstance(decoder, engine.StarGreedyDecoder):
        outputs = decoder.decode(y, args.start_idx)
    else:
        outputs = decoder.decode(y, inp32, args.start_idx)
    scaled = pertutation * (float(n_std) * math.sqrt(PNR))
    realised_noise = (y - x - scaled) if channel == 'AWGN' else None
    return outputs.clone(), scaled, realised_noise, x


def greedy_decode(args, inp, net, PNR_dB, channel='AWGN', n_std=0.1, epsilon=1, *, noise=None, noise2=None, h=None,
                  seed=None, decoder=None):
    """utlis/eval.py:11-75: FGM direction from one teacher-forced pass (gradient w.r.t. the received symbols,
    :25-33), then greedy decoding of the attacked transmission.  Returns (outputs, scaled perturbation,
    channel noise sample, channel_enc_output) like the reference; the third item is the noise realised by the
    attacked pass (the reference draws a fresh, unused sample there)."""
    tar_inp = inp[:, :-1]
    star = isinstance(net.semantic_decoder, (M.SD, M.SDecoder))
    tar_real = inp if star else inp[:, 1:]                          # star decoders emit 31 positions (D11)
    masks = create_masks(inp, tar_inp)
    _, _, g, _ = _symbol_gradient(net, inp, tar_inp, tar_real, PNR_dB, channel, n_std, masks, noise, h, wrt="received",
                                  reforward_awgn=False)
    pertutation = fgm_perturbation(g, epsilon)
    return _attacked_greedy(args, inp, net, PNR_dB, channel, n_std, pertutation, noise2, h, seed, decoder)


def greedy_decode_gan(args, inp, net, PNR_dB, channel='AWGN', n_std=0.1, epsilon=1, *, noise=None, noise2=None,
                      h=None, seed=None, decoder=None):
    """utlis/eval.py:120-187 (``Transeiver_GAN``): as greedy_decode with the clean branch as the attacked loss;
    additionally returns ``noa`` = teacher-forced argmax of the clean branch (:185).
    Returns (outputs, noa, scaled perturbation, noise sample, channel_enc_output)."""
    tar_inp, tar_real = inp[:, :-1], inp[:, 1:]
    masks = create_masks(inp, tar_inp)
    _, pred_r, g, _ = _symbol_gradient(net, inp, tar_inp, tar_real, PNR_dB, channel, n_std, masks, noise, h, gan=True,
                                       wrt="received", reforward_awgn=False)
    noa = _lib.argmax_rows(pred_r)
    pertutation = fgm_perturbation(g, epsilon)
    outputs, scaled, z, x = _attacked_greedy(args, inp, net, PNR_dB, channel, n_std, pertutation, noise2, h, seed, decoder)
    return outputs, noa, scaled, z, x
