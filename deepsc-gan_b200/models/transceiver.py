"""End-to-end systems and the channel: mirror of DeepSC-GAN/models/transceiver.py.

``Channels``, ``Channel_Encoder``, ``Channel_Decoder`` and the four ``Transeiver*`` wirings keep the
reference's names, attributes (``semantic_encoder``, ``channel_encoder``, ``channel_layer``,
``channel_decoder``, ``semantic_decoder``, ``generator``), positional argument order and 4-tuple
returns.  Additions are keyword-only and concern injected randomness: ``noise`` (unit-normal tensor,
shape of the symbols), ``h`` (the two unit-normal draws of the fading coefficient) and ``seed`` for the
on-device Philox stream used when ``noise`` is None.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch
from torch import nn

from .. import _lib
from .. import autograd as AG
from . import modules as _M
from .gan import G
from .modules import Decoder, Dense, Encoder, LayerNormalization, SD, SDecoder, SE, SEncoder, _add_ln

_DETECTORS = {"LS": 1, "MMSE": 2}


class Channels(nn.Module):
    """models/transceiver.py:13-83, one fused kernel per call (dsc_channel).

    Reference behaviour kept by default: the fading coefficient is one complex scalar per call
    (:48-50), ``p`` and ``PNR_dB`` are ignored in fading, and the LS/MMSE estimate is computed but the
    *unequalised* y is returned (:74-75).  ``apply_detector=True`` returns the estimate instead."""

    def __init__(self, apply_detector: bool = False, seed: int = 0):
        super().__init__()
        self.apply_detector = apply_detector
        self.seed = seed
        self._calls = 0

    def _next_offset(self) -> int:
        self._calls += 1
        return self._calls

    def forward(self, inputs, p, PNR_dB, n_std=0.1, channel="AWGN", K=0, detector="MMSE", *, noise=None, h=None):
        if channel == "AWGN":
            return self.awgn(inputs, p, PNR_dB, n_std, noise=noise)
        elif channel == "Rayleigh":
            return self.fading(inputs, p, PNR_dB, 0, n_std, detector, noise=noise, h=h)
        else:
            return self.fading(inputs, p, PNR_dB, 1, n_std, detector, noise=noise, h=h)

    call = forward

    def awgn(self, inputs, p, PNR_dB, n_std=0.1, *, noise=None, scale_by_sqrt_size: bool = True):
        """y = x + N(0, n_std) + n_std*sqrt(PNR)*sqrt(size)*p  (:25-33).  ``scale_by_sqrt_size=False`` is
        the inline form of utlis/eval.py:90-93."""
        x = inputs.contiguous()
        dev = x.device
        PNR = 10 ** (PNR_dB / 10)
        scale = float(n_std) * math.sqrt(PNR) * (math.sqrt(float(x.numel())) if scale_by_sqrt_size else 1.0)
        ns = torch.full((1,), float(n_std), device=dev, dtype=torch.float32)
        ps = torch.full((1,), scale, device=dev, dtype=torch.float32)
        noise = None if noise is None else noise.contiguous()
        if _M.is_differentiable():
            return AG.Channel.apply(x, p, 1, ns, noise, self.seed, self._next_offset(), ps, None, 0)
        y, _ = _lib.channel(x, 1, ns, noise=noise, seed=self.seed,
                            offset=self._next_offset(), p=None if p is None else p.contiguous(), p_scale=ps)
        return y

    def fading(self, inputs, p, PNR_dB, K=0, n_std=0.1, detector="MMSE", *, noise=None, h: Optional[Sequence[float]] = None):
        """y = x*h + n on I/Q pairs, h = N(mean,std) + j N(mean,std) with mean = sqrt(K/(2(K+1))),
        std = sqrt(1/(2(K+1))) (:35-83).  ``h`` = the two unit-normal draws (z1, z2)."""
        if detector not in _DETECTORS:
            raise ValueError("detector must in LS and MMSE")
        x = inputs.contiguous()
        dev = x.device
        mean = math.sqrt(K / (2 * (K + 1)))
        std = math.sqrt(1 / (2 * (K + 1)))
        if h is None:
            z = torch.randn(2).tolist()
        else:
            z = [float(h[0]), float(h[1])]
        hh = torch.tensor([mean + std * z[0], mean + std * z[1]], device=dev, dtype=torch.float32)
        ns = torch.full((1,), float(n_std), device=dev, dtype=torch.float32)
        noise = None if noise is None else noise.contiguous()
        det = _DETECTORS[detector] if self.apply_detector else 0
        if _M.is_differentiable():
            return AG.Channel.apply(x, None, 1, ns, noise, self.seed, self._next_offset(), None, hh, det)
        y, _ = _lib.channel(x, 1, ns, noise=noise, seed=self.seed, offset=self._next_offset(), h=hh, detector=det)
        return y


class Channel_Encoder(nn.Module):
    """models/transceiver.py:85-98: Dense(256, relu) -> Dense(16) -> x / sqrt(mean(x^2)) over the whole
    batch tensor (padding positions included)."""

    def __init__(self, size1=256, size2=16, d_model=128):
        super().__init__()
        self.dense0 = Dense(d_model, size1, activation="relu")
        self.dense1 = Dense(size1, size2)

    def raw(self, inputs):
        return self.dense1(self.dense0(inputs))

    def forward(self, inputs):
        if _M.is_differentiable():
            return AG.PowerNormalize.apply(self.raw(inputs), 1, 1.0)
        return _lib.power_normalize(self.raw(inputs).contiguous(), 1, factor=1.0)

    call = forward


class Channel_Decoder(nn.Module):
    """models/transceiver.py:100-113: x1 = Dense(128, relu), x2 = Dense(512, relu), x3 = Dense(128),
    LayerNorm(x1 + x3)."""

    def __init__(self, size1, size2, in_features=16):
        super().__init__()
        self.dense1 = Dense(in_features, size1, activation="relu")
        self.dense2 = Dense(size1, size2, activation="relu")
        self.dense3 = Dense(size2, size1)
        self.layernorm1 = LayerNormalization(size1)

    def forward(self, receives):
        x1 = self.dense1(receives)
        x3 = self.dense3(self.dense2(x1))
        lead = x1.shape[:-1]
        out = _add_ln(x1.reshape(1, -1, 128), x3.reshape(1, -1, 128), self.layernorm1)
        return out.reshape(*lead, 128)

    call = forward


def _vocab(args) -> int:
    return int(getattr(args, "vocab_size"))


class _TranseiverBase(nn.Module):
    """Shared helpers (not in the reference): TF-name state access for oracle/ckpt interchange."""

    def tf_state_dict(self):
        """Parameters keyed by the TF checkpoint names ('/' separated, relative to the model root)."""
        return {k.replace(".", "/"): v.detach() for k, v in self.named_parameters()}

    def load_tf_state_dict(self, params) -> None:
        own = dict(self.named_parameters())
        missing = set(k.replace(".", "/") for k in own) ^ set(params)
        if missing:
            raise KeyError(f"parameter name mismatch: {sorted(missing)[:8]} ...")
        with torch.no_grad():
            for k, v in own.items():
                v.copy_(params[k.replace(".", "/")].to(v.device, v.dtype))

    def num_parameters(self) -> int:
        return sum(p.numel() for p in self.parameters())

    @staticmethod
    def _tape(symbols: torch.Tensor) -> torch.Tensor:
        """Inside ``differentiable()`` the channel symbols are always a tape variable, so that the evaluators can take
        d(loss)/d(symbols) with frozen parameters (tf.GradientTape watches every intermediate, utlis/eval.py:197-213)."""
        if _M.is_differentiable() and not symbols.requires_grad:
            symbols.requires_grad_(True)
        return symbols


class Transeiver(_TranseiverBase):
    """models/transceiver.py:115-161: DeepSC baseline (4+4 post-LN layers with identity feed-forward)."""

    def __init__(self, args):
        super().__init__()
        self.semantic_encoder = Encoder(args.encoder_num_layer, args.encoder_num_heads, args.encoder_d_model,
                                        args.encoder_d_ff, _vocab(args), dropout_pro=args.encoder_dropout)
        self.semantic_decoder = Decoder(args.decoder_num_layer, args.decoder_d_model, args.decoder_num_heads,
                                        args.decoder_d_ff, _vocab(args), dropout_pro=args.decoder_dropout)
        self.channel_encoder = Channel_Encoder(256, 16)
        self.channel_decoder = Channel_Decoder(args.decoder_d_model, 512)
        self.channel_layer = Channels()

    def forward(self, inputs, tar_inp, p, PNR_dB, channel="AWGN", n_std=0.1, training=False, enc_padding_mask=None,
                combined_mask=None, dec_padding_mask=None, *, noise=None, h=None):
        sema_enc_output = self.semantic_encoder.call(inputs, training, enc_padding_mask)
        channel_enc_output = self._tape(self.channel_encoder.call(sema_enc_output))
        received = self.channel_layer(channel_enc_output, p, PNR_dB, n_std, channel, noise=noise, h=h)
        received_dec = self.channel_decoder.call(received)
        predictions = self.semantic_decoder.call(tar_inp, received_dec, training, combined_mask, dec_padding_mask)
        return predictions, channel_enc_output, received, received

    call = forward


class Transeiver_star(_TranseiverBase):
    """models/transceiver.py:163-206: 4-layer star codec (SEncoder / SDecoder)."""

    def __init__(self, args):
        super().__init__()
        self.semantic_encoder = SEncoder(args.cycle_num, args.encoder_num_layer, args.encoder_num_heads,
                                         args.encoder_d_model, args.encoder_d_ff, _vocab(args),
                                         dropout_pro=args.encoder_dropout)
        self.semantic_decoder = SDecoder(args.cycle_num, args.decoder_num_layer, args.decoder_d_model,
                                         args.decoder_num_heads, args.decoder_d_ff, _vocab(args),
                                         dropout_pro=args.decoder_dropout)
        self.channel_encoder = Channel_Encoder(256, 16)
        self.channel_decoder = Channel_Decoder(args.decoder_d_model, 512)
        self.channel_layer = Channels()

    def forward(self, inputs, tar_inp, p, PNR_dB, channel="AWGN", n_std=0.1, training=False, enc_padding_mask=None,
                combined_mask=None, dec_padding_mask=None, *, noise=None, h=None):
        sema_enc_output = self.semantic_encoder.call(inputs, training, enc_padding_mask)
        channel_enc_output = self._tape(self.channel_encoder.call(sema_enc_output))
        if channel == "AWGN":
            received = self.channel_layer.awgn(channel_enc_output, p, PNR_dB, n_std, noise=noise)
        elif channel == "Rayleigh":
            received = self.channel_layer.fading(channel_enc_output, p, PNR_dB, 0, n_std, noise=noise, h=h)
        else:
            received = self.channel_layer.fading(channel_enc_output, p, PNR_dB, 1, n_std, noise=noise, h=h)
        received_dec = self.channel_decoder.call(received)
        predictions = self.semantic_decoder.call(tar_inp, received_dec, combined_mask, training, dec_padding_mask)
        return predictions, channel_enc_output, received, received

    call = forward


class Transeiver_Star(_TranseiverBase):
    """models/transceiver.py:208-245: single STE / STD codec (the checkpoint at checkpoint/ckpt-9)."""

    def __init__(self, args):
        super().__init__()
        self.semantic_encoder = SE(args.cycle_num, args.cycle_layers, args.encoder_num_heads, args.encoder_d_model,
                                   args.encoder_d_ff, _vocab(args), dropout_pro=args.encoder_dropout)
        self.semantic_decoder = SD(args.cycle_num, args.cycle_layers, args.decoder_d_model, args.decoder_num_heads,
                                   args.decoder_d_ff, _vocab(args), dropout_pro=args.decoder_dropout)
        self.channel_encoder = Channel_Encoder(256, 16)
        self.channel_decoder = Channel_Decoder(args.decoder_d_model, 512)
        self.channel_layer = Channels()

    def forward(self, inputs, tar_inp, p, PNR_dB, channel="AWGN", n_std=0.1, training=False, enc_padding_mask=None,
                combined_mask=None, dec_padding_mask=None, *, noise=None, h=None):
        sema_enc_output = self.semantic_encoder.call(inputs, training, enc_padding_mask)
        channel_enc_output = self._tape(self.channel_encoder.call(sema_enc_output))
        received = self.channel_layer(channel_enc_output, p, PNR_dB, n_std, channel, noise=noise, h=h)
        received_dec = self.channel_decoder.call(received)
        predictions = self.semantic_decoder.call(tar_inp, received_dec, training, combined_mask, dec_padding_mask)
        return predictions, channel_enc_output, received, received

    call = forward


class Transeiver_GAN(_TranseiverBase):
    """models/transceiver.py:247-300: baseline codec + generator ``G``; the channel, channel decoder and
    semantic decoder run twice (perturbed branch, clean branch)."""

    def __init__(self, args):
        super().__init__()
        self.semantic_encoder = Encoder(args.encoder_num_layer, args.encoder_num_heads, args.encoder_d_model,
                                        args.encoder_d_ff, _vocab(args), dropout_pro=args.encoder_dropout)
        self.semantic_decoder = Decoder(args.decoder_num_layer, args.decoder_d_model, args.decoder_num_heads,
                                        args.decoder_d_ff, _vocab(args), dropout_pro=args.decoder_dropout)
        self.generator = G()
        self.channel_encoder = Channel_Encoder(256, 16)
        self.channel_decoder = Channel_Decoder(args.decoder_d_model, 512)
        self.channel_layer = Channels()

    def forward(self, inputs, tar_inp, pertutation, PNR_dB, channel="AWGN", n_std=0.1, training=False,
                enc_padding_mask=None, combined_mask=None, dec_padding_mask=None, traingan=False, *,
                noise=None, h=None, noise_r=None, h_r=None):
        sema_enc_output = self.semantic_encoder.call(inputs, training, enc_padding_mask)
        channel_enc_output = self._tape(self.channel_encoder.call(sema_enc_output))
        p = self.generator.call(channel_enc_output) if traingan else pertutation
        y_p = self.channel_layer(channel_enc_output, p, PNR_dB, n_std, channel, noise=noise, h=h)
        y_r = self.channel_layer(channel_enc_output, None, PNR_dB, n_std, channel, noise=noise_r, h=h_r)   # p = zeros :288
        dec_p = self.channel_decoder.call(y_p)
        dec_r = self.channel_decoder.call(y_r)
        predictions_p = self.semantic_decoder.call(tar_inp, dec_p, training, combined_mask, dec_padding_mask)
        predictions_r = self.semantic_decoder.call(tar_inp, dec_r, training, combined_mask, dec_padding_mask)
        return predictions_p, predictions_r, channel_enc_output, y_r

    call = forward
