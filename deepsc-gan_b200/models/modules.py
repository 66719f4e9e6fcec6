"""Semantic codec building blocks: the PyTorch-side mirror of DeepSC-GAN/models/modules.py.

Same class names, attribute names, argument order and return arity as the reference's Keras
layers, so code written against ``models.modules`` keeps working; every forward pass runs on the
sm_100a kernels of libdeepsc_b200.so (see ``.._lib``), there is no eager/CPU path.  Parameters use
the Keras layouts and the TF checkpoint names (Dense ``kernel`` [in, out] + ``bias``, Embedding
``embeddings``, LayerNormalization ``gamma``/``beta``), so ``state_dict`` keys are the dotted form
of the names in DeepSC-GAN/checkpoint/**/ckpt-9.index.

Reference quirks that are kept on purpose: the feed-forward sublayer is an
identity (modules.py:389-401), star attention is cyclic and unmasked (:289-299), ``STE`` applies
``layernorm1`` twice (:310,:314), the 4-layer star layers drive the relay with the satellite weights
(:175,:243), the star decoders emit 31 (memory-length) positions (:376).

Two execution modes share these classes.  The default (inference) mode reuses workspaces, writes in place
and takes the fused tcgen05 kernels.  Inside ``with differentiable():`` every op goes through the
``torch.autograd.Function`` wrappers of ``..autograd`` (forward and backward both libdeepsc_b200.so kernels,
SURVEY.md K17), out of place, so that ``utlis`` can take d(loss)/d(symbols) and parameter gradients;
``training=True`` additionally applies dropout at the reference's sites (Philox mask kernel).
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import numpy as np
import torch
from torch import nn

from .. import _lib
from .. import autograd as AG

D_MODEL = 128
# Arithmetic of every Dense / star-cycle kernel (the `prec` argument of include/deepsc_b200.h):
#   1 (default)  tcgen05 bf16x3: three bf16 UMMA passes per fp32-class product, fp32 accumulation in TMEM
#   2            tcgen05 bf16, one pass (throughput experiments; not parity-grade)
#   0            fp32 FFMA kernels - a debugging aid to separate arithmetic from logic errors, selected explicitly with
#                set_precision(0); it is also what the backward tape uses for the products with K < 128
PREC = 1


def set_precision(prec: int) -> None:
    """Select the arithmetic of every module (see PREC above and include/deepsc_b200.h dsc_linear)."""
    global PREC
    assert prec in (0, 1, 2)
    PREC = prec


class precision:
    """``with precision(0):`` - temporary arithmetic override.  The evaluators take d(loss)/d(symbols) under it in fp32:
    the FGM normalisation divides every sentence's gradient by its own norm (utlis/eval.py:36-44), which turns the ~1e-5
    relative error of a bf16x3 forward into percents of the perturbation for the sentences whose gradient is small."""

    def __init__(self, prec: int):
        assert prec in (0, 1, 2)
        self.prec = prec

    def __enter__(self):
        global PREC
        self._prev, PREC = PREC, self.prec
        return self

    def __exit__(self, *exc):
        global PREC
        PREC = self._prev
        return False


_DIFF = False
_DROPOUT = {"seed": 0x5EED, "calls": 0}


class differentiable:
    """Context manager: route every op through the autograd Functions (out of place, unfused star cycles)."""

    def __enter__(self):
        global _DIFF
        self._prev, _DIFF = _DIFF, True
        return self

    def __exit__(self, *exc):
        global _DIFF
        _DIFF = self._prev
        return False


def is_differentiable() -> bool:
    return _DIFF


def set_dropout_seed(seed: int) -> None:
    _DROPOUT["seed"], _DROPOUT["calls"] = int(seed), 0


def _dropout(x: torch.Tensor, rate: float, training) -> torch.Tensor:
    """tf.keras.layers.Dropout(rate)(x, training=training)."""
    if not training or rate <= 0:
        return x
    _DROPOUT["calls"] += 1
    return AG.Dropout.apply(x, float(rate), _DROPOUT["seed"], _DROPOUT["calls"])


def postional_encoder(position: int, d_model: int) -> torch.Tensor:
    """models/modules.py:5-23 (name kept, typo included).  Returns [1, position, d_model] float32."""
    pos = np.arange(position)[:, None]
    i = np.arange(d_model)[None, :]
    angle = pos / np.power(10000, (2 * i) / np.float32(d_model))
    angle[:, 0::2] = np.sin(angle[:, 0::2])
    angle[:, 1::2] = np.cos(angle[:, 1::2])
    return torch.from_numpy(angle[None, ...].astype(np.float32))


# --------------------------------------------------------------------------- Keras-shaped leaves
class Dense(nn.Module):
    """tf.keras.layers.Dense: y = act(x @ kernel + bias), kernel [in, out], glorot_uniform / zeros."""

    def __init__(self, in_features: int, units: int, activation: Optional[str] = None, use_bias: bool = True):
        super().__init__()
        lim = math.sqrt(6.0 / (in_features + units))
        self.kernel = nn.Parameter((torch.rand(in_features, units) * 2 - 1) * lim)
        self.bias = nn.Parameter(torch.zeros(units)) if use_bias else None
        self.act = {None: 0, "relu": 1}[activation]
        self._packed = None

    def padded_kernel(self) -> torch.Tensor:
        """Kernel with a row stride that is a multiple of 4 floats (what dsc_linear requires)."""
        k = self.kernel
        if k.shape[1] % 4 == 0:
            return k.detach()
        key = (k._version, _lib.WEIGHT_EPOCH, k.data_ptr(), k.device)
        if self._packed is None or self._packed[0] != key:
            n_pad = (k.shape[1] + 127) // 128 * 128
            buf = torch.zeros((k.shape[0], n_pad), device=k.device, dtype=torch.float32)
            buf[:, : k.shape[1]] = k.detach()
            self._packed = (key, buf)
        return self._packed[1]

    def forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if _DIFF:
            return AG.linear(x, self.kernel, self.bias, self.act, PREC, wfwd=self.padded_kernel())
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        y = _lib.linear(x2, self.padded_kernel(), None if self.bias is None else self.bias.detach(), self.act,
                        out=out, n=self.kernel.shape[1], prec=PREC if x2.shape[1] >= 128 else 0)
        return y.reshape(*lead, self.kernel.shape[1]) if out is None else out


class Embedding(nn.Module):
    """tf.keras.layers.Embedding with its default uniform(-0.05, 0.05) initialiser."""

    def __init__(self, vocab_size: int, d_model: int):
        super().__init__()
        self.embeddings = nn.Parameter((torch.rand(vocab_size, d_model) - 0.5) * 0.1)


class LayerNormalization(nn.Module):
    """tf.keras.layers.LayerNormalization(epsilon=1e-6): gamma ones, beta zeros."""

    def __init__(self, d_model: int = D_MODEL):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(d_model))
        self.beta = nn.Parameter(torch.zeros(d_model))


def _as_ids(x: torch.Tensor) -> torch.Tensor:
    """int32 ids with unit stride along the sentence; the row stride is free (dsc_embed takes it), so the one-column slice
    a greedy step embeds is passed as a view instead of being copied by a separate kernel every step."""
    if x.dtype != torch.int32:
        x = x.to(torch.int32)
    if x.dim() == 2 and (x.shape[1] == 1 or x.stride(1) == 1):
        return x
    return x.contiguous()


def _check_eval(training, rate: float) -> None:
    if training and rate > 0 and not _DIFF:
        raise RuntimeError("training=True with dropout runs in the differentiable mode only: wrap the call in "
                           "`with models.modules.differentiable():` (utlis.trainer / utlis.gan_train do)")


def _add_ln(x: torch.Tensor, res: Optional[torch.Tensor], ln_a: LayerNormalization,
            ln_b: Optional[LayerNormalization] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    if _DIFF:
        return AG.add_layernorm(x, res, ln_a, ln_b)
    return _lib.add_layernorm(x, res, ln_a.gamma.detach(), ln_a.beta.detach(),
                              None if ln_b is None else ln_b.gamma.detach(),
                              None if ln_b is None else ln_b.beta.detach(), out=out)


def _res_ln(attn: torch.Tensor, res: torch.Tensor, ln: LayerNormalization, training, rate: float) -> torch.Tensor:
    """LN(res + dropout(attn)): the attention sub-block of every layer (e.g. models/modules.py:424-425)."""
    return _add_ln(_dropout(attn, rate, training), res, ln)


def _res_ln2(attn: torch.Tensor, res: torch.Tensor, ln_a: LayerNormalization, ln_b: LayerNormalization, training,
             rate: float) -> torch.Tensor:
    """o1 = LN_a(res + dropout(attn)); LN_b(o1 + dropout(ffn(o1))) with the identity feed-forward (:424-429 and the
    same pattern in every layer).  Without dropout this is the fused LN_b(2 * LN_a(.)) kernel."""
    if training and rate > 0:
        o1 = _add_ln(_dropout(attn, rate, training), res, ln_a)
        return _add_ln(_dropout(o1, rate, training), o1, ln_b)
    return _add_ln(attn, res, ln_a, ln_b)


# --------------------------------------------------------------------------- multi-head attention
class sublayer1(nn.Module):
    """Multi-head attention, models/modules.py:26-123: wq/wk/wv without bias, dense with bias."""

    def __init__(self, d_model: int, num_heads: int):
        super().__init__()
        assert d_model == D_MODEL and num_heads == 8, "kernels are written for d_model=128, 8 heads"
        self.d_model, self.num_heads, self.depth = d_model, num_heads, d_model // num_heads
        self.wq = Dense(d_model, d_model, use_bias=False)
        self.wk = Dense(d_model, d_model, use_bias=False)
        self.wv = Dense(d_model, d_model, use_bias=False)
        self.dense = Dense(d_model, d_model)
        self._cache = {}

    def _packed(self, which: str) -> torch.Tensor:
        """Concatenated projection weights: 'qkv' [128,384], 'kv' [128,256]."""
        ws = {"qkv": (self.wq, self.wk, self.wv), "kv": (self.wk, self.wv),
              "qkv_grouped": (self.wq, self.wk, self.wv)}[which]
        key = (_lib.WEIGHT_EPOCH,) + tuple((w.kernel._version, w.kernel.data_ptr()) for w in ws)
        hit = self._cache.get(which)
        if hit is None or hit[0] != key:
            if which == "qkv_grouped":      # head pairs g: [wq[:,32g:32g+32] | wk[...] | wv[...]] (dsc_star_sat_tc)
                cols = [w.kernel.detach()[:, 32 * g:32 * g + 32] for g in range(4) for w in ws]
                packed = torch.cat(cols, dim=1).contiguous()
            else:
                packed = torch.cat([w.kernel.detach() for w in ws], dim=1).contiguous()
            hit = (key, packed)
            self._cache[which] = hit
        return hit[1]

    def project(self, x2: torch.Tensor, which: str) -> torch.Tensor:
        """x2 [rows,128] @ concatenated projection weights ('qkv' -> [rows,384], 'kv' -> [rows,256]), in either mode.
        In the differentiable mode the weight operand is the (differentiable) concatenation of the kernels, whose
        backward is the slicing; the forward kernel still reads the cached packed copy."""
        if _DIFF:
            ws = (self.wq, self.wk, self.wv) if which == "qkv" else (self.wk, self.wv)
            return AG.linear(x2, torch.cat([w.kernel for w in ws], dim=1), None, 0, PREC, wfwd=self._packed(which))
        return _lib.linear(x2, self._packed(which), None, prec=PREC)

    def attend(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mask=None, key_ids=None,
               causal: bool = False, q_off: int = 0) -> torch.Tensor:
        """Projected q [n,lq,128] / k,v [n,lk,128] views -> dense(softmax(qk^T/4 + mask*-1e9) v)."""
        n, lq, _ = q.shape
        if _DIFF:
            return self.dense(AG.MhaAttention.apply(q, k, v, mask, key_ids, causal, q_off))
        o = torch.empty((n, lq, D_MODEL), device=q.device, dtype=torch.float32)
        _lib.mha_attention(q, k, v, o, mask=mask, key_ids=key_ids, causal=causal, q_off=q_off)
        return self.dense(o)

    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mask) -> torch.Tensor:
        n, lq, _ = q.shape
        lk = k.shape[1]
        if q is k and k is v:
            qkv = self.project(q.reshape(-1, D_MODEL), "qkv").view(n, lq, 384)
            Q, K, V = qkv[..., 0:128], qkv[..., 128:256], qkv[..., 256:384]
        else:
            Q = self.wq(q)
            if k is v:
                kv = self.project(k.reshape(-1, D_MODEL), "kv").view(n, lk, 256)
                K, V = kv[..., 0:128], kv[..., 128:256]
            else:
                K, V = self.wk(k), self.wv(v)
        return self.attend(Q, K, V, mask)

    call = forward


class sublayer2(nn.Module):
    """The reference's feed-forward sublayer defines no ``call`` (models/modules.py:389-401), so under
    tf.keras it is the identity and owns no variables.  Kept as such."""

    def __init__(self, d_model: int, dff: int):
        super().__init__()
        self.d_model, self.dff = d_model, dff

    def forward(self, x):
        return x

    call = forward


# --------------------------------------------------------------------------- star cycles engine
class StarWorkspace:
    """Per-batch scratch for the cycle loop, reused across cycles / greedy steps."""

    def __init__(self, n_sent: int, device):
        self.n = n_sent
        f = dict(device=device, dtype=torch.float32)
        self.x = torch.empty((n_sent, 32, 128), **f)
        self.qkv = torch.empty((n_sent * 32, 384), **f)
        self.kv_e = torch.empty((n_sent * 32, 256), **f)
        self.att = torch.empty((n_sent * 32, 128), **f)
        self.att_r = torch.empty((n_sent, 128), **f)
        self.q_r = torch.empty((n_sent, 128), **f)
        self.kvei = None                                    # interleaved k|v of the e rows (tcgen05 path)
        self.kv2i = None                                    # interleaved h2 cache built from a row-major kv2
        self.xi0 = None                                     # interleaved e tile, s0 and q0 = s0 @ wq_relay: the
        self.s0 = self.q0 = None                            # cycle-0 state of the one-launch kernel (constant per e tile)
        self.xi1 = self.xi1_buf = None                      # interleaved X' of cycle 0 (satellite half precomputed, greedy decode)

    def tc_buffers(self):
        if self.kvei is None:
            f = dict(device=self.x.device, dtype=torch.float32)
            self.kvei = torch.empty((self.n * 32 * 256,), **f)
        return self.kvei


def use_tc(n_sent: int = 4) -> bool:
    """True when the star cycles run on the fused tcgen05 kernel (every precision mode but the fp32 debug mode)."""
    return PREC != 0


def _pad4(t: torch.Tensor) -> torch.Tensor:
    """Zero-pad the sentence axis to a multiple of 4 (one tile of the tcgen05 star kernel = 4 sentences).  A ragged last
    batch (dataset/dataloader.py:14 batches without drop_remainder) takes this path; whole 64-sentence units never do."""
    S = t.shape[0]
    if S % 4 == 0:
        return t
    out = torch.zeros((S + 3) // 4 * 4, *t.shape[1:], device=t.device, dtype=t.dtype)
    out[:S] = t
    return out


def prepare_kv_e(e_tile: torch.Tensor, sat: sublayer1, ws: StarWorkspace, relay: Optional[sublayer1] = None,
                 first_sat: bool = False) -> None:
    """Everything that depends on the e tile only (constant over cycles and greedy steps): k|v of the e rows under the
    satellite weights and, for the one-launch tcgen05 kernel, the interleaved e tile, s0 and q0 = s0 @ wq_relay.
    ``first_sat``: also the satellite half of the FIRST cycle (models/modules.py:287-300 reads h = e, s = mean(e) and
    the e keys only), so that a greedy decoder runs it once per batch instead of once per step.
    e_tile [S,32,128] with S = ws.n (a multiple of 4 in the tcgen05 modes)."""
    S = e_tile.shape[0]
    _lib.linear(e_tile.view(S * 32, 128), sat._packed("kv"), None, out=ws.kv_e, prec=PREC)
    if use_tc() and relay is not None:
        _lib.star_interleave(ws.kv_e.view(S // 4, 128, 256), ws.tc_buffers(), 128)
        f = dict(device=e_tile.device, dtype=torch.float32)
        if ws.xi0 is None:
            ws.xi0, ws.s0, ws.q0 = torch.empty((S * 4096,), **f), torch.empty((S, 128), **f), torch.empty((S, 128), **f)
        _lib.star_interleave(e_tile.view(S // 4, 128, 128), ws.xi0, 128)
        ws.s0.copy_(e_tile[:, 31, :])
        _lib.linear(ws.s0, relay.wq.kernel.detach(), None, out=ws.q0, prec=PREC)
        ws.xi1 = None
        if first_sat:
            # one cycle without target keys: rows 0..30 of the result are X' = relu(ATT @ Wo + b) of cycle 0
            _lib.star_cycles_tc(ws.xi0, ws.s0, ws.q0, ws.kvei, None, 0, sat._packed("qkv_grouped"),
                                sat.dense.kernel.detach(), relay._packed("kv"), relay.dense.kernel.detach(),
                                relay.wq.kernel.detach(), sat.dense.bias.detach(), relay.dense.bias.detach(),
                                ws.x, S, 1, PREC)
            if ws.xi1_buf is None:
                ws.xi1_buf = torch.empty((S * 4096,), **f)
            ws.xi1 = _lib.star_interleave(ws.x.view(S // 4, 128, 128), ws.xi1_buf, 128)


def star_cycles(e_tile: torch.Tensor, sat: sublayer1, relay: sublayer1, cycle_num: int,
                kv2: Optional[torch.Tensor] = None, n2: int = 0, ws: Optional[StarWorkspace] = None,
                kv_e_ready: bool = False, kv2i: Optional[torch.Tensor] = None, relay_row: bool = True) -> torch.Tensor:
    """The satellite/relay cycle loop of STE/STD/StarTransformer*Layer (models/modules.py:283-306,
    359-378) on star tiles, deduplicated: every node is projected once per cycle, neighbours are
    gathered by index.  e_tile [S,32,128] with row 31 = mean over tokens; returns the tile after
    ``cycle_num`` cycles (rows 0..30 = h, row 31 = s).  kv2 [S, rows, 256] holds k|v of h2 under the
    relay weights (decoder only), of which the first n2 rows are attended; ``kv2i`` is the same cache
    already in the interleaved layout of the tcgen05 kernel.  ``relay_row=False``: the caller reads rows 0..30 only
    (a greedy decoder), so the kernel skips the relay half of the last cycle (row 31 is then the relay node before its
    last update).

    One launch of dsc_star_cycles_tc runs all cycles (PREC 1 / 2).  PREC 0 (fp32 debug mode) runs the per-op fp32 kernels
    that the backward tape differentiates."""
    S = e_tile.shape[0]
    if _DIFF:
        return _star_cycles_diff(e_tile, sat, relay, cycle_num, kv2, n2)
    if use_tc():
        if S % 4:                                       # ragged batch: pad to whole tiles, drop the padding again
            assert ws is None and not kv_e_ready and kv2i is None
            x = star_cycles(_pad4(e_tile), sat, relay, cycle_num, None if kv2 is None else _pad4(kv2), n2,
                            relay_row=relay_row)
            return x[:S]
        if ws is None:
            ws = StarWorkspace(S, e_tile.device)
        if not kv_e_ready:
            prepare_kv_e(e_tile, sat, ws, relay)
        if n2 > 0 and kv2i is None:
            pad = torch.zeros((S, 32, 256), device=e_tile.device, dtype=torch.float32)
            pad[:, : kv2.shape[1]] = kv2
            kv2i = _lib.star_interleave(pad, torch.empty_like(pad).view(-1), 32)
        skip = kv_e_ready and ws.xi1 is not None and cycle_num >= 2
        return _lib.star_cycles_tc(ws.xi1 if skip else ws.xi0, ws.s0, ws.q0, ws.kvei, kv2i, n2, sat._packed("qkv_grouped"),
                                   sat.dense.kernel.detach(), relay._packed("kv"), relay.dense.kernel.detach(),
                                   relay.wq.kernel.detach(), sat.dense.bias.detach(), relay.dense.bias.detach(),
                                   ws.x, S, cycle_num, PREC | (_lib.STAR_FIRST_SAT_DONE if skip else 0)
                                   | (0 if relay_row else _lib.STAR_NO_FINAL_RELAY))
    if ws is None:
        ws = StarWorkspace(S, e_tile.device)
    if not kv_e_ready:
        prepare_kv_e(e_tile, sat, ws, relay)
    x2 = ws.x.view(S * 32, 128)
    ws.x.copy_(e_tile)
    s_rows = ws.x[:, 31, :]
    w_qkv_s, w_qkv_r = sat._packed("qkv"), relay._packed("qkv")
    for _ in range(cycle_num):
        _lib.linear(x2, w_qkv_s, None, out=ws.qkv, prec=PREC)
        _lib.star_satellite_attn(ws.qkv, ws.kv_e, ws.att, S)
        _lib.linear(ws.att, sat.dense.kernel.detach(), sat.dense.bias.detach(), act=1, out=x2,
                    row_mod=32, row_skip=31, prec=PREC)
        _lib.linear(x2, w_qkv_r, None, out=ws.qkv, prec=PREC)
        _lib.star_relay_attn(ws.qkv, kv2, n2, ws.att_r, S)
        _lib.linear(ws.att_r, relay.dense.kernel.detach(), relay.dense.bias.detach(), act=1, out=s_rows, prec=PREC)
    return ws.x


def _star_cycles_diff(e_tile: torch.Tensor, sat: "sublayer1", relay: "sublayer1", cycle_num: int,
                      kv2: Optional[torch.Tensor], n2: int) -> torch.Tensor:
    """The same cycle loop, out of place on the autograd Functions (one kernel per op, fp32 or tcgen05 Dense)."""
    S = e_tile.shape[0]
    x = e_tile
    kv_e = sat.project(e_tile.reshape(S * 32, 128), "kv")
    for _ in range(cycle_num):
        qkv = sat.project(x.reshape(S * 32, 128), "qkv")
        att = AG.StarSatelliteAttn.apply(qkv, kv_e)
        h = AG.linear(att, sat.dense.kernel, sat.dense.bias, 1, PREC, wfwd=sat.dense.kernel.detach()).view(S, 32, 128)
        x = torch.cat([h[:, :31], x[:, 31:32]], dim=1)                   # the relay row keeps s
        qkv_r = relay.project(x.reshape(S * 32, 128), "qkv")
        att_r = AG.StarRelayAttn.apply(qkv_r, kv2, n2)
        s = AG.linear(att_r, relay.dense.kernel, relay.dense.bias, 1, PREC, wfwd=relay.dense.kernel.detach())
        x = torch.cat([x[:, :31], s[:, None, :]], dim=1)
    return x


def _star_pack(e: torch.Tensor) -> torch.Tensor:
    return AG.StarPack.apply(e) if _DIFF else _lib.star_pack(e.contiguous())


def _kv_of(x2: torch.Tensor, att: "sublayer1") -> torch.Tensor:
    """k|v projection [rows, 256] of rows x2 under the weights of ``att``."""
    return att.project(x2, "kv")


def _target_branch(layer, tar: torch.Tensor, look_ahead_mask, training=False) -> torch.Tensor:
    """h2 = LN1(tar + dropout1(multi_tar(tar,tar,tar,mask))) (modules.py:219-221, 352-354)."""
    return _res_ln(layer.multi_tar(tar, tar, tar, look_ahead_mask), tar, layer.layernorm1, training, layer.drop_pro)


class _StarBase(nn.Module):
    def cycle_shift(self, x, forward=True):
        return torch.roll(x, shifts=1 if forward else -1, dims=1)


class StarTransformerEncoderLayer(_StarBase):
    """models/modules.py:126-186.  Relay update reuses multi_att_satellite (:175); multi_att_relay is
    constructed by the reference but never called, so it owns no variables and is not created here."""

    def __init__(self, cycle_num, d_model, num_heads, input_vocab_size, dff, drop_pro=0.1):
        super().__init__()
        self.d_model, self.cycle_num, self.depth, self.drop_pro = d_model, cycle_num, d_model // num_heads, drop_pro
        self.multi_att_satellite = sublayer1(d_model, num_heads)
        self.sl2 = sublayer2(d_model, dff)
        self.layernorm1 = LayerNormalization(d_model)
        self.layernorm2 = LayerNormalization(d_model)

    def forward(self, e, training, forward=True, mask=None):
        _check_eval(training, self.drop_pro)
        tile = _star_pack(e)
        x = star_cycles(tile, self.multi_att_satellite, self.multi_att_satellite, self.cycle_num)
        out = _res_ln2(x[:, :31], tile[:, :31], self.layernorm1, self.layernorm2, training, self.drop_pro)
        return out, x[:, 31, :].clone()

    call = forward


class StarTransformerDecoderLayer(_StarBase):
    """models/modules.py:188-253.  layernorm1 serves both h2 (:221) and e+h (:247)."""

    def __init__(self, cycle_num, d_model, num_heads, input_vocab_size, dff, drop_pro=0.1):
        super().__init__()
        self.d_model, self.cycle_num, self.depth, self.drop_pro = d_model, cycle_num, d_model // num_heads, drop_pro
        self.multi_tar = sublayer1(d_model, num_heads)
        self.multi_att_satellite = sublayer1(d_model, num_heads)
        self.sl2 = sublayer2(d_model, dff)
        self.layernorm1 = LayerNormalization(d_model)
        self.layernorm2 = LayerNormalization(d_model)

    def forward(self, tar, e, look_ahead_mask, training, forward=True, mask=None):
        _check_eval(training, self.drop_pro)
        h2 = _target_branch(self, tar, look_ahead_mask, training)
        b, lt, _ = h2.shape
        kv2 = _kv_of(h2.reshape(b * lt, 128), self.multi_att_satellite).view(b, lt, 256)
        tile = _star_pack(e)
        x = star_cycles(tile, self.multi_att_satellite, self.multi_att_satellite, self.cycle_num, kv2, lt)
        out = _res_ln2(x[:, :31], tile[:, :31], self.layernorm1, self.layernorm2, training, self.drop_pro)
        return out, x[:, 31, :].clone()

    call = forward


class STE(_StarBase):
    """models/modules.py:256-320 (encoder of ``Transeiver_Star``): separate relay weights, layernorm1
    applied to e+h and again to 2*output1; layernorm2/embedding/sl2 exist in the reference but own no
    variables (checkpoint/ckpt-9.index), so they are not parameters here."""

    def __init__(self, cycle_num, d_model, num_heads, input_vocab_size, dff, drop_pro=0.1):
        super().__init__()
        self.d_model, self.cycle_num, self.depth, self.drop_pro = d_model, cycle_num, d_model // num_heads, drop_pro
        self.multi_att_satellite = sublayer1(d_model, num_heads)
        self.multi_att_relay = sublayer1(d_model, num_heads)
        self.layernorm1 = LayerNormalization(d_model)
        self.sl2 = sublayer2(d_model, dff)

    def forward(self, e, training, forward=True, mask=None):
        _check_eval(training, self.drop_pro)
        tile = _star_pack(e)
        x = star_cycles(tile, self.multi_att_satellite, self.multi_att_relay, self.cycle_num)
        out = _res_ln2(x[:, :31], tile[:, :31], self.layernorm1, self.layernorm1, training, self.drop_pro)
        return out, x[:, 31, :].clone()

    call = forward


class STD(_StarBase):
    """models/modules.py:322-387 (decoder of ``Transeiver_Star``)."""

    def __init__(self, cycle_num, d_model, num_heads, input_vocab_size, dff, drop_pro=0.1):
        super().__init__()
        self.d_model, self.cycle_num, self.depth, self.drop_pro = d_model, cycle_num, d_model // num_heads, drop_pro
        self.multi_tar = sublayer1(d_model, num_heads)
        self.multi_att_satellite = sublayer1(d_model, num_heads)
        self.multi_att_relay = sublayer1(d_model, num_heads)
        self.layernorm1 = LayerNormalization(d_model)
        self.layernorm2 = LayerNormalization(d_model)
        self.layernorm3 = LayerNormalization(d_model)
        self.sl2 = sublayer2(d_model, dff)

    def forward(self, tar, e, look_ahead_mask, training, forward=True, mask=None):
        _check_eval(training, self.drop_pro)
        h2 = _target_branch(self, tar, look_ahead_mask, training)
        b, lt, _ = h2.shape
        kv2 = _kv_of(h2.reshape(b * lt, 128), self.multi_att_relay).view(b, lt, 256)
        tile = _star_pack(e)
        x = star_cycles(tile, self.multi_att_satellite, self.multi_att_relay, self.cycle_num, kv2, lt)
        out = _res_ln2(x[:, :31], tile[:, :31], self.layernorm2, self.layernorm3, training, self.drop_pro)
        return out, x[:, 31, :].clone()

    call = forward


# --------------------------------------------------------------------------- baseline transformer layers
class EncoderLayer(nn.Module):
    """models/modules.py:405-431: LN1(x + MHA(x)), LN2(2*output1) (identity feed-forward)."""

    def __init__(self, d_model, num_heads, dff, drop_pro=0.1):
        super().__init__()
        self.drop_pro = drop_pro
        self.sl1 = sublayer1(d_model, num_heads)
        self.sl2 = sublayer2(d_model, dff)
        self.layernorm1 = LayerNormalization(d_model)
        self.layernorm2 = LayerNormalization(d_model)

    def forward(self, x, training, mask):
        _check_eval(training, self.drop_pro)
        return _res_ln2(self.sl1(x, x, x, mask), x, self.layernorm1, self.layernorm2, training, self.drop_pro)

    call = forward


class DecoderLayer(nn.Module):
    """models/modules.py:433-469."""

    def __init__(self, d_model, num_heads, dff, drop_pro=0.1):
        super().__init__()
        self.drop_pro = drop_pro
        self.sl11 = sublayer1(d_model, num_heads)
        self.sl12 = sublayer1(d_model, num_heads)
        self.ffn = sublayer2(d_model, dff)
        self.layernorm1 = LayerNormalization(d_model)
        self.layernorm2 = LayerNormalization(d_model)
        self.layernorm3 = LayerNormalization(d_model)

    def forward(self, x, enc_output, training, look_ahead_mask, padding_mask):
        _check_eval(training, self.drop_pro)
        o1 = _res_ln(self.sl11(x, x, x, look_ahead_mask), x, self.layernorm1, training, self.drop_pro)
        return _res_ln2(self.sl12(o1, enc_output, enc_output, padding_mask), o1, self.layernorm2, self.layernorm3,
                        training, self.drop_pro)

    call = forward


# --------------------------------------------------------------------------- encoders / decoders
class _Codec(nn.Module):
    def _embed(self, ids: torch.Tensor, pos0: int = 0, training=False) -> torch.Tensor:
        """Embedding * sqrt(d_model) + positional rows, then the codec's input Dropout (e.g. :497-505)."""
        if self.pos_encoding.device != self.embedding.embeddings.device:
            self.pos_encoding = self.pos_encoding.to(self.embedding.embeddings.device)
        if _DIFF:
            x = AG.Embed.apply(_as_ids(ids), self.embedding.embeddings, self.pos_encoding[0], pos0)
            return _dropout(x, self.dropout_pro, training)
        return _lib.embed(_as_ids(ids), self.embedding.embeddings.detach(), self.pos_encoding[0], pos0)


class Encoder(_Codec):
    """models/modules.py:471-511."""

    def __init__(self, num_layers, num_heads, d_model, dff, input_vocab_size,
                 maximum_position_encoding=512, dropout_pro=0.1):
        super().__init__()
        self.d_model, self.dff, self.num_layers, self.target_vocab_size = d_model, dff, num_layers, input_vocab_size
        self.dropout_pro = dropout_pro
        self.embedding = Embedding(input_vocab_size, d_model)
        self.register_buffer("pos_encoding", postional_encoder(maximum_position_encoding, d_model), persistent=False)
        self.encoder = nn.ModuleList([EncoderLayer(d_model, num_heads, dff, dropout_pro) for _ in range(num_layers)])

    def forward(self, x, training, mask):
        _check_eval(training, self.dropout_pro)
        x = self._embed(x, training=training)
        for layer in self.encoder:
            x = layer(x, training, mask)
        return x

    call = forward


class Decoder(_Codec):
    """models/modules.py:513-552."""

    def __init__(self, num_layers, d_model, num_heads, dff, target_vocab_size,
                 maximum_position_encoding=512, dropout_pro=0.1):
        super().__init__()
        self.d_model, self.num_layers, self.dropout_pro = d_model, num_layers, dropout_pro
        self.embedding = Embedding(target_vocab_size, d_model)
        self.register_buffer("pos_encoding", postional_encoder(maximum_position_encoding, d_model), persistent=False)
        self.dec_layers = nn.ModuleList([DecoderLayer(d_model, num_heads, dff, dropout_pro) for _ in range(num_layers)])
        self.final_layer = Dense(d_model, target_vocab_size)

    def hidden(self, x, enc_output, training, look_ahead_mask, padding_mask):
        x = self._embed(x, training=training)
        for layer in self.dec_layers:
            x = layer(x, enc_output, training, look_ahead_mask, padding_mask)
        return x

    def forward(self, x, enc_output, training, look_ahead_mask, padding_mask):
        _check_eval(training, self.dropout_pro)
        return self.final_layer(self.hidden(x, enc_output, training, look_ahead_mask, padding_mask))

    call = forward


class SEncoder(_Codec):
    """models/modules.py:554-590: 4 stacked StarTransformerEncoderLayer."""

    def __init__(self, cycle_num, num_layers, num_heads, d_model, dff, input_vocab_size,
                 maximum_position_encoding=512, dropout_pro=0.1):
        super().__init__()
        self.d_model, self.dff, self.num_layers, self.target_vocab_size = d_model, dff, num_layers, input_vocab_size
        self.dropout_pro = dropout_pro
        self.embedding = Embedding(input_vocab_size, d_model)
        self.register_buffer("pos_encoding", postional_encoder(maximum_position_encoding, d_model), persistent=False)
        self.encoder = nn.ModuleList([StarTransformerEncoderLayer(cycle_num, d_model, num_heads, input_vocab_size, dff)
                                      for _ in range(num_layers)])

    def forward(self, x, training, mask):
        _check_eval(training, self.dropout_pro)
        x = self._embed(x, training=training)
        for layer in self.encoder:
            x, _ = layer(x, training, True, mask)
        return x

    call = forward


class SDecoder(_Codec):
    """models/modules.py:592-633.  NOTE the argument order (tar, x, look_ahead_mask, training, mask)."""

    def __init__(self, cycle_num, num_layers, d_model, num_heads, dff, target_vocab_size,
                 maximum_position_encoding=512, dropout_pro=0.1):
        super().__init__()
        self.d_model, self.num_layers, self.dropout_pro = d_model, num_layers, dropout_pro
        self.embedding = Embedding(target_vocab_size, d_model)
        self.register_buffer("pos_encoding", postional_encoder(maximum_position_encoding, d_model), persistent=False)
        self.dec_layers = nn.ModuleList([StarTransformerDecoderLayer(cycle_num, d_model, num_heads, target_vocab_size, dff)
                                         for _ in range(num_layers)])
        self.final_layer = Dense(d_model, target_vocab_size)

    def hidden(self, tar, x, look_ahead_mask, training, mask):
        tar = self._embed(tar, training=training)
        for layer in self.dec_layers:
            x, _ = layer(tar, x, look_ahead_mask, training, True, mask)
        return x

    def forward(self, tar, x, look_ahead_mask, training, mask):
        _check_eval(training, self.dropout_pro)
        return self.final_layer(self.hidden(tar, x, look_ahead_mask, training, mask))

    call = forward


class SE(_Codec):
    """models/modules.py:635-674: one STE whatever ``num_layers`` says (:653)."""

    def __init__(self, cycle_num, num_layers, num_heads, d_model, dff, input_vocab_size,
                 maximum_position_encoding=512, dropout_pro=0.1):
        super().__init__()
        self.d_model, self.dff, self.num_layers, self.target_vocab_size = d_model, dff, num_layers, input_vocab_size
        self.dropout_pro = dropout_pro
        self.embedding = Embedding(input_vocab_size, d_model)
        self.register_buffer("pos_encoding", postional_encoder(maximum_position_encoding, d_model), persistent=False)
        self.encoder = STE(cycle_num, d_model, num_heads, input_vocab_size, dff)

    def forward(self, x, training, mask):
        _check_eval(training, self.dropout_pro)
        x, _ = self.encoder(self._embed(x, training=training), training, True, mask)
        return x

    call = forward


class SD(_Codec):
    """models/modules.py:677-718: one STD, argument order (tar, x, training, look_ahead_mask, mask)."""

    def __init__(self, cycle_num, num_layers, d_model, num_heads, dff, target_vocab_size,
                 maximum_position_encoding=512, dropout_pro=0.1):
        super().__init__()
        self.d_model, self.num_layers, self.dropout_pro = d_model, num_layers, dropout_pro
        self.embedding = Embedding(target_vocab_size, d_model)
        self.register_buffer("pos_encoding", postional_encoder(maximum_position_encoding, d_model), persistent=False)
        self.dec_layers = STD(cycle_num, d_model, num_heads, target_vocab_size, dff)
        self.final_layer = Dense(d_model, target_vocab_size)

    def hidden(self, tar, x, training, look_ahead_mask, mask):
        x, _ = self.dec_layers(self._embed(tar, training=training), x, look_ahead_mask, training, True, mask)
        return x

    def forward(self, tar, x, training, look_ahead_mask, mask):
        _check_eval(training, self.dropout_pro)
        return self.final_layer(self.hidden(tar, x, training, look_ahead_mask, mask))

    call = forward


# --------------------------------------------------------------------------- schedule / loss / masks
class CustomSchedule:
    """models/modules.py:719-736: lr = d_model^-0.5 * min(step^-0.5, step * warmup^-1.5)."""

    def __init__(self, d_model, warmup_steps=4000):
        self.d_model, self.warmup_steps = float(d_model), warmup_steps

    def __call__(self, steps):
        steps = float(steps)
        return self.d_model ** -0.5 * min(steps ** -0.5, steps * self.warmup_steps ** -1.5)


def loss_function(real: torch.Tensor, pred: torch.Tensor) -> torch.Tensor:
    """models/modules.py:738-755: masked sparse CE (PAD only, the id-4/5 masks are overwritten
    :749-750), mean over all positions.  Row losses come from dsc_masked_ce_rows."""
    rows = AG.MaskedCeRows.apply(pred, real) if _DIFF else _lib.masked_ce_rows(pred, real)
    return rows.sum() / rows.numel()


def create_padding_mask(seq: torch.Tensor) -> torch.Tensor:
    """models/modules.py:757-759 -> [batch, 1, 1, seq_len] float."""
    return (seq == 0).to(torch.float32)[:, None, None, :]


def create_look_ahead_mask(size: int, device=None) -> torch.Tensor:
    """models/modules.py:761-767."""
    return 1.0 - torch.tril(torch.ones((size, size), device=device))


def create_masks(inp: torch.Tensor, tar: torch.Tensor):
    """models/modules.py:769-777."""
    enc_padding_mask = create_padding_mask(inp)
    dec_padding_mask = create_padding_mask(inp)
    look_ahead_mask = create_look_ahead_mask(tar.shape[1], device=tar.device)
    combined_mask = torch.maximum(create_padding_mask(tar), look_ahead_mask)
    return enc_padding_mask, combined_mask, dec_padding_mask
