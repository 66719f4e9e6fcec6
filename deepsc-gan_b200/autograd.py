"""Differentiable forms of the kernel groups: ``torch.autograd.Function`` is used as the tape (what
``tf.GradientTape`` is to the reference, utlis/eval.py:25-33, utlis/trainer.py:17-25, utlis/gan_train.py:15-44);
every forward AND backward computation is a libdeepsc_b200.so kernel (include/deepsc_b200.h, section K17).
There is no autograd-on-eager-ops fallback: these Functions raise on CPU tensors like the rest of ``_lib``.

Tensors are fp32 and contiguous unless stated.  Views, ``cat`` and slicing between the Functions are left to
torch (memory plumbing).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import _check, _ptr, _stream, load


TC_GEMM_MIN_MACS = 10 ** 9        # the transposing dW route pays two extra launches: vocabulary-sized products only


def _c(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


def gemm(a: torch.Tensor, b: torch.Tensor, trans_a: bool = False, trans_b: bool = False,
         out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """out[M,N] (+)= op(a) @ op(b) through dsc_gemm; a, b 2-D with unit inner stride."""
    _lib._need_cuda(a, b, out)
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = (a.shape[1], a.shape[0]) if trans_a else a.shape
    K2, N = (b.shape[1], b.shape[0]) if trans_b else b.shape
    assert K == K2, (a.shape, b.shape, trans_a, trans_b)
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32)
    if trans_a and not trans_b and M * N * K >= TC_GEMM_MIN_MACS and K % 2 == 0:
        # dW = X^T @ dY at vocabulary size: both operands are contiguous along the wrong axis for a K-major tensor-core
        # operand, so they are transposed once ([M,K] and [N,K], K = rows of the batch) and the product runs as A @ B^T
        at = torch.empty((M, K), device=a.device, dtype=torch.float32)
        bt = torch.empty((N, K), device=a.device, dtype=torch.float32)
        _check(load().dsc_transpose(a.data_ptr(), a.stride(0), at.data_ptr(), K, K, M, _stream()), "dsc_transpose")
        _check(load().dsc_transpose(b.data_ptr(), b.stride(0), bt.data_ptr(), K, K, N, _stream()), "dsc_transpose")
        _check(load().dsc_gemm_nt_tc(at.data_ptr(), K, bt.data_ptr(), K, out.data_ptr(), out.stride(0), M, N, K,
                                     int(accumulate), _stream()), "dsc_gemm_nt_tc")
        return out
    _check(load().dsc_gemm(a.data_ptr(), a.stride(0), int(trans_a), b.data_ptr(), b.stride(0), int(trans_b),
                           out.data_ptr(), out.stride(0), M, N, K, int(accumulate), _stream()), "dsc_gemm")
    return out


class Linear(torch.autograd.Function):
    """y = act(x @ w[:, :n] + bias): dsc_linear / dsc_linear_tc forward, dsc_bias_act_backward + dsc_gemm backward."""

    @staticmethod
    def forward(ctx, x, w, bias, act: int, prec: int, wfwd):
        x = _c(x)
        K, N = w.shape
        wp = w.detach() if wfwd is None else wfwd       # wfwd: the same weight with a 16-byte aligned row stride
        if wp.stride(0) % 4:
            wp = torch.zeros((K, (N + 3) // 4 * 4), device=w.device, dtype=torch.float32)
            wp[:, :N] = w.detach()
        y = _lib.linear(x, wp, None if bias is None else bias.detach(), act, n=N, prec=prec if K % 128 == 0 else 0)
        ctx.save_for_backward(x, w, y if act else None)
        ctx.act, ctx.has_bias = act, bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dy = _c(dy)
        M, N = dy.shape
        dz, db = dy, None
        if ctx.act or ctx.has_bias:
            if ctx.act:
                dz = torch.empty_like(dy)
            if ctx.has_bias:
                db = torch.empty((N,), device=dy.device, dtype=torch.float32)
            _check(load().dsc_bias_act_backward(dy.data_ptr(), dy.stride(0), _ptr(y), 0 if y is None else y.stride(0), ctx.act,
                                                dz.data_ptr() if ctx.act else None, dz.stride(0), _ptr(db), M, N, _stream()),
                   "dsc_bias_act_backward")
        dx = gemm(dz, w.detach(), trans_b=True) if ctx.needs_input_grad[0] else None
        dw = gemm(x, dz, trans_a=True) if ctx.needs_input_grad[1] else None
        return dx, dw, db, None, None, None


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], act: int = 0, prec: int = 0,
           wfwd: Optional[torch.Tensor] = None) -> torch.Tensor:
    lead = x.shape[:-1]
    y = Linear.apply(x.reshape(-1, x.shape[-1]), w, bias, act, prec, wfwd)
    return y.reshape(*lead, w.shape[1])


class AddLayerNorm(torch.autograd.Function):
    """LN_b(2 * LN_a(x + res)) or LN_a(x + res) on [rows, 128]."""

    @staticmethod
    def forward(ctx, x, res, ga, ba, gb, bb):
        x = _c(x).view(1, -1, 128)
        r = None if res is None else _c(res).view(1, -1, 128)
        out = _lib.add_layernorm(x, r, ga.detach(), ba.detach(), None if gb is None else gb.detach(),
                                 None if bb is None else bb.detach())
        ctx.save_for_backward(x, r, ga, ba, gb, bb)
        return out.view(-1, 128)

    @staticmethod
    def backward(ctx, dout):
        x, r, ga, ba, gb, bb = ctx.saved_tensors
        dout = _c(dout)
        n = x.shape[1]
        dv = torch.empty((n, 128), device=x.device, dtype=torch.float32)
        dg = torch.zeros((4, 128), device=x.device, dtype=torch.float32)
        two = gb is not None
        _check(load().dsc_add_layernorm_backward(
            x.data_ptr(), 0, _ptr(r), 0, ga.data_ptr(), ba.data_ptr(), _ptr(gb), _ptr(bb), dout.data_ptr(), 0,
            dv.data_ptr(), 0, dg[0].data_ptr(), dg[1].data_ptr(), dg[2].data_ptr() if two else None,
            dg[3].data_ptr() if two else None, n, n, _stream()), "dsc_add_layernorm_backward")
        return dv, (dv if r is not None else None), dg[0], dg[1], (dg[2] if two else None), (dg[3] if two else None)


def add_layernorm(x, res, ln_a, ln_b=None):
    shape = x.shape
    out = AddLayerNorm.apply(x.reshape(-1, 128), None if res is None else res.reshape(-1, 128), ln_a.gamma, ln_a.beta,
                             None if ln_b is None else ln_b.gamma, None if ln_b is None else ln_b.beta)
    return out.view(shape)


class MhaAttention(torch.autograd.Function):
    """softmax(q k^T / 4 + mask * -1e9) v per head; q [n,lq,128], k/v [n,lk,128] (already projected)."""

    @staticmethod
    def forward(ctx, q, k, v, mask, key_ids, causal: bool, q_off: int):
        q, k, v = _c(q), _c(k), _c(v)
        n, lq, _ = q.shape
        out = torch.empty((n, lq, 128), device=q.device, dtype=torch.float32)
        m = None
        if mask is not None:
            m = mask.to(torch.float32)
            if m.dim() == 2:
                m = m[None, None]
            m = m.broadcast_to((n, 1, lq, k.shape[1])).contiguous()
        _lib.mha_attention(q, k, v, out, mask=m, key_ids=key_ids, causal=causal, q_off=q_off)
        ctx.save_for_backward(q, k, v, m, key_ids)
        ctx.causal, ctx.q_off = causal, q_off
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, m, key_ids = ctx.saved_tensors
        dout = _c(dout)
        n, lq, _ = q.shape
        lk = k.shape[1]
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        _check(load().dsc_mha_attention_backward(
            q.data_ptr(), q.stride(1), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(1), k.stride(0),
            dout.data_ptr(), dout.stride(1), dout.stride(0), _ptr(m), 0 if m is None else m.stride(0),
            0 if m is None else m.stride(2), _ptr(key_ids), 0 if key_ids is None else key_ids.stride(0), int(ctx.causal),
            ctx.q_off, dq.data_ptr(), dq.stride(1), dq.stride(0), dk.data_ptr(), dv.data_ptr(), dk.stride(1), dk.stride(0),
            n, lq, lk, _stream()), "dsc_mha_attention_backward")
        return dq, dk, dv, None, None, None, None


class StarSatelliteAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, kv_e):
        qkv, kv_e = _c(qkv), _c(kv_e)
        S = qkv.shape[0] // 32
        att = torch.empty((S * 32, 128), device=qkv.device, dtype=torch.float32)
        _lib.star_satellite_attn(qkv, kv_e, att, S)
        ctx.save_for_backward(qkv, kv_e)
        return att

    @staticmethod
    def backward(ctx, datt):
        qkv, kv_e = ctx.saved_tensors
        datt = _c(datt)
        dqkv, dkv_e = torch.empty_like(qkv), torch.empty_like(kv_e)
        _check(load().dsc_star_satellite_attn_backward(qkv.data_ptr(), kv_e.data_ptr(), datt.data_ptr(), dqkv.data_ptr(),
                                                       dkv_e.data_ptr(), qkv.shape[0] // 32, _stream()),
               "dsc_star_satellite_attn_backward")
        return dqkv, dkv_e


class StarRelayAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv_r, kv2, n2: int):
        qkv_r = _c(qkv_r)
        kv2 = None if kv2 is None or n2 == 0 else _c(kv2)
        S = qkv_r.shape[0] // 32
        out = torch.empty((S, 128), device=qkv_r.device, dtype=torch.float32)
        _lib.star_relay_attn(qkv_r, kv2, n2 if kv2 is not None else 0, out, S)
        ctx.save_for_backward(qkv_r, kv2)
        ctx.n2 = n2 if kv2 is not None else 0
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv_r, kv2 = ctx.saved_tensors
        dout = _c(dout)
        dq = torch.empty_like(qkv_r)
        dkv2 = None if kv2 is None else torch.empty_like(kv2)
        _check(load().dsc_star_relay_attn_backward(qkv_r.data_ptr(), _ptr(kv2), 0 if kv2 is None else kv2.shape[1], ctx.n2,
                                                   dout.data_ptr(), dq.data_ptr(), _ptr(dkv2), qkv_r.shape[0] // 32, _stream()),
               "dsc_star_relay_attn_backward")
        return dq, dkv2, None


class Embed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, table, pos_table, pos0: int):
        out = _lib.embed(ids, table.detach(), pos_table, pos0)
        ctx.save_for_backward(ids, table)
        return out

    @staticmethod
    def backward(ctx, dout):
        ids, table = ctx.saved_tensors
        dout = _c(dout)
        dt = torch.zeros_like(table)
        n, ln = ids.shape
        _check(load().dsc_embed_backward(ids.data_ptr(), ids.stride(0), dout.data_ptr(), 128, dt.data_ptr(), table.shape[0],
                                         n, ln, _stream()), "dsc_embed_backward")
        return None, dt, None, None


class StarPack(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src):
        return _lib.star_pack(_c(src))

    @staticmethod
    def backward(ctx, dtile):
        dtile = _c(dtile)
        n = dtile.shape[0]
        dsrc = torch.empty((n, 31, 128), device=dtile.device, dtype=torch.float32)
        _check(load().dsc_star_pack_backward(dtile.data_ptr(), dsrc.data_ptr(), n, _stream()), "dsc_star_pack_backward")
        return dsrc


class MaskedCeRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target):
        N = logits.shape[-1]
        flat = _c(logits.reshape(-1, N))
        tgt = target.reshape(-1).to(torch.int32).contiguous()
        rows = _lib.masked_ce_rows(flat, tgt)
        ctx.save_for_backward(flat, tgt)
        ctx.shape = logits.shape
        return rows

    @staticmethod
    def backward(ctx, drows):
        flat, tgt = ctx.saved_tensors
        drows = _c(drows.reshape(-1))
        d = torch.empty_like(flat)
        _check(load().dsc_masked_ce_backward(flat.data_ptr(), flat.stride(0), tgt.data_ptr(), drows.data_ptr(), d.data_ptr(),
                                             d.stride(0), flat.shape[0], flat.shape[1], _stream()), "dsc_masked_ce_backward")
        return d.view(ctx.shape), None


def unit_dot(a: torch.Tensor, b: torch.Tensor, n_units: int) -> torch.Tensor:
    out = torch.empty((n_units,), device=a.device, dtype=torch.float32)
    _check(load().dsc_unit_dot(a.data_ptr(), b.data_ptr(), out.data_ptr(), n_units, a.numel() // n_units, _stream()),
           "dsc_unit_dot")
    return out


class PowerNormalize(torch.autograd.Function):
    """x / sqrt(factor * mean_unit(x^2)) (models/transceiver.py:91, models/gan.py:9)."""

    @staticmethod
    def forward(ctx, x, n_units: int, factor: float):
        x = _c(x)
        ss = _lib.unit_sumsq(x, n_units)
        ctx.save_for_backward(x, ss)
        ctx.n_units, ctx.factor = n_units, factor
        return _lib.power_normalize(x, n_units, factor, sumsq=ss)

    @staticmethod
    def backward(ctx, dy):
        x, ss = ctx.saved_tensors
        dy = _c(dy)
        dot = unit_dot(x, dy, ctx.n_units)
        dx = torch.empty_like(x)
        _check(load().dsc_power_normalize_backward(x.data_ptr(), ss.data_ptr(), dot.data_ptr(), ctx.factor, dy.data_ptr(),
                                                   dx.data_ptr(), ctx.n_units, x.numel() // ctx.n_units, _stream()),
               "dsc_power_normalize_backward")
        return dx, None, None


class Channel(torch.autograd.Function):
    """dsc_channel on normalised symbols x and (optionally) a perturbation p; differentiable in both."""

    @staticmethod
    def forward(ctx, x, p, n_units: int, n_std, noise, seed: int, offset: int, p_scale, h, detector: int):
        x = _c(x)
        p = None if p is None else _c(p)
        y, _ = _lib.channel(x, n_units, n_std, noise=noise, seed=seed, offset=offset, p=p, p_scale=p_scale, h=h,
                            detector=detector)
        ctx.save_for_backward(n_std, p_scale, h)
        ctx.n_units, ctx.detector, ctx.has_p = n_units, detector, p is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        n_std, p_scale, h = ctx.saved_tensors
        dy = _c(dy)
        dx = torch.empty_like(dy) if ctx.needs_input_grad[0] else None
        dp = torch.empty_like(dy) if (ctx.has_p and ctx.needs_input_grad[1]) else None
        if dx is not None or dp is not None:
            _check(load().dsc_channel_backward(dy.data_ptr(), _ptr(h), n_std.data_ptr(), ctx.detector, _ptr(p_scale), _ptr(dx),
                                               _ptr(dp), ctx.n_units, dy.numel() // ctx.n_units, _stream()),
                   "dsc_channel_backward")
        return dx, dp, None, None, None, None, None, None, None, None


class Dropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rate: float, seed: int, offset: int):
        x = _c(x)
        out = torch.empty_like(x)
        step_dev = _lib.STEP_DEV
        _check(load().dsc_dropout(x.data_ptr(), out.data_ptr(), rate, seed, offset, _ptr(step_dev), x.numel(), _stream()),
               "dsc_dropout")
        ctx.args = (rate, seed, offset, step_dev)
        return out

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        dx = torch.empty_like(dy)
        rate, seed, offset, step_dev = ctx.args
        _check(load().dsc_dropout(dy.data_ptr(), dx.data_ptr(), rate, seed, offset, _ptr(step_dev), dy.numel(), _stream()),
               "dsc_dropout")
        return dx, None, None, None


def adam_step(param: torch.Tensor, grad: torch.Tensor, m: torch.Tensor, v: torch.Tensor, lr: float, step: int,
              beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-7, grad_scale: float = 1.0,
              grad2: Optional[torch.Tensor] = None, grad2_scale: float = 0.0) -> None:
    """In-place tf.keras Adam update of a flat fp32 buffer with g = grad*grad_scale + grad2*grad2_scale."""
    _lib._need_cuda(param, grad, m, v, grad2)
    for t in (param, grad, m, v):
        assert t.is_contiguous() and t.numel() == param.numel()
    assert grad2 is None or (grad2.is_contiguous() and grad2.numel() == param.numel())
    _check(load().dsc_adam_step(param.data_ptr(), grad.data_ptr(), _ptr(grad2), m.data_ptr(), v.data_ptr(), lr, beta1, beta2,
                                eps, step, _ptr(_lib.STEP_DEV), _lib.ADAM_APPLIES_PER_STEP, grad_scale, grad2_scale,
                                param.numel(), _stream()), "dsc_adam_step")
    _lib.weights_changed()
