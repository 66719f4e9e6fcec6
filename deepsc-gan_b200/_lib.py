"""ctypes binding of libdeepsc_b200.so (the C ABI declared in include/deepsc_b200.h).

PyTorch is used here for device memory and streams only: every wrapper takes CUDA tensors,
passes raw device pointers plus the current stream, and raises if the library is missing or
returns an error.  There is no CPU or eager fallback.
"""
from __future__ import annotations

import collections
import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libdeepsc_b200.so")

_lib: Optional[C.CDLL] = None

i32, i64, u64, f32, vp = C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_void_p

_SIGNATURES = {
    "dsc_version": (C.c_int, []),
    "dsc_last_error": (C.c_char_p, []),
    "dsc_device_arch": (C.c_int, []),
    "dsc_embed": (C.c_int, [vp, i64, vp, i32, vp, vp, i64, i32, i32, i32, vp]),
    "dsc_linear": (C.c_int, [vp, i64, vp, i64, vp, vp, i64, i32, i32, i32, i32, i32, i32, i32, vp]),
    "dsc_packed_weight_bytes": (C.c_int64, [i32, i32]),
    "dsc_pack_weight": (C.c_int, [vp, i64, i32, i32, vp, vp]),
    "dsc_linear_tc": (C.c_int, [vp, i64, vp, vp, vp, i64, i32, i32, i32, i32, i32, i32, i32, vp]),
    "dsc_add_layernorm": (C.c_int, [vp, i64, vp, i64, vp, vp, vp, vp, vp, i64, i32, i32, vp]),
    "dsc_star_pack": (C.c_int, [vp, vp, i32, vp]),
    "dsc_star_satellite_attn": (C.c_int, [vp, vp, vp, i32, vp]),
    "dsc_star_interleave": (C.c_int, [vp, i64, vp, i32, i32, i32, vp]),
    "dsc_star_kv2_put": (C.c_int, [vp, vp, i32, i32, vp]),
    "dsc_target_tail_tc": (C.c_int, [vp, i64, vp, i64, vp, vp, vp, vp, vp, vp, i32, vp, i64, vp, i64, i32, i32, vp]),
    "dsc_star_cycles_tc": (C.c_int, [vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "dsc_star_relay_attn": (C.c_int, [vp, vp, i32, i32, vp, i32, vp]),
    "dsc_mha_attention": (C.c_int, [vp, i64, i64, vp, vp, i64, i64, vp, i64, i64, vp, i64, i64, vp, i64, i32, i32,
                                    i32, i32, i32, vp]),
    "dsc_unit_sumsq": (C.c_int, [vp, vp, i32, i64, vp]),
    "dsc_power_normalize": (C.c_int, [vp, vp, f32, vp, i32, i64, vp]),
    "dsc_channel": (C.c_int, [vp, vp, f32, vp, u64, u64, vp, vp, f32, vp, vp, vp, i32, vp, vp, i32, i64, vp]),
    "dsc_vocab_argmax": (C.c_int, [vp, i64, vp, i64, vp, vp, i64, vp, i64, vp, i64, i32, i32, i32, vp]),
    "dsc_vocab_argmax_workspace": (C.c_int64, [i32, i32]),
    "dsc_vocab_argmax_tc_workspace": (C.c_int64, [i32, i32]),
    "dsc_vocab_argmax_tc": (C.c_int, [vp, i64, vp, vp, vp, i64, vp, i64, i32, i32, i32, vp]),
    "dsc_argmax_rows": (C.c_int, [vp, i64, vp, i64, i32, i32, vp]),
    "dsc_masked_ce_rows": (C.c_int, [vp, i64, vp, vp, i32, i32, vp]),
    "dsc_bleu_counts": (C.c_int, [vp, i32, vp, i32, vp, i32, vp]),
    "dsc_fgm_normalize": (C.c_int, [vp, vp, f32, i32, i32, i32, vp]),
    # backward kernels (K17)
    "dsc_gemm": (C.c_int, [vp, i64, i32, vp, i64, i32, vp, i64, i32, i32, i32, i32, vp]),
    "dsc_gemm_nt_tc": (C.c_int, [vp, i64, vp, i64, vp, i64, i32, i32, i32, i32, vp]),
    "dsc_transpose": (C.c_int, [vp, i64, vp, i64, i32, i32, vp]),
    "dsc_bias_act_backward": (C.c_int, [vp, i64, vp, i64, i32, vp, i64, vp, i32, i32, vp]),
    "dsc_add_layernorm_backward": (C.c_int, [vp, i64, vp, i64, vp, vp, vp, vp, vp, i64, vp, i64, vp, vp, vp, vp, i32, i32, vp]),
    "dsc_mha_attention_backward": (C.c_int, [vp, i64, i64, vp, vp, i64, i64, vp, i64, i64, vp, i64, i64, vp, i64, i32, i32,
                                             vp, i64, i64, vp, vp, i64, i64, i32, i32, i32, vp]),
    "dsc_star_satellite_attn_backward": (C.c_int, [vp, vp, vp, vp, vp, i32, vp]),
    "dsc_star_relay_attn_backward": (C.c_int, [vp, vp, i32, i32, vp, vp, vp, i32, vp]),
    "dsc_embed_backward": (C.c_int, [vp, i64, vp, i64, vp, i32, i32, i32, vp]),
    "dsc_star_pack_backward": (C.c_int, [vp, vp, i32, vp]),
    "dsc_masked_ce_backward": (C.c_int, [vp, i64, vp, vp, vp, i64, i32, i32, vp]),
    "dsc_unit_dot": (C.c_int, [vp, vp, vp, i32, i64, vp]),
    "dsc_power_normalize_backward": (C.c_int, [vp, vp, vp, f32, vp, vp, i32, i64, vp]),
    "dsc_channel_backward": (C.c_int, [vp, vp, vp, i32, vp, vp, vp, i32, i64, vp]),
    "dsc_dropout": (C.c_int, [vp, vp, f32, u64, u64, vp, i64, vp]),
    "dsc_adam_step": (C.c_int, [vp, vp, vp, vp, vp, f32, f32, f32, f32, i32, vp, i32, f32, f32, i64, vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def use_debug_library() -> None:
    """Developer tools only (tools/star_trace.py, tools/umma_probe.py): bind libdeepsc_b200_debug.so - the product sources
    compiled with -DDSC_DEBUG_TOOLS=1 plus csrc/debug/*.cu (`python deepsc-gan_b200/build.py --debug`) - instead of the
    product library.  Must be called before the first load()."""
    global LIB_PATH
    assert _lib is None, "use_debug_library() must precede the first load()"
    LIB_PATH = os.path.join(_HERE, "csrc", "libdeepsc_b200_debug.so")
    _SIGNATURES["dsc_umma_probe"] = (C.c_int, [i32, i32, i32, vp, vp])
    _SIGNATURES["dsc_debug_star_trace"] = (C.c_int, [vp])


def load() -> C.CDLL:
    """Load the shared library (no GPU needed for this step) and set the prototypes."""
    global _lib, LIB_PATH
    if _lib is None:
        if os.environ.get("DSC_LIB_PATH"):              # developer A/B builds (tools/pp_variants.sh); never set in product use
            LIB_PATH = os.environ["DSC_LIB_PATH"]
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "deepsc_gan_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class DscError(RuntimeError):
    pass


# kernel launches issued through this binding (bench.py reports them as gpu_launches)
_KERNELS_PER_CALL = {"dsc_vocab_argmax": 2}
STATS = {"launches": 0}
# when set to a list, the wrappers named in PROFILE_OPS append (op, start_event, end_event, meta) for bench.py's
# roofline accounting (CUDA events on the launching stream around the single kernel launch)
PROFILE = None
STAR_FIRST_SAT_DONE = 0x100          # include/deepsc_b200.h DSC_STAR_FIRST_SAT_DONE
STAR_NO_FINAL_RELAY = 0x200          # include/deepsc_b200.h DSC_STAR_NO_FINAL_RELAY
STAR_FORM_ONE_TILE = 0x400           # include/deepsc_b200.h DSC_STAR_FORM_ONE_TILE / _TWO_TILE: force a kernel form
STAR_FORM_TWO_TILE = 0x800
STAR_FORM = 0                        # OR-ed into every dsc_star_cycles_tc call (0 = let the library pick by batch size)
PROFILE_OPS = ("dsc_star_cycles_tc",)


class _timed:
    """with _timed("dsc_x", meta): launch  -> PROFILE.append(("dsc_x", ev0, ev1, meta)) when profiling is on."""

    def __init__(self, op: str, meta):
        self.on = PROFILE is not None and op in PROFILE_OPS and not torch.cuda.is_current_stream_capturing()
        self.op, self.meta = op, meta

    def __enter__(self):
        if self.on:
            self.ev0, self.ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.ev0.record()

    def __exit__(self, *exc):
        if self.on:
            self.ev1.record()
            PROFILE.append((self.op, self.ev0, self.ev1, self.meta))
        return False


def _check(rc: int, what: str) -> None:
    STATS["launches"] += _KERNELS_PER_CALL.get(what, 1)
    if rc != 0:
        msg = load().dsc_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(msg or what)
        raise DscError(f"{what} failed ({rc}): {msg}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("deepsc_gan_b200 kernels need CUDA tensors (there is no CPU fallback)")


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"expected float32, got {t.dtype}")
    return t


# --------------------------------------------------------------------------- wrappers
def version() -> int:
    return load().dsc_version()


def embed(ids: torch.Tensor, table: torch.Tensor, pos_table: torch.Tensor, pos0: int = 0,
          out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ids [n, len] int32 (last-dim stride 1, any row stride) -> [n, len, 128]."""
    _need_cuda(ids, table, pos_table)
    assert ids.dtype == torch.int32 and ids.dim() == 2 and (ids.shape[1] == 1 or ids.stride(1) == 1)
    n, ln = ids.shape
    if out is None:
        out = torch.empty((n, ln, 128), device=ids.device, dtype=torch.float32)
    _check(load().dsc_embed(ids.data_ptr(), ids.stride(0), _f32(table).data_ptr(), table.shape[0],
                            pos_table.data_ptr(), out.data_ptr(), 128, n, ln, pos0, _stream()), "dsc_embed")
    return out


_PACK_CACHE = collections.OrderedDict()      # key -> [stamp, blob, source weight, pinned]
_PACK_CACHE_MAX = 512
# bumped by anything that rewrites parameters through raw pointers (the Adam kernel): torch's own _version counter
# does not see those writes, and the packed / padded weight caches key on both
WEIGHT_EPOCH = 0


# device-side count of completed training steps (int64 scalar tensor) while a graph-replayed training step is being
# captured / replayed; None otherwise.  The dropout and Adam kernels read it (see include/deepsc_b200.h).
STEP_DEV = None
ADAM_APPLIES_PER_STEP = 1


def weights_changed() -> None:
    global WEIGHT_EPOCH
    WEIGHT_EPOCH += 1


def packed_weight(w: torch.Tensor, n: int) -> torch.Tensor:
    """bf16 hi/lo UMMA image of a Keras-layout weight, cached per storage and re-packed IN PLACE when the weight has
    been modified (torch version counter or WEIGHT_EPOCH), so the blob address is stable for the life of the entry.
    An entry touched while a CUDA graph is being captured is pinned: the graph has the blob's raw pointer baked in
    (and re-packs into it on replay), so it is never evicted.  Other entries are evicted least-recently-used beyond
    _PACK_CACHE_MAX; eviction only drops the cache's reference, a caller still holding the blob keeps it alive."""
    key = (w.data_ptr(), tuple(w.shape), w.stride(0), n, w.device.index)
    stamp = (w._version, WEIGHT_EPOCH)
    hit = _PACK_CACHE.get(key)
    capturing = torch.cuda.is_current_stream_capturing()
    if hit is None:
        blob = torch.empty((load().dsc_packed_weight_bytes(w.shape[0], n),), device=w.device, dtype=torch.uint8)
        hit = [None, blob, w, capturing]     # the source is kept alive so its data_ptr cannot be recycled under the key
        _PACK_CACHE[key] = hit
        if len(_PACK_CACHE) > _PACK_CACHE_MAX:
            for k in [k for k, e in _PACK_CACHE.items() if not e[3] and k != key][: len(_PACK_CACHE) - _PACK_CACHE_MAX * 3 // 4]:
                del _PACK_CACHE[k]
    else:
        _PACK_CACHE.move_to_end(key)
        hit[3] = hit[3] or capturing
    if hit[0] != stamp:
        _check(load().dsc_pack_weight(_f32(w).data_ptr(), w.stride(0), w.shape[0], n, hit[1].data_ptr(), _stream()),
               "dsc_pack_weight")
        hit[0] = stamp
    return hit[1]


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], act: int = 0,
           out: Optional[torch.Tensor] = None, n: Optional[int] = None, row_mod: int = 0, row_skip: int = 0,
           prec: int = 1) -> torch.Tensor:
    """y = act(x @ w + bias).  x [M, K] (row stride free), w [K, >=N] Keras layout, out [M, N] (row stride free).
    prec 1 / 2: tcgen05 (bf16x3 / bf16) when K is a multiple of 128; prec 0, or any other K: the fp32 FFMA kernel."""
    _need_cuda(x, w, bias, out)
    assert x.dim() == 2 and w.dim() == 2 and x.stride(1) == 1 and w.stride(1) == 1
    M, K = x.shape
    N = w.shape[1] if n is None else n
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=torch.float32)
    assert out.dim() == 2 and out.shape[0] == M and out.shape[1] == N and out.stride(1) == 1
    if M == 0:                       # empty batch: torch hands out a null data_ptr, nothing to launch
        return out
    if prec != 0 and K % 128 == 0:
        blob = packed_weight(w, N)
        with _timed("dsc_linear_tc", (M, K, N)):
            _check(load().dsc_linear_tc(_f32(x).data_ptr(), x.stride(0), blob.data_ptr(), _ptr(bias), out.data_ptr(),
                                        out.stride(0), M, K, N, act, row_mod, row_skip, prec, _stream()),
                   "dsc_linear_tc")
    else:
        with _timed("dsc_linear", (M, K, N)):
            _check(load().dsc_linear(_f32(x).data_ptr(), x.stride(0), _f32(w).data_ptr(), w.stride(0), _ptr(bias),
                                     out.data_ptr(), out.stride(0), M, K, N, act, row_mod, row_skip, 0, _stream()),
                   "dsc_linear")
    return out


def add_layernorm(x: torch.Tensor, res: Optional[torch.Tensor], gamma_a, beta_a, gamma_b=None, beta_b=None,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x, res, out: [groups, rows, 128] views whose rows are contiguous within a group (group stride free)."""
    _need_cuda(x, res, out)
    assert x.dim() == 3 and x.shape[2] == 128 and x.stride(2) == 1 and x.stride(1) == 128
    g, r, _ = x.shape
    if out is None:
        out = torch.empty((g, r, 128), device=x.device, dtype=torch.float32)
    assert out.shape == x.shape and out.stride(1) == 128
    if res is not None:
        assert res.shape == x.shape and res.stride(1) == 128 and res.stride(2) == 1
    _check(load().dsc_add_layernorm(x.data_ptr(), x.stride(0), _ptr(res), 0 if res is None else res.stride(0),
                                    gamma_a.data_ptr(), beta_a.data_ptr(), _ptr(gamma_b), _ptr(beta_b),
                                    out.data_ptr(), out.stride(0), g * r, r, _stream()), "dsc_add_layernorm")
    return out


def star_pack(src: torch.Tensor, tile: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[n, 31, 128] contiguous -> star tile [n, 32, 128] with row 31 = mean of the 31 rows."""
    _need_cuda(src)
    assert src.is_contiguous() and src.shape[1:] == (31, 128)
    n = src.shape[0]
    if tile is None:
        tile = torch.empty((n, 32, 128), device=src.device, dtype=torch.float32)
    _check(load().dsc_star_pack(_f32(src).data_ptr(), tile.data_ptr(), n, _stream()), "dsc_star_pack")
    return tile


def star_satellite_attn(qkv: torch.Tensor, kv_e: torch.Tensor, att: torch.Tensor, n_sent: int) -> torch.Tensor:
    _need_cuda(qkv, kv_e, att)
    assert qkv.is_contiguous() and kv_e.is_contiguous() and att.is_contiguous()
    _check(load().dsc_star_satellite_attn(qkv.data_ptr(), kv_e.data_ptr(), att.data_ptr(), n_sent, _stream()),
           "dsc_star_satellite_attn")
    return att


def star_interleave(src: torch.Tensor, dst: torch.Tensor, group_rows: int) -> torch.Tensor:
    """Row-major [n_groups, group_rows, width] (contiguous within a group) -> interleaved [group][width/4][rows][4]."""
    _need_cuda(src, dst)
    assert src.dim() == 3 and src.shape[1] == group_rows and src.stride(2) == 1 and src.stride(1) == src.shape[2]
    assert dst.is_contiguous() and dst.numel() == src.numel()
    _check(load().dsc_star_interleave(_f32(src).data_ptr(), src.stride(0), dst.data_ptr(), src.shape[0], group_rows,
                                      src.shape[2], _stream()), "dsc_star_interleave")
    return dst


def star_kv2_put(vals: torch.Tensor, kv2i: torch.Tensor, row_index: int) -> None:
    """vals [n_sent, 256] -> row `row_index` of the interleaved h2 cache [n_sent, 64, 32, 4]."""
    _need_cuda(vals, kv2i)
    assert vals.is_contiguous() and vals.shape[1] == 256 and kv2i.is_contiguous()
    _check(load().dsc_star_kv2_put(vals.data_ptr(), kv2i.data_ptr(), row_index, vals.shape[0], _stream()),
           "dsc_star_kv2_put")


def star_cycles_tc(xi0: torch.Tensor, s0: torch.Tensor, q0: torch.Tensor, kvei: torch.Tensor, kv2i: Optional[torch.Tensor],
                   n2: int, w_grouped: torch.Tensor, wo: torch.Tensor, wkv_relay: torch.Tensor, wo_relay: torch.Tensor,
                   wq_relay: torch.Tensor, bias_o: torch.Tensor, bias_o_relay: torch.Tensor, x_rowmajor: torch.Tensor,
                   n_sent: int, n_cycles: int, prec: int) -> torch.Tensor:
    """All star cycles of a layer in one persistent launch (see include/deepsc_b200.h dsc_star_cycles_tc)."""
    _need_cuda(xi0, s0, q0, kvei, kv2i, w_grouped, wo, wkv_relay, wo_relay, wq_relay, bias_o, bias_o_relay, x_rowmajor)
    for t in (xi0, s0, q0, kvei, x_rowmajor, bias_o, bias_o_relay):
        assert t.is_contiguous()
    assert kv2i is None or (kv2i.is_contiguous() and kv2i.numel() == n_sent * 8192)
    assert x_rowmajor.numel() == n_sent * 4096 and xi0.numel() >= n_sent * 4096 and kvei.numel() == n_sent * 8192
    with _timed("dsc_star_cycles_tc", (n_sent, n_cycles, n2, bool(prec & STAR_FIRST_SAT_DONE),
                                       bool(prec & STAR_NO_FINAL_RELAY))):
        _check(load().dsc_star_cycles_tc(xi0.data_ptr(), s0.data_ptr(), q0.data_ptr(), kvei.data_ptr(), _ptr(kv2i), n2,
                                         packed_weight(w_grouped, 384).data_ptr(), packed_weight(wo, 128).data_ptr(),
                                         packed_weight(wkv_relay, 256).data_ptr(), packed_weight(wo_relay, 128).data_ptr(),
                                         packed_weight(wq_relay, 128).data_ptr(), bias_o.data_ptr(),
                                         bias_o_relay.data_ptr(), x_rowmajor.data_ptr(), n_sent, n_cycles, prec | STAR_FORM,
                                         _stream()),
               "dsc_star_cycles_tc")
    return x_rowmajor


def target_tail_tc(attn: torch.Tensor, resid: torch.Tensor, wo: torch.Tensor, bias_o: torch.Tensor, gamma: torch.Tensor,
                   beta: torch.Tensor, wkv_relay: torch.Tensor, kv2i: Optional[torch.Tensor], row_index: int, prec: int,
                   kv_rows: Optional[torch.Tensor] = None, h2_out: Optional[torch.Tensor] = None) -> None:
    """Fused Dense + residual + LayerNorm + relay k|v projection + key-cache write (include/deepsc_b200.h)."""
    _need_cuda(attn, resid, wo, bias_o, gamma, beta, wkv_relay, kv2i, kv_rows, h2_out)
    M = attn.shape[0]
    assert attn.dim() == 2 and attn.shape[1] == 128 and attn.stride(1) == 1 and resid.shape == attn.shape and resid.stride(1) == 1
    assert kv2i is None or (kv2i.is_contiguous() and kv2i.numel() == M * 8192)
    assert kv_rows is None or (kv_rows.shape == (M, 256) and kv_rows.stride(1) == 1)
    assert h2_out is None or (h2_out.shape == (M, 128) and h2_out.stride(1) == 1)
    _check(load().dsc_target_tail_tc(attn.data_ptr(), attn.stride(0), resid.data_ptr(), resid.stride(0),
                                     packed_weight(wo, 128).data_ptr(), bias_o.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                     packed_weight(wkv_relay, 256).data_ptr(), _ptr(kv2i), row_index,
                                     _ptr(kv_rows), 0 if kv_rows is None else kv_rows.stride(0),
                                     _ptr(h2_out), 0 if h2_out is None else h2_out.stride(0), M, prec, _stream()),
           "dsc_target_tail_tc")


def star_relay_attn(qkv_r: torch.Tensor, kv2: Optional[torch.Tensor], n2: int, out: torch.Tensor, n_sent: int):
    _need_cuda(qkv_r, kv2, out)
    kv2_rows = 0 if kv2 is None else kv2.shape[1]
    if kv2 is not None:
        assert kv2.is_contiguous() and kv2.shape[2] == 256
    _check(load().dsc_star_relay_attn(qkv_r.data_ptr(), _ptr(kv2), kv2_rows, n2, out.data_ptr(), n_sent, _stream()),
           "dsc_star_relay_attn")
    return out


def mha_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor,
                  mask: Optional[torch.Tensor] = None, key_ids: Optional[torch.Tensor] = None,
                  causal: bool = False, q_off: int = 0) -> torch.Tensor:
    """q [n, lq, 128], k/v [n, lk, 128] (views with free row/batch strides, k and v sharing strides);
    mask broadcastable to [n, 1, lq, lk] float 0/1; key_ids [n, >=lk] int32."""
    _need_cuda(q, k, v, out, mask, key_ids)
    n, lq, _ = q.shape
    lk = k.shape[1]
    assert k.stride() == v.stride() and q.stride(2) == 1 and k.stride(2) == 1 and out.stride(2) == 1
    mb = mq = 0
    mptr = None
    if mask is not None:
        m = _f32(mask)
        if m.dim() == 2:
            m = m[None, None]
        m = m.broadcast_to((n, 1, lq, lk))
        if m.stride(3) != 1 and lk > 1:
            m = m.contiguous()
        mb, mq, mptr = m.stride(0), m.stride(2), m.data_ptr()
        mask = m  # keep alive
    kis = 0
    if key_ids is not None:
        assert key_ids.dtype == torch.int32 and key_ids.shape[0] == n and key_ids.shape[1] >= lk
        assert key_ids.shape[1] == 1 or key_ids.stride(1) == 1
        kis = key_ids.stride(0)
    _check(load().dsc_mha_attention(q.data_ptr(), q.stride(1), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(1),
                                    k.stride(0), out.data_ptr(), out.stride(1), out.stride(0), mptr, mb, mq,
                                    _ptr(key_ids), kis, int(causal), q_off, n, lq, lk, _stream()),
           "dsc_mha_attention")
    return out


def unit_sumsq(x: torch.Tensor, n_units: int) -> torch.Tensor:
    _need_cuda(x)
    assert x.is_contiguous() and x.numel() % n_units == 0
    out = torch.empty((n_units,), device=x.device, dtype=torch.float32)
    _check(load().dsc_unit_sumsq(_f32(x).data_ptr(), out.data_ptr(), n_units, x.numel() // n_units, _stream()),
           "dsc_unit_sumsq")
    return out


def power_normalize(x: torch.Tensor, n_units: int, factor: float = 1.0, sumsq: Optional[torch.Tensor] = None):
    """x / sqrt(factor * mean(x^2)) per unit (contiguous x split evenly into n_units)."""
    _need_cuda(x, sumsq)
    assert x.is_contiguous()
    if sumsq is None:
        sumsq = unit_sumsq(x, n_units)
    out = torch.empty_like(x)
    _check(load().dsc_power_normalize(_f32(x).data_ptr(), sumsq.data_ptr(), factor, out.data_ptr(), n_units,
                                      x.numel() // max(n_units, 1), _stream()), "dsc_power_normalize")
    return out


def channel(x: torch.Tensor, n_units: int, n_std: torch.Tensor, *, x_sumsq=None, x_factor: float = 1.0,
            noise=None, seed: int = 0, offset: int = 0, p=None, p_sumsq=None, p_factor: float = 1.0,
            p_scale=None, h=None, detector: int = 0, want_x_norm: bool = False):
    """The fused channel kernel; returns (y, x_norm or None)."""
    _need_cuda(x, n_std, x_sumsq, noise, p, p_sumsq, p_scale, h)
    assert x.is_contiguous() and x.numel() % max(n_units, 1) == 0
    for t in (noise, p):
        assert t is None or (t.is_contiguous() and t.numel() == x.numel())
    assert n_std.numel() == n_units and (h is None or (h.is_contiguous() and h.numel() == 2 * n_units))
    y = torch.empty_like(x)
    xn = torch.empty_like(x) if want_x_norm else None
    _check(load().dsc_channel(_f32(x).data_ptr(), _ptr(x_sumsq), x_factor, _ptr(noise), seed, offset, _ptr(p),
                              _ptr(p_sumsq), p_factor, _ptr(p_scale), _ptr(h), _f32(n_std).data_ptr(), detector,
                              y.data_ptr(), _ptr(xn), n_units, x.numel() // max(n_units, 1), _stream()),
           "dsc_channel")
    return y, xn


_VOCAB_WS = {}


def _vocab_tc_workspace(M: int, n_vocab: int, device) -> torch.Tensor:
    """Zero-filled once; the kernel's arrival counters reset themselves, calls on one stream are ordered."""
    key = (M, n_vocab, torch.device(device).index, torch.cuda.current_stream().cuda_stream)
    ws = _VOCAB_WS.get(key)
    if ws is None:
        ws = torch.zeros((load().dsc_vocab_argmax_tc_workspace(M, n_vocab),), device=device, dtype=torch.uint8)
        _VOCAB_WS[key] = ws
    return ws


def vocab_argmax(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, n_vocab: int, ids_out: torch.Tensor,
                 logits: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None,
                 prec: int = 1) -> torch.Tensor:
    """ids_out [M] int32 view (any stride) <- argmax(x @ w[:, :n_vocab] + bias)."""
    _need_cuda(x, w, bias, ids_out, logits, workspace)
    M = x.shape[0]
    assert x.dim() == 2 and x.shape[1] == 128 and x.stride(1) == 1 and ids_out.dtype == torch.int32
    assert ids_out.dim() == 1 and ids_out.shape[0] == M
    ld_logits = 0
    if logits is not None:
        assert logits.dim() == 2 and logits.shape == (M, n_vocab) and logits.stride(1) == 1
        ld_logits = logits.stride(0)
    elif workspace is None and prec == 0:
        workspace = torch.empty((load().dsc_vocab_argmax_workspace(M, n_vocab),), device=x.device,
                                dtype=torch.float32)
    if prec != 0 and logits is None:
        # fused tcgen05 projection + running argmax: the logits never reach HBM
        ws = _vocab_tc_workspace(M, n_vocab, x.device)
        with _timed("dsc_vocab_argmax_tc", (M, n_vocab)):
            _check(load().dsc_vocab_argmax_tc(_f32(x).data_ptr(), x.stride(0), packed_weight(w, n_vocab).data_ptr(),
                                              bias.data_ptr(), ids_out.data_ptr(), ids_out.stride(0) if M > 1 else 1,
                                              ws.data_ptr(), ws.numel(), M, n_vocab, prec, _stream()),
                   "dsc_vocab_argmax_tc")
        return ids_out
    if prec != 0:
        # `predictions` requested: tensor-core projection into the caller's logits rows, then the row argmax kernel
        linear(x, w, bias, out=logits, n=n_vocab, prec=prec)
        _check(load().dsc_argmax_rows(logits.data_ptr(), logits.stride(0), ids_out.data_ptr(),
                                      ids_out.stride(0) if M > 1 else 1, M, n_vocab, _stream()), "dsc_argmax_rows")
        return ids_out
    _check(load().dsc_vocab_argmax(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), bias.data_ptr(),
                                   ids_out.data_ptr(), ids_out.stride(0) if M > 1 else 1, _ptr(logits), ld_logits,
                                   _ptr(workspace), 0 if workspace is None else workspace.numel(),
                                   M, n_vocab, prec, _stream()), "dsc_vocab_argmax")
    return ids_out


def argmax_rows(logits: torch.Tensor) -> torch.Tensor:
    """[..., N] float32 (last dim contiguous, uniform row stride) -> [...] int32."""
    _need_cuda(logits)
    N = logits.shape[-1]
    flat = logits.reshape(-1, N)
    ids = torch.empty((flat.shape[0],), device=logits.device, dtype=torch.int32)
    _check(load().dsc_argmax_rows(_f32(flat).data_ptr(), flat.stride(0), ids.data_ptr(), 1, flat.shape[0], N,
                                  _stream()), "dsc_argmax_rows")
    return ids.reshape(logits.shape[:-1])


def masked_ce_rows(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    _need_cuda(logits, target)
    N = logits.shape[-1]
    flat = logits.reshape(-1, N)
    tgt = target.reshape(-1).to(torch.int32).contiguous()
    assert tgt.numel() == flat.shape[0]
    out = torch.empty((flat.shape[0],), device=logits.device, dtype=torch.float32)
    _check(load().dsc_masked_ce_rows(_f32(flat).data_ptr(), flat.stride(0), tgt.data_ptr(), out.data_ptr(),
                                     flat.shape[0], N, _stream()), "dsc_masked_ce_rows")
    return out.reshape(target.shape)


def bleu_counts(ref: torch.Tensor, hyp: torch.Tensor) -> torch.Tensor:
    """ref [n, Lr], hyp [n, Lh] int32 contiguous -> counts [n, 10] int32."""
    _need_cuda(ref, hyp)
    assert ref.dtype == torch.int32 and hyp.dtype == torch.int32 and ref.is_contiguous() and hyp.is_contiguous()
    n = ref.shape[0]
    out = torch.empty((n, 10), device=ref.device, dtype=torch.int32)
    if n == 0:
        return out
    _check(load().dsc_bleu_counts(ref.data_ptr(), ref.shape[1], hyp.data_ptr(), hyp.shape[1], out.data_ptr(), n,
                                  _stream()), "dsc_bleu_counts")
    return out


def fgm_normalize(g: torch.Tensor, n_units: int, epsilon: float = 1.0) -> torch.Tensor:
    """g [n_units*samples, ...] contiguous float32 -> p of the same shape."""
    _need_cuda(g)
    assert g.is_contiguous() and g.shape[0] % n_units == 0
    samples = g.shape[0] // n_units
    p = torch.empty_like(g)
    _check(load().dsc_fgm_normalize(_f32(g).data_ptr(), p.data_ptr(), float(epsilon), n_units, samples,
                                    g.numel() // g.shape[0], _stream()), "dsc_fgm_normalize")
    return p
