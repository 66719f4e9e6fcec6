"""Per-kernel parity through the C ABI (-m gpu): every dsc_* entry point against the CPU oracle on the
same seeded inputs.  Integer results must be bit-exact; floating point within the stated tolerance."""
import json
import os
import random

import numpy as np
import pytest
import torch

import _cases
from oracle import bleu_oracle as B, deepsc_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-3   # north_star tolerance for symbols / logits (relative)


def rel_err(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.fixture(scope="module")
def L(dev):
    from deepsc_gan_b200 import _lib
    return _lib


def test_device_is_blackwell(L, dev):
    assert L.load().dsc_device_arch() >= 100


@pytest.mark.parametrize("M,K,N,act", [(1984, 128, 384, 0), (64, 128, 128, 1), (1984, 16, 128, 1), (1984, 256, 16, 0),
                                       (300, 512, 128, 0), (64, 128, 22234, 0), (129, 128, 256, 1)])
def test_linear_fp32(L, dev, M, K, N, act):
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(K, N, generator=g) / np.sqrt(K)
    b = torch.randn(N, generator=g)
    ref = x.double() @ w.double() + b.double()
    if act:
        ref = torch.relu(ref)
    ldw = (N + 127) // 128 * 128
    wd = torch.zeros(K, ldw, device=dev)
    wd[:, :N] = w.to(dev)
    y = L.linear(x.to(dev), wd, b.to(dev), act=act, n=N, prec=0)
    assert y.shape == (M, N)
    assert rel_err(y, ref) < 2e-6


def test_linear_row_skip_and_strided_output(L, dev):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(128, 128, generator=g).to(dev)
    w = torch.randn(128, 128, generator=g).to(dev)
    out = torch.full((128, 128), 7.0, device=dev)
    L.linear(x, w, None, out=out, row_mod=32, row_skip=31, prec=0)
    ref = (x.double() @ w.double()).float()
    keep = torch.arange(128, device=dev) % 32 != 31
    assert torch.allclose(out[keep], ref[keep], rtol=1e-5, atol=1e-4)
    assert bool((out[~keep] == 7.0).all())
    tile = torch.zeros((4, 32, 128), device=dev)
    L.linear(x[:4], w, None, out=tile[:, 31, :], prec=0)
    assert torch.allclose(tile[:, 31, :], ref[:4], rtol=1e-5, atol=1e-4) and float(tile[:, :31].abs().max()) == 0.0


def test_embed_matches_oracle(L, dev):
    P = _cases.params("Transeiver_Star")
    ids = _cases.synthetic_unit(0)
    ref = O.embed(P, "semantic_encoder", ids.long(), O.positional_table())
    out = L.embed(ids.to(dev), P["semantic_encoder/embedding/embeddings"].to(dev), O.positional_table().to(dev))
    assert rel_err(out, ref) < 1e-6
    # one token per sentence at position 5 read with a stride (the greedy step form)
    out1 = L.embed(ids.to(dev)[:, 5:6], P["semantic_encoder/embedding/embeddings"].to(dev),
                   O.positional_table().to(dev), pos0=5)
    assert rel_err(out1[:, 0], ref[:, 5]) < 1e-6


def test_add_layernorm_single_and_double(L, dev):
    P = _cases.params("Transeiver_Star")
    g = torch.Generator().manual_seed(2)
    x, r = torch.randn(6, 31, 128, generator=g), torch.randn(6, 31, 128, generator=g)
    pre = "semantic_decoder/dec_layers"
    o1 = O.layernorm(P, pre + "/layernorm2", x + r)
    o2 = O.layernorm(P, pre + "/layernorm3", o1 + o1)
    ga, ba = P[pre + "/layernorm2/gamma"].to(dev), P[pre + "/layernorm2/beta"].to(dev)
    gb, bb = P[pre + "/layernorm3/gamma"].to(dev), P[pre + "/layernorm3/beta"].to(dev)
    assert rel_err(L.add_layernorm(x.to(dev), r.to(dev), ga, ba), o1) < 2e-6
    assert rel_err(L.add_layernorm(x.to(dev), r.to(dev), ga, ba, gb, bb), o2) < 2e-6
    # tile-strided input rows, compact output
    xt = torch.zeros(6, 32, 128, device=dev)
    xt[:, :31] = x.to(dev)
    rt = torch.zeros(6, 32, 128, device=dev)
    rt[:, :31] = r.to(dev)
    assert rel_err(L.add_layernorm(xt[:, :31], rt[:, :31], ga, ba, gb, bb), o2) < 2e-6


def test_star_pack_mean_row(L, dev):
    g = torch.Generator().manual_seed(3)
    e = torch.randn(5, 31, 128, generator=g)
    t = L.star_pack(e.to(dev)).cpu()
    assert torch.equal(t[:, :31], e)
    assert torch.allclose(t[:, 31], e.mean(1), atol=1e-6)


@pytest.mark.parametrize("prec,tol", [(0, 5e-6), (1, 1e-4), (2, 5e-2)])
@pytest.mark.parametrize("n2", [0, 1, 17, 30])
def test_star_cycle_kernels_match_literal_oracle(L, dev, n2, prec, tol, monkeypatch):
    """Full cycles (satellite + relay) against the literal 5-key concat form of modules.py:289-306, for the
    fp32 kernels (prec 0) and the fused tcgen05 kernels (prec 1 bf16x3, prec 2 bf16)."""
    import deepsc_gan_b200.models.modules as M0
    monkeypatch.setattr(M0, "PREC", prec)
    P = _cases.params("Transeiver_Star", gain=3.0)
    pre = "semantic_decoder/dec_layers"
    g = torch.Generator().manual_seed(10 + n2)
    S = 8
    e = torch.randn(S, 31, 128, generator=g)
    h2 = torch.randn(S, 30, 128, generator=g)[:, :n2] if n2 else None
    h_ref, s_ref = O._star_cycles(P, pre, e, h2, 1, "multi_att_relay")
    h_ref3, s_ref3 = O._star_cycles(P, pre, e, h2, 3, "multi_att_relay")

    import deepsc_gan_b200.models.modules as M
    sat, relay = M.sublayer1(128, 8).to(dev), M.sublayer1(128, 8).to(dev)
    with torch.no_grad():
        for mod, name in ((sat, "multi_att_satellite"), (relay, "multi_att_relay")):
            for w in ("wq", "wk", "wv"):
                getattr(mod, w).kernel.copy_(P[f"{pre}/{name}/{w}/kernel"])
            mod.dense.kernel.copy_(P[f"{pre}/{name}/dense/kernel"])
            mod.dense.bias.copy_(P[f"{pre}/{name}/dense/bias"])
    tile = L.star_pack(e.to(dev))
    kv2 = None
    if n2:
        kv2 = torch.zeros(S, 30, 256, device=dev)
        kv2[:, :n2] = L.linear(h2.reshape(-1, 128).to(dev), relay._packed("kv"), None, prec=prec).view(S, n2, 256)
    x = M.star_cycles(tile, sat, relay, 1, kv2, n2).clone()
    torch.cuda.synchronize()
    assert rel_err(x[:, :31], h_ref) < tol and rel_err(x[:, 31], s_ref) < tol
    x3 = M.star_cycles(tile, sat, relay, 3, kv2, n2)
    torch.cuda.synchronize()
    assert rel_err(x3[:, :31], h_ref3) < 4 * tol and rel_err(x3[:, 31], s_ref3) < 4 * tol


@pytest.mark.parametrize("prec", [1, 2])
@pytest.mark.parametrize("S,n2", [(8, 0), (12, 17), (600, 30)])
def test_star_cycles_first_satellite_half_cached(L, dev, S, n2, prec, monkeypatch):
    """DSC_STAR_FIRST_SAT_DONE (the greedy decoder's once-per-batch satellite half of cycle 0) gives bit-identical
    tiles to the plain call: same products, same order."""
    import deepsc_gan_b200.models.modules as M
    monkeypatch.setattr(M, "PREC", prec)
    torch.manual_seed(3)
    sat, relay = M.sublayer1(128, 8).to(dev), M.sublayer1(128, 8).to(dev)
    tile = L.star_pack(torch.randn(S, 31, 128, device=dev))
    kv2 = torch.randn(S, 30, 256, device=dev) if n2 else None
    ws = M.StarWorkspace(S, dev)
    M.prepare_kv_e(tile, sat, ws, relay)
    plain = M.star_cycles(tile, sat, relay, 3, kv2, n2, ws, kv_e_ready=True).clone()
    M.prepare_kv_e(tile, sat, ws, relay, first_sat=True)
    assert ws.xi1 is not None
    cached = M.star_cycles(tile, sat, relay, 3, kv2, n2, ws, kv_e_ready=True).clone()
    torch.cuda.synchronize()
    assert torch.equal(plain, cached)
    # DSC_STAR_NO_FINAL_RELAY: satellite rows identical (row 31 then holds the relay node BEFORE the last update)
    short = M.star_cycles(tile, sat, relay, 3, kv2, n2, ws, kv_e_ready=True, relay_row=False)
    torch.cuda.synchronize()
    assert torch.equal(plain[:, :31], short[:, :31])
    M.prepare_kv_e(tile, sat, ws, relay)                 # and without the cached first half, 1 and 2 cycles
    for cyc in (1, 2):
        full = M.star_cycles(tile, sat, relay, cyc, kv2, n2, ws, kv_e_ready=True).clone()
        part = M.star_cycles(tile, sat, relay, cyc, kv2, n2, ws, kv_e_ready=True, relay_row=False)
        torch.cuda.synchronize()
        assert torch.equal(full[:, :31], part[:, :31])


@pytest.mark.parametrize("lq,lk,mode", [(31, 31, "pad"), (30, 30, "combined"), (1, 17, "ids"), (30, 31, "pad"), (7, 7, "none")])
def test_mha_attention_matches_oracle(L, dev, lq, lk, mode):
    g = torch.Generator().manual_seed(lq * 100 + lk)
    n = 5
    q, k, v = (torch.randn(n, l, 128, generator=g) for l in (lq, lk, lk))
    ids = torch.randint(0, 3, (n, lk), generator=g)
    ids[:, 0] = 1
    mask, kw = None, {}
    if mode == "pad":
        mask = O.create_padding_mask(ids)
        kw = dict(mask=mask.to(dev))
    elif mode == "combined":
        mask = torch.maximum(O.create_padding_mask(ids), O.create_look_ahead_mask(lk))
        kw = dict(mask=mask.to(dev))
    elif mode == "ids":      # newest query row of a causal prefix: all keys visible, PAD ids masked
        mask = O.create_padding_mask(ids)
        kw = dict(key_ids=ids.to(torch.int32).to(dev), causal=True, q_off=lk - 1)
    Q = q.reshape(n, lq, 8, 16).transpose(1, 2)
    K = k.reshape(n, lk, 8, 16).transpose(1, 2)
    V = v.reshape(n, lk, 8, 16).transpose(1, 2)
    lg = Q @ K.transpose(-1, -2) / 4.0
    if mask is not None:
        lg = lg + mask * -1e9
    ref = (torch.softmax(lg, -1) @ V).transpose(1, 2).reshape(n, lq, 128)
    out = torch.empty(n, lq, 128, device=dev)
    L.mha_attention(q.to(dev), k.to(dev), v.to(dev), out, **kw)
    assert rel_err(out, ref) < 3e-6


def test_channel_awgn_injected_noise_and_power_norm(L, dev):
    g = torch.Generator().manual_seed(4)
    U = 3
    u = torch.randn(U * 64, 31, 16, generator=g) * 0.3
    z = torch.randn(U * 64, 31, 16, generator=g)
    p = torch.randn(U * 64, 31, 16, generator=g)
    n_std = [O.snr_to_noise(s) for s in (0.0, 6.0, 18.0)]
    ref, refx = [], []
    for i in range(U):
        sl = slice(64 * i, 64 * (i + 1))
        xs = u[sl] / torch.sqrt(torch.mean(u[sl] * u[sl]))
        pu = p[sl] / torch.linalg.vector_norm(p[sl])
        refx.append(xs)
        ref.append(O.awgn(xs, pu, 3.0, n_std[i], z[sl]))
    ref, refx = torch.cat(ref), torch.cat(refx)
    ud = u.to(dev)
    sumsq = L.unit_sumsq(ud, U)
    psum = L.unit_sumsq(p.to(dev), U)
    # p/||p||_F = p / sqrt(1 * sumsq/elems) / sqrt(elems)  -> fold sqrt(size)*n_std*sqrt(PNR)/sqrt(elems) in p_scale
    pscale = torch.tensor([ns * np.sqrt(10 ** 0.3) for ns in n_std], dtype=torch.float32, device=dev)
    y, xn = L.channel(ud, U, torch.tensor(n_std, dtype=torch.float32, device=dev), x_sumsq=sumsq, noise=z.to(dev),
                      p=p.to(dev), p_sumsq=psum, p_factor=1.0, p_scale=pscale, want_x_norm=True)
    assert rel_err(xn, refx) < 2e-6
    assert rel_err(y, ref) < 1e-5
    assert rel_err(L.power_normalize(ud, U), refx) < 2e-6


@pytest.mark.parametrize("K,detector,apply", [(0, "MMSE", False), (1, "LS", True), (0, "MMSE", True)])
def test_channel_fading_matches_oracle(L, dev, K, detector, apply):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(128, 31, 16, generator=g)
    z = torch.randn(128, 31, 16, generator=g)
    hz = [(0.3, -1.1), (1.7, 0.2)]
    ref = torch.cat([O.fading(x[64 * i:64 * i + 64], K, 0.25, hz[i], z[64 * i:64 * i + 64], detector, apply) for i in range(2)])
    h = torch.tensor([[O.fading_coeff(K, *hz[i]).real, O.fading_coeff(K, *hz[i]).imag] for i in range(2)],
                     dtype=torch.float32, device=dev)
    y, _ = L.channel(x.to(dev), 2, torch.full((2,), 0.25, device=dev), noise=z.to(dev), h=h,
                     detector={"LS": 1, "MMSE": 2}[detector] if apply else 0)
    assert rel_err(y, ref) < 5e-6


def _philox_normal_reference(n4, seed, offset):
    """numpy restatement of the kernel's Philox4x32-10 + Box-Muller stream."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    i = np.arange(n4, dtype=np.uint64)
    c = [(i & 0xFFFFFFFF), (i >> np.uint64(32)), np.full(n4, offset & 0xFFFFFFFF, np.uint64), np.full(n4, offset >> 32, np.uint64)]
    k0, k1 = seed & 0xFFFFFFFF, seed >> 32
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & 0xFFFFFFFF, p1 >> np.uint64(32), p1 & 0xFFFFFFFF
        c = [(hi1 ^ c[1] ^ np.uint64(k0)) & 0xFFFFFFFF, lo1, (hi0 ^ c[3] ^ np.uint64(k1)) & 0xFFFFFFFF, lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    def bm(a, b):
        u1 = (a.astype(np.float64) + 1.0) * 2.0 ** -32
        u2 = b.astype(np.float64) * 2.0 ** -32
        r = np.sqrt(-2.0 * np.log(u1))
        return r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)
    z0, z1 = bm(c[0], c[1])
    z2, z3 = bm(c[2], c[3])
    return np.stack([z0, z1, z2, z3], 1).reshape(-1)


def test_channel_philox_stream(L, dev):
    x = torch.zeros(64, 31, 16, device=dev)
    y, _ = L.channel(x, 1, torch.ones(1, device=dev), seed=0x1234567890AB, offset=5)
    ref = _philox_normal_reference(64 * 31 * 4, 0x1234567890AB, 5)
    got = y.cpu().double().numpy().reshape(-1)
    assert np.abs(got - ref).max() < 2e-4
    assert abs(got.mean()) < 0.02 and abs(got.std() - 1.0) < 0.02
    y2, _ = L.channel(x, 1, torch.ones(1, device=dev), seed=0x1234567890AB, offset=6)
    assert not torch.equal(y, y2)


def test_vocab_argmax_and_ce_rows(L, dev):
    P = _cases.params("Transeiver_Star")
    g = torch.Generator().manual_seed(6)
    x = torch.randn(96, 128, generator=g)
    w, b = P["semantic_decoder/final_layer/kernel"], P["semantic_decoder/final_layer/bias"]
    logits = x.double() @ w.double() + b.double()
    V = w.shape[1]
    wd = torch.zeros(128, (V + 127) // 128 * 128, device=dev)
    wd[:, :V] = w.to(dev)
    ids = torch.zeros(96, 3, dtype=torch.int32, device=dev)
    L.vocab_argmax(x.to(dev), wd, b.to(dev), V, ids[:, 1], prec=0)
    top2 = logits.topk(2, -1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-5          # rows whose fp64 margin is not a numerical tie
    assert clear.float().mean() > 0.95
    assert torch.equal(ids[:, 1].cpu()[clear].long(), logits.argmax(-1)[clear])
    assert bool((ids[:, 0] == 0).all()) and bool((ids[:, 2] == 0).all())
    lg = torch.empty(96, V, device=dev)
    L.vocab_argmax(x.to(dev), wd, b.to(dev), V, ids[:, 2], logits=lg, prec=0)
    assert rel_err(lg, logits) < 2e-6
    tgt = torch.randint(0, V, (96,), generator=g)
    tgt[::5] = 0
    ref = (torch.logsumexp(logits, -1) - logits.gather(-1, tgt[:, None])[:, 0]) * (tgt != 0)
    assert rel_err(L.masked_ce_rows(lg, tgt.to(dev)), ref) < 5e-6
    # first-max tie rule of tf.argmax
    tie = torch.zeros(3, 1000, device=dev)
    tie[0, 17] = tie[0, 400] = 2.0
    tie[1, 999] = 1.0
    assert L.argmax_rows(tie).tolist() == [17, 999, 0]


@pytest.mark.parametrize("M", [96, 700])
@pytest.mark.parametrize("prec", [1, 2])
def test_vocab_argmax_tensor_core_fused(L, dev, M, prec):
    """dsc_vocab_argmax_tc: ids equal the fp64 argmax wherever the margin is not a numerical tie (bf16x3), exact
    ties resolve to the smallest index across tiles and across CTAs, untouched id columns stay untouched."""
    P = _cases.params("Transeiver_Star")
    g = torch.Generator().manual_seed(16)
    x = torch.randn(M, 128, generator=g)
    w, b = P["semantic_decoder/final_layer/kernel"].clone(), P["semantic_decoder/final_layer/bias"].clone()
    V = w.shape[1]
    # rows 0..2: three identical winning columns placed in different 64-column tiles / CTA ranges
    for r, cols in enumerate(([5, 900, 20000], [22233, 22200, 64], [12345, 12346, 12347])):
        for c in cols:
            w[:, c] = x[r] / x[r].norm() * 3.0
            b[c] = 0.25
    logits = x.double() @ w.double() + b.double()
    wd = torch.zeros(128, (V + 127) // 128 * 128, device=dev)
    wd[:, :V] = w.to(dev)
    ids = torch.full((M, 3), -7, dtype=torch.int32, device=dev)
    for _ in range(2):                                  # second call reuses the self-resetting workspace
        ids[:, 1] = -7
        L.vocab_argmax(x.to(dev), wd, b.to(dev), V, ids[:, 1], prec=prec)
    got = ids[:, 1].cpu().long()
    assert got[:3].tolist() == [5, 64, 12345]
    # the planted duplicate columns tie exactly (same operands, same arithmetic) for every row, so most rows test the
    # first-index rule; a row is "clear" when the gap to the best non-tied value is not a numerical tie
    mx = logits.max(-1).values
    second = logits.masked_fill(logits == mx[:, None], -float("inf")).max(-1).values
    clear = (mx - second) > (1e-5 if prec == 1 else 5e-2)
    assert clear.float().mean() > (0.95 if prec == 1 else 0.3)
    want = torch.from_numpy(np.argmax(logits.numpy(), axis=-1))          # first occurrence
    assert torch.equal(got[clear], want[clear])
    assert bool((ids[:, 0] == -7).all()) and bool((ids[:, 2] == -7).all())
    if prec == 1:                                       # same ids as the unfused tensor-core path (logits materialised)
        lg = torch.empty(M, V, device=dev)
        ids2 = torch.zeros(M, dtype=torch.int32, device=dev)
        L.vocab_argmax(x.to(dev), wd, b.to(dev), V, ids2, logits=lg, prec=1)
        assert torch.equal(ids2.cpu().long(), got)


def test_bleu_counts_bit_exact_on_real_sentences(L, dev):
    fix = json.load(open(os.path.join(_cases.GOLDEN_DIR, "europarl_sample.json")))
    from test_bleu_oracle import corrupt, pad
    rng = random.Random(11)
    ref = [pad(s) for s in fix["sentences"]]
    hyp = [corrupt(s, rng) for s in fix["sentences"]]
    hyp[0] = [2] + [5] * 30            # empty hypothesis
    hyp[1] = ref[1]                    # perfect
    hyp[2] = [1, ref[2][1], 2] + [0] * 28
    hyp[3] = [0] * 31                  # all PAD, no END
    want = B.bleu_counts(np.asarray(ref), np.asarray(hyp))
    got = L.bleu_counts(torch.tensor(ref, dtype=torch.int32, device=dev), torch.tensor(hyp, dtype=torch.int32, device=dev))
    assert np.array_equal(got.cpu().numpy(), want)


def test_fgm_normalize(L, dev):
    g = torch.Generator().manual_seed(8)
    grad = torch.randn(128, 31, 16, generator=g) * 1e-3
    ref = torch.cat([O.fgm_normalize(grad[:64], 1.0), O.fgm_normalize(grad[64:], 1.0)])
    got = L.fgm_normalize(grad.to(dev), 2, 1.0)
    assert rel_err(got, ref) < 5e-6
    assert abs(float(torch.linalg.vector_norm(got[:64])) - 1.0) < 1e-5


def test_bad_arguments_raise(L, dev):
    with pytest.raises(ValueError, match="multiple of 16"):
        L.linear(torch.zeros(4, 24, device=dev), torch.zeros(24, 16, device=dev), None)
    with pytest.raises(ValueError):
        L.bleu_counts(torch.zeros(2, 40, dtype=torch.int32, device=dev), torch.zeros(2, 40, dtype=torch.int32, device=dev))


@pytest.mark.parametrize("prec,tol", [(1, 3e-5), (2, 2e-2)])
@pytest.mark.parametrize("M,K,N,act", [(128, 128, 128, 0), (1984, 128, 384, 0), (300, 128, 256, 1), (64, 128, 22234, 0),
                                       (257, 512, 128, 1), (200, 256, 16, 0),
                                       # >= 296 output tiles: the 64 KB / three-CTAs-per-SM form of the kernel
                                       (40000, 128, 256, 1), (38001, 512, 128, 0), (40000, 256, 16, 0)])
def test_linear_tensor_core(L, dev, M, K, N, act, prec, tol):
    """tcgen05 path: bf16x3 split must be fp32-class (<= 3e-5 relative), single bf16 pass ~1e-2."""
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(K, N, generator=g) / np.sqrt(K)
    b = torch.randn(N, generator=g)
    ref = x.double() @ w.double() + b.double()
    if act:
        ref = torch.relu(ref)
    ldw = (N + 127) // 128 * 128
    wd = torch.zeros(K, ldw, device=dev)
    wd[:, :N] = w.to(dev)
    y = torch.full((M, N), 123.0, device=dev)
    L.linear(x.to(dev), wd, b.to(dev), act=act, n=N, prec=prec, out=y)
    torch.cuda.synchronize()
    assert rel_err(y, ref) < tol


def test_linear_tensor_core_row_skip_strided(L, dev):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(128, 128, generator=g).to(dev)
    w = torch.randn(128, 128, generator=g).to(dev)
    out = torch.full((128, 128), 7.0, device=dev)
    L.linear(x, w, None, out=out, row_mod=32, row_skip=31, prec=1)
    ref = (x.double() @ w.double()).float()
    keep = torch.arange(128, device=dev) % 32 != 31
    assert rel_err(out[keep], ref[keep]) < 3e-5 and bool((out[~keep] == 7.0).all())
    tile = torch.zeros((4, 32, 128), device=dev)
    L.linear(x[:4], w, None, out=tile[:, 31, :], prec=1)
    assert rel_err(tile[:, 31, :], ref[:4]) < 3e-5 and float(tile[:, :31].abs().max()) == 0.0


@pytest.mark.parametrize("prec,tol", [(1, 2e-4), (2, 1e-1)])
@pytest.mark.parametrize("S,n2", [(4, 0), (600, 17), (2368, 30)])
def test_star_cycles_at_bench_sizes_match_literal_oracle(L, dev, S, n2, prec, tol, monkeypatch):
    """dsc_star_cycles_tc at one tile, at a size that does not fill the SMs evenly and at bench.py's 592 tiles (4 per SM):
    two cycles against the literal tf.roll / concat / 5-key form of the oracle (modules.py:289-306), element-wise."""
    import deepsc_gan_b200.models.modules as M
    monkeypatch.setattr(M, "PREC", prec)
    P = _cases.params("Transeiver_Star", gain=3.0)
    pre = "semantic_decoder/dec_layers"
    g = torch.Generator().manual_seed(S)
    e = torch.randn(S, 31, 128, generator=g)
    h2 = torch.randn(S, 30, 128, generator=g)[:, :n2] if n2 else None
    h_ref, s_ref = O._star_cycles(P, pre, e, h2, 2, "multi_att_relay")
    sat, relay = M.sublayer1(128, 8).to(dev), M.sublayer1(128, 8).to(dev)
    with torch.no_grad():
        for mod, name in ((sat, "multi_att_satellite"), (relay, "multi_att_relay")):
            for w in ("wq", "wk", "wv"):
                getattr(mod, w).kernel.copy_(P[f"{pre}/{name}/{w}/kernel"])
            mod.dense.kernel.copy_(P[f"{pre}/{name}/dense/kernel"])
            mod.dense.bias.copy_(P[f"{pre}/{name}/dense/bias"])
    kv2 = None
    if n2:
        kv2 = torch.zeros(S, 30, 256, device=dev)
        kv2[:, :n2] = L.linear(h2.reshape(-1, 128).to(dev), relay._packed("kv"), None, prec=prec).view(S, n2, 256)
    x = M.star_cycles(L.star_pack(e.to(dev)), sat, relay, 2, kv2, n2)
    torch.cuda.synchronize()
    assert _cases.max_rel(x[:, :31], h_ref) < 4 * tol and _cases.max_rel(x[:, 31], s_ref) < 4 * tol


def test_star_cycles_ragged_batch_is_padded_to_whole_tiles(L, dev):
    """51 sentences (the last batch of the reference's 7,347-sentence test set, dataset/dataloader.py:14) are padded to 52
    for the tcgen05 kernel and trimmed again: same rows as the 52-sentence call."""
    import deepsc_gan_b200.models.modules as M
    torch.manual_seed(5)
    sat, relay = M.sublayer1(128, 8).to(dev), M.sublayer1(128, 8).to(dev)
    e = torch.randn(52, 31, 128, device=dev)
    e[51] = 0
    kv2 = torch.randn(52, 30, 256, device=dev)
    kv2[51] = 0
    full = M.star_cycles(L.star_pack(e), sat, relay, 3, kv2, 9).clone()
    part = M.star_cycles(L.star_pack(e[:51].contiguous()), sat, relay, 3, kv2[:51].contiguous(), 9)
    torch.cuda.synchronize()
    assert part.shape == (51, 32, 128) and torch.equal(part, full[:51])


def test_star_interleave_layout(L, dev):
    g = torch.Generator().manual_seed(0)
    for R, W in ((128, 128), (128, 256), (32, 256)):
        src = torch.randn(3, R, W, generator=g).to(dev)
        dst = L.star_interleave(src, torch.empty(src.numel(), device=dev), R)
        want = src.view(3, R, W // 4, 4).permute(0, 2, 1, 3).reshape(-1)
        assert torch.equal(dst, want)
    kv2i = torch.zeros(5 * 8192, device=dev)
    vals = torch.randn(5, 256, generator=g).to(dev)
    L.star_kv2_put(vals, kv2i, 7)
    assert torch.equal(kv2i.view(5, 64, 32, 4)[:, :, 7, :].reshape(5, 256), vals)
    assert float(kv2i.view(5, 64, 32, 4)[:, :, 8, :].abs().max()) == 0.0


@pytest.mark.parametrize("prec,tol", [(1, 1e-4), (2, 3e-2)])
def test_target_tail_fused(L, dev, prec, tol):
    """dsc_target_tail_tc = Dense + residual + LayerNorm + relay k|v projection + key-cache row write, against the
    same chain in fp64; ragged M (not a multiple of the 128-row tile)."""
    g = torch.Generator().manual_seed(77)
    Mrows = 200
    attn, resid = torch.randn(Mrows, 128, generator=g), torch.randn(Mrows, 128, generator=g)
    wo, bo = torch.randn(128, 128, generator=g) * 0.1, torch.randn(128, generator=g) * 0.1
    gamma, beta = torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g) * 0.1
    wkv = torch.randn(128, 256, generator=g) * 0.1
    h = resid.double() + attn.double() @ wo.double() + bo.double()
    h2 = (h - h.mean(-1, keepdim=True)) / torch.sqrt(h.var(-1, unbiased=False, keepdim=True) + 1e-6) * gamma.double() + beta.double()
    kv = h2 @ wkv.double()
    kv2i = torch.zeros(Mrows * 8192, device=dev)
    rows = torch.empty(Mrows, 256, device=dev)
    h2_out = torch.empty(Mrows, 128, device=dev)
    L.target_tail_tc(attn.to(dev), resid.to(dev), wo.to(dev), bo.to(dev), gamma.to(dev), beta.to(dev), wkv.to(dev), kv2i, 9,
                     prec, kv_rows=rows, h2_out=h2_out)
    torch.cuda.synchronize()
    assert rel_err(h2_out, h2) < tol and rel_err(rows, kv) < tol
    cache = kv2i.view(Mrows, 64, 32, 4)
    assert torch.equal(cache[:, :, 9, :].reshape(Mrows, 256), rows)
    assert float(cache[:, :, 8, :].abs().max()) == 0.0 and float(cache[:, :, 10, :].abs().max()) == 0.0


@pytest.mark.parametrize("N,act", [(256, 1), (512, 1), (384, 0), (100, 0)])
@pytest.mark.parametrize("prec,tol", [(1, 2e-5), (2, 1e-2)])
def test_persistent_k128_dense_matches_fp64_and_the_tiled_kernel(L, dev, N, act, prec, tol):
    """The persistent K = 128 Dense kernel (dsc_gemm_k128.cu: taken from two 128-row tiles per SM up, i.e. the channel
    codec's 73,408-row layers) against fp64, and against the tiled kernel (prec | 64; its K-block-major accumulation order
    differs, so agreement is to rounding, not bit for bit).  40,011 rows: ragged last tile, 3 tiles on some SMs, 2 on others."""
    M, K = 40011, 128
    g = torch.Generator().manual_seed(N + act)
    x = torch.randn(M, K, generator=g).to(dev)
    w = (torch.randn(K, (N + 3) // 4 * 4, generator=g) * 0.1).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    y = torch.full((M, N), float("nan"), device=dev)
    L.linear(x, w, b, act=act, n=N, prec=prec, out=y)
    y_tiled = torch.empty((M, N), device=dev)
    L.linear(x, w, b, act=act, n=N, prec=prec | 64, out=y_tiled)
    ref = x.double() @ w[:, :N].double() + b.double()
    if act:
        ref = ref.clamp_min(0)
    torch.cuda.synchronize()
    assert not torch.isnan(y).any()
    assert float((y.double() - ref).abs().max() / ref.abs().max()) < tol
    assert float((y - y_tiled).abs().max() / ref.abs().max()) < 2 * tol
