"""Backward kernels (K17), FGM evaluators and training steps against the CPU oracle differentiated by autograd
(-m gpu).  The oracle is the fp32 PyTorch restatement of the reference's forward; its autograd gradient plays the role
of tf.GradientTape.  Gradients must agree within 1e-3 relative (max-norm per tensor)."""
import math

import numpy as np
import pytest
import torch

import _cases
from oracle import deepsc_oracle as O

pytestmark = pytest.mark.gpu
GTOL = 1e-3


def rel_err(a, b):
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def rel_l2(a, b):
    """||a - b||_2 / ||b||_2: the whole-model gradient metric (a ReLU gate that flips on an fp32 rounding difference
    moves single elements by more than 1e-3 of the max-norm without changing the gradient as a vector)."""
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    return float(torch.linalg.vector_norm(a - b) / (torch.linalg.vector_norm(b) + 1e-30))


def build(kind, dev, dropout=None):
    import deepsc_gan_b200.models as models
    from deepsc_gan_b200.utlis.parameters import para_config
    args = para_config([])
    if dropout is not None:
        args.encoder_dropout = args.decoder_dropout = dropout
    net = getattr(models, kind)(args).to(dev).eval()
    net.load_tf_state_dict(_cases.params(kind))
    if dropout is not None:
        # the star layers take their Dropout rate from the constructor default 0.1, not from the flags
        # (models/modules.py:653 builds STE without a rate): switch it off explicitly for oracle comparisons
        for m in net.modules():
            for attr in ("drop_pro", "dropout_pro"):
                if hasattr(m, attr):
                    setattr(m, attr, dropout)
    return args, net


@pytest.fixture(scope="module")
def AG(dev):
    from deepsc_gan_b200 import autograd
    return autograd


# ----------------------------------------------------------------------------------------------- single kernels
@pytest.mark.parametrize("M,N,K,ta,tb", [(70, 50, 33, 0, 0), (128, 128, 4000, 1, 0), (100, 130, 257, 0, 1),
                                         (65, 64, 1000, 1, 1), (16, 22234, 128, 0, 0), (128, 40, 22234, 0, 1)])
def test_gemm_transposes_and_split_k(AG, dev, M, N, K, ta, tb):
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn((K, M) if ta else (M, K), generator=g)
    b = torch.randn((N, K) if tb else (K, N), generator=g)
    ref = (a.t() if ta else a).double() @ (b.t() if tb else b).double()
    out = AG.gemm(a.to(dev), b.to(dev), bool(ta), bool(tb))
    assert rel_err(out, ref) < 2e-5
    acc = torch.ones((M, N), device=dev)
    AG.gemm(a.to(dev), b.to(dev), bool(ta), bool(tb), out=acc, accumulate=True)
    assert rel_err(acc, ref + 1.0) < 2e-5


@pytest.mark.parametrize("M,N,K,ta,tb", [(1920, 128, 22234, 0, 1), (128, 22234, 1920, 1, 0), (300, 200, 20000, 0, 1),
                                          (130, 9000, 1000, 1, 0)])
def test_gemm_vocabulary_sized_on_tensor_cores(AG, dev, M, N, K, ta, tb):
    """The vocabulary-sized backward products (dX = dY @ W^T, dW = X^T @ dY) take the bf16x3 tcgen05 path
    (dsc_gemm_nt_tc, directly or after dsc_transpose): fp32-class against an fp64 product, with and without accumulate."""
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn((K, M) if ta else (M, K), generator=g)
    b = torch.randn((N, K) if tb else (K, N), generator=g)
    ref = (a.t() if ta else a).double() @ (b.t() if tb else b).double()
    out = AG.gemm(a.to(dev), b.to(dev), bool(ta), bool(tb))
    assert rel_err(out, ref) < 2e-5
    acc = torch.ones((M, N), device=dev)
    AG.gemm(a.to(dev), b.to(dev), bool(ta), bool(tb), out=acc, accumulate=True)
    assert rel_err(acc, ref + 1.0) < 2e-5


def test_linear_layernorm_gradients(AG, dev):
    from deepsc_gan_b200.models.modules import LayerNormalization
    g = torch.Generator().manual_seed(3)
    x = torch.randn(200, 128, generator=g)
    w = torch.randn(128, 70, generator=g) * 0.1
    b = torch.randn(70, generator=g)
    up = torch.randn(200, 70, generator=g)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    (torch.relu(xr @ wr + br) * up).sum().backward()
    xd, wd, bd = (t.to(dev).requires_grad_(True) for t in (x, w, b))
    (AG.linear(xd, wd, bd, 1, 0) * up.to(dev)).sum().backward()
    for got, ref in ((xd.grad, xr.grad), (wd.grad, wr.grad), (bd.grad, br.grad)):
        assert rel_err(got, ref) < 1e-5
    # residual + LN, single and doubled
    res = torch.randn(200, 128, generator=g)
    up2 = torch.randn(200, 128, generator=g)
    ga, ba, gb, bb = (torch.randn(128, generator=g) * 0.3 + 1 for _ in range(4))
    for double in (False, True):
        leaves = [t.clone().requires_grad_(True) for t in (x, res, ga, ba, gb, bb)]
        o = torch.nn.functional.layer_norm(leaves[0] + leaves[1], (128,), leaves[2], leaves[3], 1e-6)
        if double:
            o = torch.nn.functional.layer_norm(o + o, (128,), leaves[4], leaves[5], 1e-6)
        (o * up2).sum().backward()
        la, lb = LayerNormalization().to(dev), LayerNormalization().to(dev)
        with torch.no_grad():
            la.gamma.copy_(ga); la.beta.copy_(ba); lb.gamma.copy_(gb); lb.beta.copy_(bb)
        xd, rd = x.to(dev).requires_grad_(True), res.to(dev).requires_grad_(True)
        (AG.add_layernorm(xd, rd, la, lb if double else None) * up2.to(dev)).sum().backward()
        assert rel_err(xd.grad, leaves[0].grad) < 1e-4 and rel_err(rd.grad, leaves[1].grad) < 1e-4
        assert rel_err(la.gamma.grad, leaves[2].grad) < 1e-4 and rel_err(la.beta.grad, leaves[3].grad) < 1e-4
        if double:
            assert rel_err(lb.gamma.grad, leaves[4].grad) < 1e-4 and rel_err(lb.beta.grad, leaves[5].grad) < 1e-4


@pytest.mark.parametrize("lq,lk,mode", [(30, 30, "causal+pad"), (30, 31, "pad"), (31, 31, "none"), (1, 62, "none")])
def test_mha_attention_gradients(AG, dev, lq, lk, mode):
    n = 5
    g = torch.Generator().manual_seed(lq * 100 + lk)
    q, k, v = torch.randn(n, lq, 128, generator=g), torch.randn(n, lk, 128, generator=g), torch.randn(n, lk, 128, generator=g)
    up = torch.randn(n, lq, 128, generator=g)
    mask = None
    if "pad" in mode:
        ids = torch.randint(1, 50, (n, lk), generator=g)
        ids[:, lk - 4:] = 0
        mask = O.create_padding_mask(ids)
        if "causal" in mode:
            mask = torch.maximum(mask, O.create_look_ahead_mask(lq))
    leaves = [t.clone().requires_grad_(True) for t in (q, k, v)]
    split = lambda t: t.view(n, -1, 8, 16).transpose(1, 2)
    s = split(leaves[0]) @ split(leaves[1]).transpose(-1, -2) / 4.0
    if mask is not None:
        s = s + mask * -1e9
    o = (torch.softmax(s, -1) @ split(leaves[2])).transpose(1, 2).reshape(n, lq, 128)
    (o * up).sum().backward()
    qd, kd, vd = (t.to(dev).requires_grad_(True) for t in (q, k, v))
    od = AG.MhaAttention.apply(qd, kd, vd, None if mask is None else mask.to(dev), None, False, 0)
    assert rel_err(od, o) < 1e-5
    (od * up.to(dev)).sum().backward()
    for got, ref in ((qd.grad, leaves[0].grad), (kd.grad, leaves[1].grad), (vd.grad, leaves[2].grad)):
        assert rel_err(got, ref) < 1e-4


@pytest.mark.parametrize("n2", [0, 30])
def test_star_cycle_gradients(AG, dev, n2):
    """One STE/STD-style cycle loop (2 cycles) through the differentiable path vs autograd on the oracle's literal
    roll/concat formulation: gradients w.r.t. the input tile e, the h2 keys and every weight."""
    import deepsc_gan_b200.models.modules as Mod
    S = 6
    g = torch.Generator().manual_seed(40 + n2)
    P = {}
    for name in ("x/multi_att_satellite", "x/multi_att_relay"):
        for w in ("wq", "wk", "wv"):
            P[f"{name}/{w}/kernel"] = torch.randn(128, 128, generator=g) * 0.08
        P[f"{name}/dense/kernel"] = torch.randn(128, 128, generator=g) * 0.08
        P[f"{name}/dense/bias"] = torch.randn(128, generator=g) * 0.1
    e = torch.randn(S, 31, 128, generator=g)
    h2 = torch.randn(S, 30, 128, generator=g) if n2 else None
    up = torch.randn(S, 32, 128, generator=g)
    Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    er = e.clone().requires_grad_(True)
    h2r = None if h2 is None else h2.clone().requires_grad_(True)
    h, s = O._star_cycles(Pr, "x", er, h2r, 2, "multi_att_relay")       # literal roll/concat 5-key form
    (torch.cat([h, s[:, None, :]], 1) * up).sum().backward()

    sat, rel = Mod.sublayer1(128, 8).to(dev), Mod.sublayer1(128, 8).to(dev)
    with torch.no_grad():
        for mod, name in ((sat, "x/multi_att_satellite"), (rel, "x/multi_att_relay")):
            mod.wq.kernel.copy_(P[f"{name}/wq/kernel"]); mod.wk.kernel.copy_(P[f"{name}/wk/kernel"])
            mod.wv.kernel.copy_(P[f"{name}/wv/kernel"]); mod.dense.kernel.copy_(P[f"{name}/dense/kernel"])
            mod.dense.bias.copy_(P[f"{name}/dense/bias"])
    ed = e.to(dev).requires_grad_(True)
    h2d = None if h2 is None else h2.to(dev).requires_grad_(True)
    with Mod.differentiable():
        tile = AG.StarPack.apply(ed)
        kv2 = None if h2d is None else rel.project(h2d.reshape(-1, 128), "kv").view(S, 30, 256)
        x = Mod.star_cycles(tile, sat, rel, 2, kv2, n2)
        (x * up.to(dev)).sum().backward()
    assert rel_err(x[:, :31], h) < 1e-4 and rel_err(x[:, 31], s) < 1e-4
    assert rel_err(ed.grad, er.grad) < GTOL
    if h2 is not None:
        assert rel_err(h2d.grad, h2r.grad) < GTOL
    for mod, name in ((sat, "x/multi_att_satellite"), (rel, "x/multi_att_relay")):
        for attr in ("wq", "wk", "wv"):
            assert rel_err(getattr(mod, attr).kernel.grad, Pr[f"{name}/{attr}/kernel"].grad) < GTOL, (name, attr)
        assert rel_err(mod.dense.kernel.grad, Pr[f"{name}/dense/kernel"].grad) < GTOL
        assert rel_err(mod.dense.bias.grad, Pr[f"{name}/dense/bias"].grad) < GTOL


def test_channel_powernorm_ce_embed_gradients(AG, dev):
    g = torch.Generator().manual_seed(77)
    n_units, S = 2, 128
    x = torch.randn(S, 31, 16, generator=g)
    p = torch.randn(S, 31, 16, generator=g)
    z = torch.randn(S, 31, 16, generator=g)
    up = torch.randn(S, 31, 16, generator=g)
    n_std = torch.tensor([0.3, 0.7])
    ps = torch.tensor([1.5, 0.25])
    elems = 64 * 31 * 16

    def norm(t, factor):
        t3 = t.reshape(n_units, -1)
        return (t3 / torch.sqrt(factor * (t3 * t3).mean(1, keepdim=True))).reshape(t.shape)

    # AWGN with power-normalised symbols and perturbation
    xr, pr = x.clone().requires_grad_(True), p.clone().requires_grad_(True)
    sc = ps.repeat_interleave(64)[:, None, None]
    y = norm(xr, 1.0) + n_std.repeat_interleave(64)[:, None, None] * z + sc * norm(pr, 2.0)
    (y * up).sum().backward()
    xd, pd = x.to(dev).requires_grad_(True), p.to(dev).requires_grad_(True)
    yd = AG.Channel.apply(AG.PowerNormalize.apply(xd, n_units, 1.0), AG.PowerNormalize.apply(pd, n_units, 2.0), n_units,
                          n_std.to(dev), z.to(dev), 0, 0, ps.to(dev), None, 0)
    assert rel_err(yd, y) < 1e-5
    (yd * up.to(dev)).sum().backward()
    assert rel_err(xd.grad, xr.grad) < 1e-4 and rel_err(pd.grad, pr.grad) < 1e-4
    # fading with each detector
    hh = torch.tensor([[0.6, -0.4], [-0.2, 0.9]])
    for det in (0, 1, 2):
        xr = x.clone().requires_grad_(True)
        xc = torch.view_as_complex(xr.reshape(S, 248, 2))
        hc = torch.complex(hh[:, 0], hh[:, 1]).repeat_interleave(64)[:, None]
        nc = torch.view_as_complex((n_std.repeat_interleave(64)[:, None, None] * z).reshape(S, 248, 2).contiguous())
        yc = xc * hc + nc
        if det:
            den = (hc * hc.conj()).real + (2 * (n_std.repeat_interleave(64)[:, None] ** 2) if det == 2 else 0.0)
            yc = yc * hc.conj() / den
        yr = torch.view_as_real(yc).reshape(S, 31, 16)
        (yr * up).sum().backward()
        xd = x.to(dev).requires_grad_(True)
        yd = AG.Channel.apply(xd, None, n_units, n_std.to(dev), z.to(dev), 0, 0, None, hh.to(dev), det)
        assert rel_err(yd, yr) < 1e-5
        (yd * up.to(dev)).sum().backward()
        assert rel_err(xd.grad, xr.grad) < 1e-4, det
    # masked CE rows
    lg = torch.randn(40, 1000, generator=g)
    tgt = torch.randint(0, 1000, (40,), generator=g)
    tgt[::4] = 0
    wrow = torch.randn(40, generator=g)
    lr_ = lg.clone().requires_grad_(True)
    ce = (torch.logsumexp(lr_, -1) - lr_.gather(-1, tgt[:, None])[:, 0]) * (tgt != 0)
    (ce * wrow).sum().backward()
    ld = lg.to(dev).requires_grad_(True)
    (AG.MaskedCeRows.apply(ld, tgt.to(dev)) * wrow.to(dev)).sum().backward()
    assert rel_err(ld.grad, lr_.grad) < 1e-5
    # embedding (duplicate ids accumulate) and star pack
    table = torch.randn(300, 128, generator=g)
    ids = torch.randint(0, 300, (9, 31), generator=g)
    ids[:, 5] = 7
    upe = torch.randn(9, 31, 128, generator=g)
    tr = table.clone().requires_grad_(True)
    ((tr[ids] * math.sqrt(128.0)) * upe).sum().backward()
    td = table.to(dev).requires_grad_(True)
    pos = torch.zeros(512, 128, device=dev)
    (AG.Embed.apply(ids.to(dev).int(), td, pos, 0) * upe.to(dev)).sum().backward()
    assert rel_err(td.grad, tr.grad) < 1e-5
    src = torch.randn(9, 31, 128, generator=g)
    upt = torch.randn(9, 32, 128, generator=g)
    sr = src.clone().requires_grad_(True)
    (torch.cat([sr, sr.mean(1, keepdim=True)], 1) * upt).sum().backward()
    sd = src.to(dev).requires_grad_(True)
    (AG.StarPack.apply(sd) * upt.to(dev)).sum().backward()
    assert rel_err(sd.grad, sr.grad) < 1e-5


def test_dropout_mask_and_backward(AG, dev):
    x = torch.ones(1 << 16, device=dev).requires_grad_(True)
    y = AG.Dropout.apply(x, 0.1, 1234, 5)
    kept = (y != 0).float().mean().item()
    assert abs(kept - 0.9) < 0.01 and torch.allclose(y[y != 0], torch.full_like(y[y != 0], 1 / 0.9))
    y.sum().backward()
    assert torch.equal(x.grad, y.detach())                     # same mask, same scale
    y2 = AG.Dropout.apply(x, 0.1, 1234, 6)
    assert not torch.equal(y2, y)


def test_adam_kernel_matches_keras_formula(AG, dev):
    g = torch.Generator().manual_seed(5)
    p, gr, g2 = torch.randn(1000, generator=g), torch.randn(1000, generator=g), torch.randn(1000, generator=g)
    m, v = torch.zeros(1000), torch.zeros(1000)
    pd, md, vd = p.to(dev), m.to(dev), v.to(dev)
    lr, b1, b2, eps = 5e-4, 0.9, 0.98, 1e-8
    pr = p.double().clone()
    mr, vr = m.double(), v.double()
    for step in (1, 2, 3):
        gg = 0.5 * gr.double() + 0.25 * g2.double()
        mr = b1 * mr + (1 - b1) * gg
        vr = b2 * vr + (1 - b2) * gg * gg
        pr = pr - lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step) * mr / (vr.sqrt() + eps)
        AG.adam_step(pd, gr.to(dev), md, vd, lr, step, b1, b2, eps, 0.5, g2.to(dev), 0.25)
    assert rel_err(pd, pr) < 1e-6


# ----------------------------------------------------------------------------------------------- whole models
def oracle_param_grads(kind, inp, z, z_r=None, p=None, PNR_dB=3.0, traingan=False, which="r"):
    spec = O.Spec(kind)
    P = {k: v.clone().requires_grad_(True) for k, v in _cases.params(kind).items()}
    tar_inp = inp[:, :-1]
    masks = O.create_masks(inp, tar_inp)
    n_std = O.snr_to_noise(_cases.SNR_DB)
    pp = torch.zeros(64, 31, 16) if p is None else p
    fw = O.transceiver_forward(P, spec, inp, tar_inp, pp, PNR_dB, "AWGN", n_std, *masks, z=z, z_r=z_r, traingan=traingan)
    tar_real = inp if spec.is_star else inp[:, 1:]
    if kind == "Transeiver_GAN":
        ce_p, ce_r = O.loss_function(tar_real, fw[0]), O.loss_function(tar_real, fw[1])
        return P, ce_p, ce_r
    loss = O.loss_function(tar_real, fw[0])
    return P, loss, None


@pytest.mark.parametrize("prec,gtol", [(1, 5e-3), (0, GTOL)])
@pytest.mark.parametrize("kind", ["Transeiver_Star", "Transeiver", "Transeiver_star"])
def test_parameter_gradients_match_oracle_autograd(dev, kind, prec, gtol):
    """d CE / d(every parameter) of a teacher-forced AWGN forward with a perturbation, whole model: in the default
    arithmetic (prec 1; a bf16x3 product carries ~2^-17 relative error, ten times fp32's, and the star recurrences
    amplify it: 5e-3 of the gradient's L2 norm) and with fp32 products (prec 0: 1e-3)."""
    import deepsc_gan_b200.models.modules as Mod
    Mod.set_precision(prec)
    args, net = build(kind, dev)
    inp = _cases.synthetic_unit(1)
    z, _, p, _, _ = _cases.draws()
    P, loss_ref, _ = oracle_param_grads(kind, inp.long(), z, p=p)
    loss_ref.backward()
    tar_inp = inp[:, :-1].to(dev)
    masks = Mod.create_masks(inp.to(dev), tar_inp)
    n_std = O.snr_to_noise(_cases.SNR_DB)
    with Mod.differentiable():
        outs = net(inp.to(dev), tar_inp, p.to(dev), 3.0, channel="AWGN", n_std=n_std, training=False,
                   enc_padding_mask=masks[0], combined_mask=masks[1], dec_padding_mask=masks[2], noise=z.to(dev))
        tar_real = inp.to(dev) if O.Spec(kind).is_star else inp[:, 1:].to(dev)
        loss = Mod.loss_function(tar_real, outs[0])
        loss.backward()
    assert rel_err(loss, loss_ref) < 1e-5
    worst, worst_max = [], []
    for name, prm in net.named_parameters():
        ref = P[name.replace(".", "/")].grad
        assert prm.grad is not None, name
        worst.append((rel_l2(prm.grad, ref), name))
        worst_max.append((rel_err(prm.grad, ref), name))
    worst.sort(reverse=True)
    worst_max.sort(reverse=True)
    assert worst[0][0] < gtol, worst[:5]
    assert worst_max[0][0] < 10 * gtol, worst_max[:5]


def test_fgm_eval_steps_match_oracle(dev):
    """eval_step_star / eval_step_normal: losses, predictions and the FGM perturbation (hence d loss / d symbols)."""
    from deepsc_gan_b200.utlis import eval as E
    z, z2, _, h_z, _ = _cases.draws()
    n_std = O.snr_to_noise(_cases.SNR_DB)
    for kind, fn, channel in (("Transeiver_Star", E.eval_step_star, "AWGN"), ("Transeiver", E.eval_step_normal, "AWGN"),
                              ("Transeiver_Star", E.eval_step_star, "Rayleigh")):
        args, net = build(kind, dev)
        inp = _cases.synthetic_unit(2)
        ref = O.eval_step(_cases.params(kind), O.Spec(kind), inp.long(), inp.long(), 3.0, channel, n_std, z, z2, h_z)
        got = fn(inp.to(dev), inp.to(dev), net, 3.0, channel=channel, n_std=n_std, epsilon=1, noise=z.to(dev),
                 noise2=z2.to(dev), h=h_z)
        assert rel_err(got[0], ref[0]) < 1e-4 and rel_err(got[1], ref[1]) < 1e-3, (kind, channel)
        assert rel_err(got[2], ref[2]) < 1e-3 and rel_err(got[3], ref[3]) < 1e-3
        # the perturbation itself, recomputed from the same symbol gradient
        tar_inp = inp[:, :-1].to(dev)
        masks = E.create_masks(inp.to(dev), tar_inp)
        tar_real = inp.to(dev) if O.Spec(kind).is_star else inp[:, 1:].to(dev)
        _, _, g, _ = E._symbol_gradient(net, inp.to(dev), tar_inp, tar_real, 3.0, channel, n_std, masks, z.to(dev), h_z)
        pert = E.fgm_perturbation(g, 1)
        assert rel_err(pert, ref[4]) < 2e-3, (kind, channel)
        assert abs(float(torch.linalg.vector_norm(pert)) - 1.0) < 1e-4
        assert all(p.requires_grad for p in net.parameters())         # _frozen restores the flags


def test_fgm_attacked_greedy_decoders_match_oracle(dev):
    """greedy_decode (utlis/eval.py:11-75) on a star and on the baseline system and greedy_decode_gan (:120-187): the
    scaled FGM perturbation (hence d loss / d received symbols), the symbols, the realised noise, the teacher-forced
    clean-branch argmax ``noa`` and the greedy ids against the oracle's restatement of the same functions.  The oracle
    runs in fp64 here: the per-sample normalisation of the direction is ill-conditioned for a sentence whose gradient is
    small, and the fp32 oracle itself is 1.7e-2 away from the fp64 one on the Rayleigh case below."""
    from deepsc_gan_b200.utlis import eval as E
    z, z2, _, h_z, _ = _cases.draws()
    n_std = O.snr_to_noise(_cases.SNR_DB)
    inp = _cases.synthetic_unit(3)
    P64 = lambda kind: O.to_dtype(_cases.params(kind), torch.float64)
    for kind in ("Transeiver_Star", "Transeiver"):
        args, net = build(kind, dev)
        outputs, scaled, noise, x = E.greedy_decode(args, inp.to(dev), net, 6.0, channel="AWGN", n_std=n_std, epsilon=1,
                                                    noise=z.to(dev), noise2=z2.to(dev))
        ref_ids, ref_scaled, ref_x = O.greedy_decode(P64(kind), O.Spec(kind), inp.long(), 6.0, "AWGN", n_std, z.double(),
                                                     z2.double())
        assert tuple(outputs.shape) == (64, 31) and outputs.dtype == torch.int32
        assert rel_err(x, ref_x) < 1e-3 and rel_err(noise, n_std * z2) < 1e-3
        assert rel_err(scaled, ref_scaled) < 2e-3, kind
        assert (outputs.cpu() == ref_ids).all(1).float().mean() >= 62 / 64, kind
    # Rayleigh: the fading branch ignores the perturbation (:52-55) but still returns it
    args, net = build("Transeiver_Star", dev)
    outputs, scaled, noise, x = E.greedy_decode(args, inp.to(dev), net, 6.0, channel="Rayleigh", n_std=n_std, noise=z.to(dev),
                                                noise2=z2.to(dev), h=h_z)
    ref_ids, ref_scaled, _ = O.greedy_decode(P64("Transeiver_Star"), O.Spec("Transeiver_Star"), inp.long(), 6.0, "Rayleigh",
                                             n_std, z.double(), z2.double(), h_z)
    assert noise is None and rel_err(scaled, ref_scaled) < 2e-3
    assert (outputs.cpu() == ref_ids).all(1).float().mean() >= 62 / 64
    # GAN model
    args, gan = build("Transeiver_GAN", dev)
    spec = O.Spec("Transeiver_GAN")
    out = E.greedy_decode_gan(args, inp.to(dev), gan, 6.0, channel="AWGN", n_std=n_std, noise=z.to(dev), noise2=z2.to(dev))
    ref = O.greedy_decode_gan(P64("Transeiver_GAN"), spec, inp.long(), 6.0, "AWGN", n_std, z.double(), z2.double())
    assert len(out) == 5 and tuple(out[0].shape) == (64, 31) and tuple(out[1].shape) == (64, 30)
    assert (out[0].cpu() == ref[0]).all(1).float().mean() >= 62 / 64
    assert (out[1].cpu() == ref[1]).float().mean() > 0.995                  # noa: teacher-forced argmax of the clean branch
    assert rel_err(out[2], ref[2]) < 2e-3 and rel_err(out[4], ref[3]) < 1e-3


@pytest.mark.parametrize("channel", ["AWGN", "Rayleigh"])
def test_eval_step_FGM_matches_oracle(dev, channel):
    """utlis/eval.py:367-408 on ``Transeiver_GAN``: clean-branch loss, direction w.r.t. y_r (AWGN) or w.r.t. the channel
    symbols of an extra AWGN forward (fading), attacked loss on the perturbed branch, both prediction tensors."""
    from deepsc_gan_b200.utlis import eval as E
    z, z2, _, h_z, _ = _cases.draws()
    n_std = O.snr_to_noise(_cases.SNR_DB)
    inp = _cases.synthetic_unit(3)
    args, gan = build("Transeiver_GAN", dev)
    ref = O.eval_step_FGM(_cases.params("Transeiver_GAN"), O.Spec("Transeiver_GAN"), inp.long(), inp.long(), 6.0, channel,
                          n_std, z, z2, z, h_z)
    got = E.eval_step_FGM(inp.to(dev), inp.to(dev), gan, 6.0, channel=channel, n_std=n_std, noise=z.to(dev),
                          noise2=z2.to(dev), noise2_r=z.to(dev), h=h_z)
    assert len(got) == 4
    assert rel_err(got[0], ref[0]) < 1e-4 and rel_err(got[1], ref[1]) < 1e-3
    assert rel_err(got[2], ref[2]) < 1e-3 and rel_err(got[3], ref[3]) < 1e-3
    assert all(p.requires_grad for p in gan.parameters())


@pytest.mark.parametrize("channel", ["AWGN", "Rayleigh"])
def test_pgd_bisection_matches_oracle(dev, channel, capsys):
    """utlis/eval.py:235-318: the device-side bisection takes the same ten decisions as the reference's host loop:
    same epsilon (printed, :312), same losses and predictions of the first and the tenth forward."""
    from deepsc_gan_b200.utlis import eval as E
    z, z2, p_extra, h_z, _ = _cases.draws()
    g = torch.Generator().manual_seed(77)
    zs = [torch.randn(64, 31, 16, generator=g) for _ in range(10)]
    n_std = O.snr_to_noise(_cases.SNR_DB)
    inp = _cases.synthetic_unit(3)
    args, base = build("Transeiver", dev)
    ref = O.eval_step_normal_pgd(_cases.params("Transeiver"), O.Spec("Transeiver"), inp.long(), inp.long(), 6.0, channel,
                                 n_std, z, zs, h_z)
    got = E.eval_step_normal_pgd(inp.to(dev), inp.to(dev), base, 6.0, channel=channel, n_std=n_std, noise=z.to(dev),
                                 noises=[t.to(dev) for t in zs], h=h_z, verbose=True)
    printed = capsys.readouterr().out
    assert len(got) == 4 and tuple(got[3].shape) == (64, 30, 22234)
    assert abs(float(printed.split("epsilon=")[1].split()[0]) - ref[4]) < 1e-9, (printed, ref[4])
    assert rel_err(got[0], ref[0]) < 1e-4 and rel_err(got[1], ref[1]) < 1e-3
    assert rel_err(got[2], ref[2]) < 1e-3 and rel_err(got[3], ref[3]) < 1e-3


def test_eval_noise_advances_between_calls(dev):
    """Without an injected tensor every greedy call draws fresh channel noise (tf.random.normal per call,
    utlis/eval.py:90-93): the Philox offset advances with the channel layer's call counter."""
    from deepsc_gan_b200.utlis import eval as E
    args, net = build("Transeiver_Star", dev)
    net.load_tf_state_dict(_cases.params("Transeiver_Star", gain=3.0, emb_gain=4.0))
    inp = _cases.synthetic_unit(3).to(dev)
    c0 = net.channel_layer._calls
    a = E.greedy_decode_noattack(args, inp, net, 0.0, "AWGN", O.snr_to_noise(0.0))
    b = E.greedy_decode_noattack(args, inp, net, 0.0, "AWGN", O.snr_to_noise(0.0))
    assert net.channel_layer._calls == c0 + 2 and not torch.equal(a, b)
    c = E.greedy_decode_noattack(args, inp, net, 0.0, "AWGN", O.snr_to_noise(0.0), seed=5)
    net.channel_layer._calls -= 1
    d = E.greedy_decode_noattack(args, inp, net, 0.0, "AWGN", O.snr_to_noise(0.0), seed=5)
    assert torch.equal(c, d)                                          # same (seed, offset): same stream


def _keras_adam(p, g, m, v, step, lr=5e-4, b1=0.9, b2=0.98, eps=1e-8):
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    return p - lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step) * m / (v.sqrt() + eps), m, v


def test_train_step_noattack_updates_like_oracle_plus_adam(dev):
    from deepsc_gan_b200.utlis import trainer as T
    kind = "Transeiver_Star"
    args, net = build(kind, dev, dropout=0.0)
    opt = T.make_optimizer(net)
    inp = _cases.synthetic_unit(4)
    z = _cases.draws()[0]
    n_std = O.snr_to_noise(3.0)
    spec = O.Spec(kind)
    P = {k: v.clone().requires_grad_(True) for k, v in _cases.params(kind).items()}
    masks = O.create_masks(inp.long(), inp[:, :-1].long())
    fw = O.transceiver_forward(P, spec, inp.long(), inp[:, :-1].long(), torch.zeros(64, 31, 16), 0, "AWGN", n_std, *masks, z=z)
    # train_step_noattack uses tar[:, 1:] for every model (utlis/trainer.py:14); the star decoder emits 31 positions,
    # so the reference step is only shape-consistent for the baseline: use the star-consistent target on both sides
    loss_ref = O.loss_function(inp.long(), fw[0])
    loss_ref.backward()
    # our step, star-consistent target through train_attack_step's convention is tested below; here call the pieces
    import deepsc_gan_b200.models.modules as Mod
    fp = opt.fp
    fp.grad_bucket[0].zero_(); fp.point_grads(0)
    with Mod.differentiable():
        m = Mod.create_masks(inp.to(dev), inp[:, :-1].to(dev))
        outs = net(inp.to(dev), inp[:, :-1].to(dev), None, 0, channel="AWGN", n_std=n_std, training=True,
                   enc_padding_mask=m[0], combined_mask=m[1], dec_padding_mask=m[2], noise=z.to(dev))
        loss = Mod.loss_function(inp.to(dev), outs[0])
        loss.backward()
    opt.apply(fp.ranges(lambda n: True), fp.grad_bucket[0], 1.0)
    assert rel_err(loss, loss_ref) < 1e-4
    # The first Adam step is -lr * g / (|g| + eps'): essentially -lr * sign(g).  Compare element-wise where the oracle
    # gradient is not small against the tensor's scale (an fp32 rounding difference can flip the sign of a tiny
    # element, and a flipped ReLU gate moves isolated elements): at most 0.1% of those elements may disagree.
    n_big = n_bad = 0
    for name, prm in net.named_parameters():
        ref = P[name.replace(".", "/")]
        want, _, _ = _keras_adam(ref.detach(), ref.grad, torch.zeros_like(ref), torch.zeros_like(ref), 1)
        step_ref = (want - ref.detach())
        step_got = prm.detach().cpu() - ref.detach()
        big = ref.grad.abs() > 5e-2 * ref.grad.abs().max()
        n_big += int(big.sum())
        n_bad += int(((step_got - step_ref).abs()[big] > 0.02 * 5e-4).sum())
    assert n_big > 100000 and n_bad <= 1e-3 * n_big, (n_bad, n_big)


def test_baseline_train_step_and_gan_train_step(dev):
    """train_step_noattack on the baseline model (its shapes are consistent in the reference) and gan_train_step:
    losses against the oracle, parameter updates against oracle-autograd + the Keras Adam formula applied in the
    reference's order (A: all but g on CE_r; B: g on 10-CE_p; C: receiver on l*CE_r + (1-l)*CE_p)."""
    from deepsc_gan_b200.utlis import gan_train as GT, trainer as T
    n_std = O.snr_to_noise(3.0)
    z, z_r, p_draw, _, _ = _cases.draws()
    inp = _cases.synthetic_unit(5)
    # ---- baseline, plain step
    args, net = build("Transeiver", dev, dropout=0.0)
    opt = T.make_optimizer(net)
    P = {k: v.clone().requires_grad_(True) for k, v in _cases.params("Transeiver").items()}
    masks = O.create_masks(inp.long(), inp[:, :-1].long())
    fw = O.transceiver_forward(P, O.Spec("Transeiver"), inp.long(), inp[:, :-1].long(), torch.zeros(64, 31, 16), 0, "AWGN",
                               n_std, *masks, z=z)
    loss_ref = O.loss_function(inp[:, 1:].long(), fw[0])
    loss = T.train_step_noattack(inp.to(dev), inp.to(dev), None, net, opt, channel="AWGN", n_std=n_std, noise=z.to(dev))
    assert rel_err(loss, loss_ref) < 1e-4 and opt.iterations == 1
    # ---- GAN step with the generator in the loop
    args, gan = build("Transeiver_GAN", dev, dropout=0.0)
    opt = GT.make_optimizer(gan)
    lam = 0.5
    pn = p_draw / torch.linalg.vector_norm(p_draw)
    P = {k: v.clone().requires_grad_(True) for k, v in _cases.params("Transeiver_GAN").items()}
    fw = O.transceiver_forward(P, O.Spec("Transeiver_GAN"), inp.long(), inp[:, :-1].long(), pn, 40, "AWGN", n_std, *masks,
                               z=z, z_r=z_r, traingan=True)
    ce_p, ce_r = O.loss_function(inp[:, 1:].long(), fw[0]), O.loss_function(inp[:, 1:].long(), fw[1])
    names = list(P)
    g_r = dict(zip(names, torch.autograd.grad(ce_r, [P[n] for n in names], retain_graph=True, allow_unused=True)))
    g_p = dict(zip(names, torch.autograd.grad(ce_p, [P[n] for n in names], allow_unused=True)))
    loss, g_loss, d_loss = GT.gan_train_step(inp.to(dev), inp.to(dev), None, gan, opt, lam, channel="AWGN", n_std=n_std,
                                             training=True, traingan=True, noise=z.to(dev), noise_r=z_r.to(dev),
                                             p_draw=p_draw.to(dev))
    assert rel_err(loss, ce_r) < 1e-4 and rel_err(g_loss, 10 - ce_p) < 1e-4
    assert rel_err(d_loss, lam * ce_r + (1 - lam) * ce_p) < 1e-4 and opt.iterations == 3
    n_big = n_bad = 0
    for name, prm in gan.named_parameters():
        key = name.replace(".", "/")
        ref = P[key].detach()
        is_g = key.startswith("generator/")
        is_rx = key.startswith("channel_decoder/") or key.startswith("semantic_decoder/")
        val, m, v = ref.clone(), torch.zeros_like(ref), torch.zeros_like(ref)
        small = torch.zeros_like(ref, dtype=torch.bool)
        if not is_g:
            g = g_r[key] if g_r[key] is not None else torch.zeros_like(ref)
            small |= g.abs() < 5e-2 * g.abs().max()
            val, m, v = _keras_adam(val, g, m, v, 1)
        if is_g:
            small |= g_p[key].abs() < 5e-2 * g_p[key].abs().max()
            val, m, v = _keras_adam(val, -g_p[key], m, v, 2)
        if is_rx:
            g = lam * g_r[key] + (1 - lam) * g_p[key]
            small |= g.abs() < 5e-2 * g.abs().max()
            val, m, v = _keras_adam(val, g, m, v, 3)
        ok = ~small
        n_big += int(ok.sum())
        n_bad += int((((prm.detach().cpu() - ref) - (val - ref)).abs()[ok] > 0.05 * 5e-4).sum())
    assert n_big > 100000 and n_bad <= 2e-3 * n_big, (n_bad, n_big)


def test_train_attack_step_matches_oracle(dev):
    """utlis/trainer.py:30-64 (FGM adversarial training) with dropout off: both losses and the Adam update of every
    parameter against oracle autograd + the Keras Adam formula."""
    from deepsc_gan_b200.utlis import trainer as T
    kind = "Transeiver_Star"
    args, net = build(kind, dev, dropout=0.0)
    opt = T.make_optimizer(net)
    inp = _cases.synthetic_unit(6)
    z, z2 = _cases.draws()[0], _cases.draws()[1]
    n_std = O.snr_to_noise(3.0)
    P = {k: v.clone().requires_grad_(True) for k, v in _cases.params(kind).items()}
    loss_ref, loss_m_ref, grads = O.train_attack_step(P, O.Spec(kind), inp.long(), inp.long(), 3.0, "AWGN", n_std, z, z2)
    loss, loss_m = T.train_attack_step(inp.to(dev), inp.to(dev), None, 3.0, net, opt, channel="AWGN", n_std=n_std,
                                       noise=z.to(dev), noise2=z2.to(dev))
    assert rel_err(loss, loss_ref) < 1e-4 and rel_err(loss_m, loss_m_ref) < 1e-3 and opt.iterations == 1
    n_big = n_bad = 0
    for name, prm in net.named_parameters():
        ref, g = P[name.replace(".", "/")].detach(), grads[name.replace(".", "/")]
        want, _, _ = _keras_adam(ref, g, torch.zeros_like(ref), torch.zeros_like(ref), 1)
        big = g.abs() > 5e-2 * g.abs().max()
        n_big += int(big.sum())
        n_bad += int((((prm.detach().cpu() - ref) - (want - ref)).abs()[big] > 0.05 * 5e-4).sum())
    assert n_big > 100000 and n_bad <= 2e-3 * n_big, (n_bad, n_big)


def test_gan_train_step_literal_layer_names_also_update_the_channel_encoder(dev):
    """``step_c_literal_names=True``: Keras names a Channel_Encoder instance 'channel__encoder', so the reference's
    freeze-by-name (utlis/gan_train.py:41) misses it and step C moves the channel encoder too; the default (intent)
    leaves it to step A alone."""
    from deepsc_gan_b200.utlis import gan_train as GT
    n_std = O.snr_to_noise(3.0)
    z, z_r, p_draw, _, _ = _cases.draws()
    inp = _cases.synthetic_unit(5).to(dev)
    res = {}
    for literal in (False, True):
        args, gan = build("Transeiver_GAN", dev, dropout=0.0)
        opt = GT.make_optimizer(gan)
        GT.gan_train_step(inp, inp, None, gan, opt, 0.5, channel="AWGN", n_std=n_std, training=True, traingan=True,
                          noise=z.to(dev), noise_r=z_r.to(dev), p_draw=p_draw.to(dev), step_c_literal_names=literal)
        res[literal] = {n: p.detach().clone() for n, p in gan.named_parameters()}
    for n in res[False]:
        # (two runs differ in isolated elements anyway: atomics in the embedding / split-K accumulations)
        moved = ((res[False][n] - res[True][n]).abs() > 1e-4).float().mean().item()
        assert (moved > 0.5) if n.startswith("channel_encoder.") else (moved < 0.01), (n, moved)


def test_training_with_dropout_runs(dev):
    """training=True applies dropout at the reference's sites (star layers use their constructor default 0.1): the loss
    changes with the dropout seed and stays finite."""
    import deepsc_gan_b200.models.modules as Mod
    args, net = build("Transeiver_Star", dev)
    inp = _cases.synthetic_unit(6).to(dev)
    z = _cases.draws()[0].to(dev)
    n_std = O.snr_to_noise(3.0)
    losses = []
    for seed in (1, 2):
        Mod.set_dropout_seed(seed)
        with Mod.differentiable(), torch.no_grad():
            m = Mod.create_masks(inp, inp[:, :-1])
            outs = net(inp, inp[:, :-1], None, 0, channel="AWGN", n_std=n_std, training=True, enc_padding_mask=m[0],
                       combined_mask=m[1], dec_padding_mask=m[2], noise=z)
            losses.append(float(Mod.loss_function(inp, outs[0])))
    assert all(math.isfinite(v) for v in losses) and abs(losses[0] - losses[1]) > 1e-6


def test_graph_replayed_gan_train_step(dev):
    """GraphedGanTrainStep: one CUDA-graph launch per step; Adam's bias correction and the dropout masks follow the
    device-side step counter; training on a repeated batch lowers the clean-branch loss; optimizer iterations advance by
    three per step as in the reference (three apply_gradients calls)."""
    import deepsc_gan_b200.models.modules as Mod
    from deepsc_gan_b200.utlis import gan_train as GT
    args, gan = build("Transeiver_GAN", dev)
    gan.train()
    opt = GT.make_optimizer(gan)
    inp = _cases.synthetic_unit(8).to(dev)
    step = GT.GraphedGanTrainStep(gan, opt, 0.5, n_std=O.snr_to_noise(3.0), traingan=True, warmup=2)
    losses = []
    for i in range(8):
        before = opt.fp.flat.clone()
        loss, g_loss, d_loss = step(inp, inp)
        losses.append(float(loss))
        assert math.isfinite(losses[-1]) and math.isfinite(float(g_loss)) and math.isfinite(float(d_loss))
        moved = (opt.fp.flat - before).abs().max()
        assert 1e-6 < float(moved) < 5e-3, (i, float(moved))
    assert opt.iterations == 3 * step.steps and step.steps == 10 and int(step.step_dev) == 8
    assert losses[-1] < losses[0]
    # the eager path still works afterwards and sees the updated weights (cache epoch bumped after every replay)
    Mod.set_dropout_seed(1)
    out = GT.gan_train_step(inp, inp, None, gan, opt, 0.5, channel="AWGN", n_std=O.snr_to_noise(3.0), training=True, traingan=True)
    assert math.isfinite(float(out[0])) and float(out[0]) < losses[0] + 0.5
