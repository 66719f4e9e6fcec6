"""Model-level parity (-m gpu): the nn.Module mirror of models/ and utlis/ against the CPU oracle and the
frozen goldens, through the C ABI.  Token ids and BLEU counts must be bit-exact (asserted with torch.equal on the
margin-enforced cases of tests/golden/make_margin_cases.py; on the plain seeded cases a differing sentence must be
explained by a numerical tie in the oracle's own fp64 logits); symbols and logits within 1e-3 relative.
Every test runs in the package's default arithmetic (prec 1: tcgen05 bf16x3) unless it says otherwise."""
import numpy as np
import pytest
import torch

import _cases
from oracle import bleu_oracle as B, deepsc_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-3


def rel_err(a, b):
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def build(kind, dev, params=None):
    import deepsc_gan_b200.models as models
    from deepsc_gan_b200.utlis.parameters import para_config
    args = para_config([])
    net = getattr(models, kind)(args).to(dev).eval()
    net.load_tf_state_dict(_cases.params(kind) if params is None else params)
    return args, net


def explained_mismatches(kind, channel, ids_gpu, ids_ref, inp):
    """Rows whose first differing step is a numerical tie in the oracle's fp64 logits are reported, not
    failed: returns (n_mismatching_sentences, n_unexplained)."""
    bad = (ids_gpu != ids_ref).any(1).nonzero()[:, 0].tolist()
    if not bad:
        return 0, 0
    spec = O.Spec(kind)
    P64 = O.to_dtype(_cases.params(kind), torch.float64)
    z, _, _, h_z, _ = _cases.draws()
    n_std = O.snr_to_noise(_cases.SNR_DB)
    _, lg = O.greedy_decode_noattack(P64, spec, inp.long(), 0.0, channel, n_std, z.double(), h_z, return_logits=True)
    unexplained = 0
    for b in bad:
        t = int((ids_gpu[b] != ids_ref[b]).nonzero()[0]) - 1
        row = lg[b, t]
        gap = abs(float(row[int(ids_gpu[b, t + 1])] - row[int(ids_ref[b, t + 1])]))
        if gap > 1e-4 * float(row.abs().max()):
            unexplained += 1
    return len(bad), unexplained


@pytest.mark.parametrize("prec", [1, 0])
@pytest.mark.parametrize("kind,channel", [("Transeiver_Star", "AWGN"), ("Transeiver_Star", "Rayleigh"),
                                          ("Transeiver", "AWGN"), ("Transeiver_star", "AWGN"), ("Transeiver_GAN", "AWGN")])
def test_forward_and_greedy_against_goldens(dev, kind, channel, prec, monkeypatch):
    """Teacher-forced forward (symbols, received symbols, logits, loss) and greedy decode of every system against the
    frozen oracle outputs, in the default tcgen05 arithmetic (prec 1) and in the fp32 debug mode (prec 0)."""
    import deepsc_gan_b200.models.modules as M0
    monkeypatch.setattr(M0, "PREC", prec)
    from deepsc_gan_b200.models.modules import create_masks, loss_function
    from deepsc_gan_b200.utlis.eval import greedy_decode_noattack
    from deepsc_gan_b200.utlis.tools import BleuScore
    gold = np.load(_cases.golden_path(kind, channel))
    args, net = build(kind, dev)
    inp = torch.from_numpy(gold["inp"]).to(dev)
    z, z_r, p, h_z, h_z_r = _cases.draws()
    n_std = O.snr_to_noise(_cases.SNR_DB)
    tar_inp = inp[:, :-1]
    masks = create_masks(inp, tar_inp)
    with torch.no_grad():
        if kind == "Transeiver_GAN":
            pred, pred_r, x, y = net(inp, tar_inp, p.to(dev), 3.0, channel, n_std, False, *masks, traingan=True,
                                     noise=z.to(dev), h=h_z, noise_r=z_r.to(dev), h_r=h_z_r)
            assert rel_err(torch.logsumexp(pred_r, -1), gold["pred_r_lse"]) < RTOL
        else:
            pred, x, y, y_again = net(inp, tar_inp, p.to(dev), 3.0, channel, n_std, False, *masks, noise=z.to(dev), h=h_z)
            assert y_again is y
    assert tuple(pred.shape) == (64, gold["lse"].shape[1], 22234)
    assert rel_err(x[:8], gold["symbols"]) < RTOL and _cases.max_rel(x[:8], gold["symbols"]) < RTOL
    assert rel_err(y[:8], gold["received"]) < RTOL and _cases.max_rel(y[:8], gold["received"]) < RTOL
    assert rel_err(pred[:4, :, :64], gold["logits_slice"]) < RTOL
    assert _cases.max_rel(pred[:4, :, :64], gold["logits_slice"]) < RTOL
    assert rel_err(torch.logsumexp(pred, -1), gold["lse"]) < RTOL
    tar_real = inp if kind in ("Transeiver_star", "Transeiver_Star") else inp[:, 1:]
    assert abs(float(loss_function(tar_real, pred)) - float(gold["loss"])) < RTOL * float(gold["loss"])
    from deepsc_gan_b200 import _lib
    agree = (_lib.argmax_rows(pred).cpu().numpy() == gold["tf_argmax"]).mean()
    assert agree > 0.995, f"teacher-forced argmax agreement {agree}"

    ids = greedy_decode_noattack(args, inp, net, 0.0, channel, n_std, noise=z.to(dev), h=h_z)
    assert ids.dtype == torch.int32 and tuple(ids.shape) == (64, 31)
    ids_c = ids.cpu()
    ids_ref = torch.from_numpy(gold["greedy_ids"])
    n_bad, unexplained = explained_mismatches(kind, channel, ids_c, ids_ref, inp.cpu())
    assert unexplained == 0, f"{unexplained} sentences differ from the oracle beyond a numerical tie"
    assert n_bad <= (3 if prec else 1), f"{n_bad} sentences hit ties; expected at most {3 if prec else 1} in this seeded case"
    counts = BleuScore.counts_from_ids(inp, ids).cpu().numpy()
    same = (ids_c == ids_ref).all(1).numpy()
    assert np.array_equal(counts[same], gold["bleu_counts"][same])
    assert np.array_equal(counts, B.bleu_counts(gold["inp"], ids_c.numpy()))


def _margin_run(dev, name):
    """Greedy decode of one margin case through the product API -> (ids, counts, fixture)."""
    from deepsc_gan_b200 import sweep
    from deepsc_gan_b200.utlis.eval import greedy_decode_noattack
    from deepsc_gan_b200.utlis.tools import BleuScore
    kind, channel, _, snr_db, psr_db, _ = _cases.MARGIN_CASES[name]
    fx = np.load(_cases.margin_path(name))
    args, net = build(kind, dev, _cases.margin_params(name))
    inp = torch.from_numpy(fx["inp"]).to(dev)
    z = _cases.margin_noise(fx["seeds"]).to(dev)
    n_std = O.snr_to_noise(snr_db)
    if psr_db is None:
        ids = greedy_decode_noattack(args, inp, net, 0.0, channel, n_std, noise=z, h=_cases.draws()[3])
        counts = BleuScore.counts_from_ids(inp, ids)
    else:                                               # configs[3]: the generator's perturbation at a fixed PSR
        runner = sweep.SweepRunner(net, 1, channel="AWGN", attack="generator", psr_db=psr_db)
        ids, counts = runner.run(inp, torch.full((1,), n_std, device=dev), noise=z)
    return ids.cpu(), counts.cpu().numpy(), fx


@pytest.mark.parametrize("name", list(_cases.MARGIN_CASES))
def test_strict_bit_exact_ids_on_margin_cases(dev, name):
    """STRICT: all 64 x 31 greedy ids and all 64 x 10 BLEU counts equal the oracle's, in the default arithmetic (prec 1),
    for every system and every BASELINE.json eval config (baseline AWGN, star AWGN at 0 / 6 / 12 / 18 dB, star Rayleigh,
    4-layer star, generator attack at fixed PSR), on the seeded Keras initialisation and on the "lively" weight variant
    whose decoded ids depend on the input.  The cases hold only sentences whose fp64 top-2 logit margin is >= 2e-3 of the
    logit scale at every step (tests/test_margin_cases.py re-derives that on the CPU), so no tie can excuse a difference."""
    import deepsc_gan_b200.models.modules as M0
    assert M0.PREC == 1
    ids, counts, fx = _margin_run(dev, name)
    assert ids.dtype == torch.int32 and torch.equal(ids, torch.from_numpy(fx["ids"]))
    assert np.array_equal(counts, fx["counts"])


def test_strict_bit_exact_ids_two_units_two_snrs_one_launch(dev):
    """The two star AWGN margin cases (6 dB and 0 dB) stacked as two units of one launch: per-unit noise std and power
    normalisation, ids equal to the per-unit oracle runs."""
    from deepsc_gan_b200 import engine
    names = ("Transeiver_Star_AWGN", "Transeiver_Star_AWGN_0dB")
    fxs = [np.load(_cases.margin_path(n)) for n in names]
    args, net = build("Transeiver_Star", dev)
    inp = torch.cat([torch.from_numpy(f["inp"]) for f in fxs]).to(dev)
    z = torch.cat([_cases.margin_noise(f["seeds"]) for f in fxs]).to(dev)
    n_std = torch.tensor([O.snr_to_noise(_cases.MARGIN_CASES[n][3]) for n in names], dtype=torch.float32, device=dev)
    ids = engine.greedy_units(net, inp, 2, n_std, noise=z).cpu()
    assert torch.equal(ids, torch.cat([torch.from_numpy(f["ids"]) for f in fxs]))


def test_multi_unit_greedy_equals_per_unit_oracle(dev):
    """Three units at three SNR points in one set of launches == the oracle run unit by unit."""
    from deepsc_gan_b200 import engine
    from deepsc_gan_b200.dataset.synthetic import synthetic_units
    kind = "Transeiver_Star"
    args, net = build(kind, dev)
    P = _cases.params(kind)
    inp = synthetic_units(3, 3)
    g = torch.Generator().manual_seed(99)
    z = torch.randn(192, 31, 16, generator=g)
    snrs = [0.0, 9.0, 18.0]
    n_std = torch.tensor([O.snr_to_noise(s) for s in snrs], dtype=torch.float32)
    ids = engine.greedy_units(net, inp.to(dev), 3, n_std.to(dev), noise=z.to(dev)).cpu()
    for u in range(3):
        sl = slice(64 * u, 64 * u + 64)
        ref = O.greedy_decode_noattack(P, O.Spec(kind), inp[sl].long(), 0.0, "AWGN", float(n_std[u]), z[sl])
        assert (ids[sl] == ref).all(1).float().mean() >= 63 / 64


def test_submodule_calls_used_by_eval(dev):
    """utlis/eval.py reaches inside the model (semantic_encoder.call, channel_encoder.call,
    channel_layer.fading, channel_decoder.call, semantic_decoder.call): same attribute names and results."""
    from deepsc_gan_b200.models.modules import create_padding_mask, create_look_ahead_mask
    kind = "Transeiver_Star"
    args, net = build(kind, dev)
    P = _cases.params(kind)
    spec = O.Spec(kind)
    inp = _cases.synthetic_unit(2)
    z = _cases.draws()[0]
    sem = net.semantic_encoder.call(inp.to(dev), False, create_padding_mask(inp.to(dev)))
    assert rel_err(sem, O.semantic_encoder(P, spec, inp.long(), None)) < 1e-4
    x = net.channel_encoder.call(sem)
    x_ref = O.channel_encoder(P, O.semantic_encoder(P, spec, inp.long(), None))
    assert rel_err(x, x_ref) < 1e-4
    y = net.channel_layer.fading(x, None, 0, 1, 0.2, noise=z.to(dev), h=(0.4, -0.6))
    assert rel_err(y, O.fading(x_ref, 1, 0.2, (0.4, -0.6), z)) < 1e-4
    with pytest.raises(ValueError, match="detector must in LS and MMSE"):
        net.channel_layer.fading(x, None, 0, 1, 0.2, "ZF")
    mem = net.channel_decoder.call(y)
    mem_ref = O.channel_decoder(P, O.fading(x_ref, 1, 0.2, (0.4, -0.6), z))
    assert rel_err(mem, mem_ref) < 1e-4
    prefix = inp[:, :9].to(dev)
    comb = torch.maximum(create_padding_mask(prefix), create_look_ahead_mask(9, device=dev))
    lg = net.semantic_decoder.call(prefix, mem, False, comb, None)
    lg_ref = O.semantic_decoder(P, spec, inp[:, :9].long(), mem_ref, torch.maximum(O.create_padding_mask(inp[:, :9]), O.create_look_ahead_mask(9)), None)
    assert tuple(lg.shape) == (64, 31, 22234) and rel_err(lg, lg_ref) < RTOL


def test_training_mode_raises(dev):
    args, net = build("Transeiver_Star", dev)
    inp = _cases.synthetic_unit(0).to(dev)
    with pytest.raises(RuntimeError, match="differentiable"):      # dropout + backward live in the differentiable mode
        net.semantic_encoder.call(inp, True, None)


@pytest.mark.parametrize("kind", ["Transeiver", "Transeiver_Star"])
def test_graph_replay_decoder_equals_eager(dev, kind):
    """engine.GraphedDecoder (one CUDA-graph launch per batch) returns the ids of the eager loop, also on replay with
    new inputs."""
    from deepsc_gan_b200 import engine
    args, net = build(kind, dev)
    n_std = torch.full((2,), float(O.snr_to_noise(6.0)), device=dev)
    eager = engine.make_decoder(net, 128, graph=False)
    graphed = engine.make_decoder(net, 128, graph=True)
    assert isinstance(graphed, engine.GraphedDecoder)
    g = torch.Generator().manual_seed(5)
    for first in (0, 2, 4):
        inp = _cases.synthetic_unit(first)
        inp = torch.cat([inp, _cases.synthetic_unit(first + 1)]).to(dev)
        z = torch.randn(128, 31, 16, generator=g).to(dev)
        a = engine.greedy_units(net, inp, 2, n_std, noise=z, decoder=eager).clone()
        b = engine.greedy_units(net, inp, 2, n_std, noise=z, decoder=graphed).clone()
        assert torch.equal(a, b)
    assert graphed.launches > 100
