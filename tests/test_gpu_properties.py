"""Edge cases and size-independent properties at BASELINE.json's full bench size (-m gpu).

The oracle finishes a 64-sentence unit in seconds, not the 2,368-sentence super-batch of bench.py, so at full size
the CUDA path is checked through properties that do not need it: a super-batch equals its units run one by one (a
unit never sees its neighbours), runs are deterministic, graph replay equals eager, the channel is linear in the
perturbation, BLEU counts of a sentence against itself equal its totals, and a sample of units is still compared with
the oracle bit for bit.  Edge cases: empty batches, a ragged last unit count, sentences that fill all 31 positions,
sentences that are all padding."""
import numpy as np
import pytest
import torch

import _cases
from oracle import bleu_oracle as B, deepsc_oracle as O

pytestmark = pytest.mark.gpu
UNITS = 37          # bench.py default: 2,368 sentences = 592 tiles = 4 per SM


def build(kind, dev, prec=1):
    import deepsc_gan_b200.models as models
    import deepsc_gan_b200.models.modules as Mod
    from deepsc_gan_b200.utlis.parameters import para_config
    Mod.set_precision(prec)
    args = para_config([])
    net = getattr(models, kind)(args).to(dev).eval()
    net.load_tf_state_dict(_cases.params(kind))
    return args, net


def test_full_size_sweep_equals_unit_by_unit_and_oracle_sample(dev):
    from deepsc_gan_b200 import engine, sweep
    from deepsc_gan_b200.dataset.synthetic import synthetic_units
    kind = "Transeiver_Star"
    args, net = build(kind, dev)
    inp = synthetic_units(100, UNITS)
    g = torch.Generator().manual_seed(321)
    z = torch.randn(UNITS * 64, 31, 16, generator=g)
    snrs = [float(u % 19) for u in range(UNITS)]
    n_std = torch.tensor([O.snr_to_noise(s) for s in snrs], dtype=torch.float32)
    dec = engine.make_decoder(net, UNITS * 64)
    ids = engine.greedy_units(net, inp.to(dev), UNITS, n_std.to(dev), noise=z.to(dev), decoder=dec).clone()
    # determinism
    ids2 = engine.greedy_units(net, inp.to(dev), UNITS, n_std.to(dev), noise=z.to(dev), decoder=dec).clone()
    assert torch.equal(ids, ids2)
    # a unit never sees its neighbours: three units run alone give the same ids
    one = engine.make_decoder(net, 64)
    for u in (0, 17, 36):
        sl = slice(64 * u, 64 * u + 64)
        alone = engine.greedy_units(net, inp[sl].to(dev), 1, n_std[u:u + 1].to(dev), noise=z[sl].to(dev), decoder=one)
        assert torch.equal(alone, ids[sl]), u
    # BLEU counts: the device table equals the host restatement on every sentence of the super-batch
    counts = sweep._lib.bleu_counts(inp.to(dev), ids)
    assert np.array_equal(counts.cpu().numpy(), B.bleu_counts(inp.numpy(), ids.cpu().numpy()))
    # and eight of the 37 units against the oracle: every sentence equal, except where the first differing step is a
    # numerical tie in the oracle's own fp64 logits (random sentences; the strict no-excuse comparison is
    # test_gpu_models.py::test_strict_bit_exact_ids_on_margin_cases)
    P = _cases.params(kind)
    P64 = O.to_dtype(P, torch.float64)
    n_diff = 0
    for u in (0, 5, 9, 14, 19, 24, 30, 36):
        sl = slice(64 * u, 64 * u + 64)
        ref = O.greedy_decode_noattack(P, O.Spec(kind), inp[sl].long(), 0.0, "AWGN", float(n_std[u]), z[sl])
        got = ids[sl].cpu()
        bad = (got != ref).any(1).nonzero()[:, 0].tolist()
        n_diff += len(bad)
        if bad:
            _, lg = O.greedy_decode_noattack(P64, O.Spec(kind), inp[sl].long(), 0.0, "AWGN", float(n_std[u]), z[sl].double(),
                                             return_logits=True)
            for b in bad:
                t = int((got[b] != ref[b]).nonzero()[0]) - 1
                row = lg[b, t]
                gap = abs(float(row[int(got[b, t + 1])] - row[int(ref[b, t + 1])]))
                assert gap <= 1e-4 * float(row.abs().max()), (u, b, t, gap)
    assert n_diff <= 8, n_diff


def test_ragged_unit_counts_and_extreme_sentences(dev):
    """Unit counts that do not fill the SMs evenly (1, 3, 5 units = 16, 48, 80 tiles) and sentences at the extremes: all
    31 positions used without END, all padding, a single word."""
    from deepsc_gan_b200 import engine
    kind = "Transeiver_Star"
    args, net = build(kind, dev)
    P = _cases.params(kind)
    inp = _cases.synthetic_unit(7).clone()
    inp[0] = torch.randint(5, 22234, (31,), generator=torch.Generator().manual_seed(1)).to(torch.int32)   # no START/END, no PAD
    inp[0, 0] = 1
    inp[1] = 0                                                      # all padding
    inp[2] = 0
    inp[2, :3] = torch.tensor([1, 77, 2], dtype=torch.int32)        # one word
    z = _cases.draws()[0]
    n_std = float(O.snr_to_noise(9.0))
    ref = O.greedy_decode_noattack(P, O.Spec(kind), inp.long(), 0.0, "AWGN", n_std, z)
    for units in (1, 3, 5):
        big = torch.cat([inp] * units).to(dev)
        zz = torch.cat([z] * units).to(dev)
        ids = engine.greedy_units(net, big, units, torch.full((units,), n_std, device=dev), noise=zz).cpu()
        for u in range(units):
            got = ids[64 * u:64 * u + 64]
            assert (got == ref).all(1).float().mean() >= 63 / 64, (units, u)
            assert torch.equal(got[:3], ref[:3])                    # the three extreme sentences exactly
    counts = B.bleu_counts(inp.numpy(), ref.numpy())
    assert counts[1, 8] >= 0 and counts[1, 9] == 0                  # all-PAD reference has length 0


def test_empty_batches_are_no_ops(dev):
    from deepsc_gan_b200 import _lib as L
    f = dict(device=dev, dtype=torch.float32)
    assert L.linear(torch.empty((0, 128), **f), torch.zeros((128, 128), **f), None).shape == (0, 128)
    assert L.bleu_counts(torch.empty((0, 31), dtype=torch.int32, device=dev),
                         torch.empty((0, 31), dtype=torch.int32, device=dev)).shape == (0, 10)
    lib = L.load()
    # n_sent = 0 returns before any pointer is dereferenced (the fake pointers only have to pass the alignment checks)
    assert lib.dsc_star_cycles_tc(16, 16, 16, 16, None, 0, 128, 128, 128, 128, 128, 16, 16, 16, 0, 8, 1, None) == 0
    assert lib.dsc_embed(4, 31, 16, 22234, 16, 16, 128, 0, 31, 0, None) == 0
    assert lib.dsc_channel(16, None, 1.0, None, 0, 0, None, None, 1.0, None, None, 16, 0, 16, None, 0, 64, None) == 0


def test_channel_is_linear_in_the_perturbation_and_philox_is_stream_stable(dev):
    """y(x, p1 + p2) - y(x, 0) = (y(x, p1) - y(x, 0)) + (y(x, p2) - y(x, 0)) with injected noise; the Philox stream of
    a (seed, offset) pair does not depend on the batch size (element e always draws the same number)."""
    from deepsc_gan_b200 import _lib as L
    g = torch.Generator().manual_seed(4)
    S = UNITS * 64
    x = torch.randn(S, 31, 16, generator=g).to(dev)
    z = torch.randn(S, 31, 16, generator=g).to(dev)
    p1 = torch.randn(S, 31, 16, generator=g).to(dev)
    p2 = torch.randn(S, 31, 16, generator=g).to(dev)
    ns = torch.rand(UNITS, generator=g).to(dev)
    ps = torch.rand(UNITS, generator=g).to(dev)
    ss = L.unit_sumsq(x, UNITS)
    y = lambda p: L.channel(x, UNITS, ns, x_sumsq=ss, noise=z, p=p, p_scale=ps)[0]
    y0 = y(None)
    lhs = y((p1 + p2).contiguous()) - y0
    rhs = (y(p1) - y0) + (y(p2) - y0)
    assert float((lhs - rhs).abs().max()) < 1e-4 * float(lhs.abs().max())
    a, _ = L.channel(x, UNITS, ns, seed=99, offset=3)
    b, _ = L.channel(x[:640].contiguous(), 10, ns[:10].contiguous(), seed=99, offset=3)
    assert torch.equal(a[:640], b)


def test_bleu_self_counts_equal_totals_at_full_size(dev):
    from deepsc_gan_b200 import _lib as L
    from deepsc_gan_b200.dataset.synthetic import synthetic_units
    inp = synthetic_units(0, UNITS).to(dev)
    c = L.bleu_counts(inp, inp).cpu().numpy()
    assert np.array_equal(c[:, 8], c[:, 9])                         # hyp_len == ref_len
    for n in range(4):                                              # every n-gram of a sentence matches itself
        expect = np.maximum(c[:, 8] - n, 0)
        assert np.array_equal(c[:, n], expect), n
        assert np.array_equal(c[:, 4 + n], np.maximum(expect, 1)), n
