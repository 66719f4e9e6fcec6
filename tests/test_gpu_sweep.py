"""SURVEY.md 8(f1) on hardware (-m gpu): the SNR-sweep driver ``sweep.evaluate_sweep`` with a real ``SweepRunner`` against the
oracle run item by item, and the real-data path (the reference's own test set, tests/golden/europarl_test.npz)."""
import math

import numpy as np
import pytest
import torch

import _cases
from oracle import bleu_oracle as B, deepsc_oracle as O

pytestmark = pytest.mark.gpu


def build(kind, dev):
    import deepsc_gan_b200.models as models
    from deepsc_gan_b200.utlis.parameters import para_config
    args = para_config([])
    net = getattr(models, kind)(args).to(dev).eval()
    net.load_tf_state_dict(_cases.params(kind))
    return args, net


@pytest.mark.parametrize("channel", ["AWGN", "Rayleigh"])
def test_evaluate_sweep_equals_oracle_item_by_item(dev, channel):
    """2 units x 3 SNR points = 6 work items through evaluate_sweep (3 launches of 2 items; the items of one launch sit at
    different SNR points): the int32 count table equals the oracle's bit for bit and the result rows
    [snr_idx, BLEU-1, BLEU-4] equal the rows formed from the oracle's counts.  Margin-enforced case (make_margin_cases.py)."""
    from deepsc_gan_b200 import sweep
    fx = np.load(_cases.sweep_margin_path(channel))
    args, net = build("Transeiver_Star", dev)
    units = torch.from_numpy(fx["units"])
    n_items = len(_cases.SWEEP_SNRS) * _cases.SWEEP_UNITS
    h_z = _cases.draws()[3]
    h_all = torch.tensor([[math.sqrt(0.5) * h_z[0], math.sqrt(0.5) * h_z[1]]] * n_items, dtype=torch.float32)
    runner = sweep.SweepRunner(net, 2, channel=channel)
    rows, counts, snr_index = sweep.evaluate_sweep(runner, units, _cases.SWEEP_SNRS, channel=channel, K=0,
                                                   noise_for_item=lambda i: _cases.margin_noise(fx["seeds"][i]),
                                                   h_all=h_all if channel != "AWGN" else None)
    assert counts.shape == (n_items * 64, 10) and counts.dtype == torch.int32
    assert np.array_equal(counts.numpy(), fx["counts"])
    assert np.array_equal(snr_index, np.repeat([0, 0, 1, 1, 2, 2], 64))
    want = sweep.bleu_table(fx["counts"], snr_index, 3)
    assert rows == want and [r[0] for r in rows] == [0, 1, 2]
    # the oracle's float scores from the same counts (independent formula: Fractions, nltk method0)
    for s in range(3):
        sel = fx["counts"][snr_index == s]
        assert abs(rows[s][1] - B.bleu_scores(sel, (1, 0, 0, 0)).mean()) < 1e-12
        assert abs(rows[s][2] - B.bleu_scores(sel, (0.25, 0.25, 0.25, 0.25)).mean()) < 1e-12
    # a ragged launch size (U = 4 over 6 items: the tail launch is padded and trimmed) gives the same table
    runner4 = sweep.SweepRunner(net, 4, channel=channel)
    _, counts4, _ = sweep.evaluate_sweep(runner4, units, _cases.SWEEP_SNRS, channel=channel, K=0,
                                         noise_for_item=lambda i: _cases.margin_noise(fx["seeds"][i]),
                                         h_all=h_all if channel != "AWGN" else None)
    assert torch.equal(counts4, counts)


def test_real_test_set_units_through_the_sweep_runner(dev):
    """The reference's own sentences (first 3 units of data/txt/test_data.pkl, dataset/dataloader.py padding) at 3 SNR points:
    BLEU counts equal the host restatement on the decoded ids, a unit alone equals the unit inside a super-batch, and the
    ragged 51-sentence tail batch of the test set decodes (padded to whole tiles inside the star decoder)."""
    from deepsc_gan_b200 import engine, sweep
    from deepsc_gan_b200.utlis.eval import greedy_decode_noattack
    ids, vocab = _cases.europarl_test()
    args, net = build("Transeiver_Star", dev)
    units = torch.from_numpy(ids[:192])
    runner = sweep.SweepRunner(net, 3, channel="AWGN", seed=3)
    n_std = torch.tensor([sweep.snr_to_noise(s) for s in (0.0, 9.0, 18.0)], dtype=torch.float32, device=dev)
    g = torch.Generator().manual_seed(12)
    z = torch.randn(192, 31, 16, generator=g).to(dev)
    out, counts = runner.run(units.to(dev), n_std, noise=z)
    out = out.clone()
    assert np.array_equal(counts.cpu().numpy(), B.bleu_counts(ids[:192], out.cpu().numpy()))
    alone = engine.greedy_units(net, units[64:128].to(dev), 1, n_std[1:2], noise=z[64:128])
    assert torch.equal(alone, out[64:128])
    tail = torch.from_numpy(ids[7296:]).to(dev)                    # 51 sentences: 7,347 = 114 * 64 + 51
    assert tail.shape[0] == 51
    got = greedy_decode_noattack(args, tail, net, 0.0, "AWGN", 0.2, noise=z[:51])
    assert tuple(got.shape) == (51, 31)
    ref = O.greedy_decode_noattack(_cases.params("Transeiver_Star"), O.Spec("Transeiver_Star"), tail.cpu().long(), 0.0, "AWGN",
                                   0.2, z[:51].cpu())
    assert (got.cpu() == ref).all(1).float().mean() >= 49 / 51


def test_bleu_score_string_interface_on_the_whole_test_set(dev):
    """BleuScore.compute_score (utlis/tools.py:37-43, lists of strings in, list of scores out) on all 7,347 reference
    sentences against the string-domain oracle, and the id-domain path against the same."""
    import random
    from deepsc_gan_b200.utlis.tools import BleuScore, SeqtoText
    ids, vocab = _cases.europarl_test()
    rev = {i: t for t, i in vocab.items()}
    st = SeqtoText(vocab, 2)
    rng = random.Random(3)
    hyp = np.asarray([corrupt_row(row, rng, ids) for row in ids], dtype=np.int32)
    real = [st.sequence_to_text(r) for r in ids]
    pred = [st.sequence_to_text(h) for h in hyp]
    for w in ((1, 0, 0, 0), (0.25, 0.25, 0.25, 0.25)):
        bs = BleuScore(*w)
        got = np.asarray(bs.compute_score(real, pred))
        want = np.asarray([B.string_bleu(r, h, rev, w)[0] for r, h in zip(ids, hyp)])
        assert got.shape == (7347,) and np.abs(got - want).max() < 1e-12
        got_ids = np.asarray(bs.score_from_ids(torch.from_numpy(ids).to(dev), torch.from_numpy(hyp).to(dev)))
        assert np.abs(got_ids - want).max() < 1e-12


def corrupt_row(row, rng, ids):
    out = []
    for t in row:
        t = int(t)
        r = rng.random()
        if r < 0.12:
            out.append(int(ids[rng.randrange(len(ids)), 1 + rng.randrange(5)]))
        elif r < 0.18:
            continue
        elif r < 0.24:
            out.extend([t, t])
        else:
            out.append(t)
    out = out[:31]
    return out + [0] * (31 - len(out))
