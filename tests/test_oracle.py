"""CPU checks of the oracle itself: structure against the checkpoint inventory, internal consistency,
and drift against the frozen goldens.  (The reference ships no tests or vectors: parity unpinned.)"""
import json
import os

import numpy as np
import pytest
import torch

import _cases
from oracle import deepsc_oracle as O

INV = json.load(open(os.path.join(_cases.GOLDEN_DIR, "ckpt_inventory.json")))


@pytest.mark.parametrize("kind,root", [("Transeiver_Star", "Transceiver_Star"), ("Transeiver_star", "Transceiver_star")])
def test_parameter_inventory_matches_reference_checkpoint(kind, root):
    """Names and shapes equal the variables of the reference's own checkpoints (App. C): this pins the
    identity FFN (no sl2 variables), the single STE/STD, and which LayerNorms exist."""
    P = O.init_params(O.Spec(kind))
    ref = {k[len(root) + 1:]: tuple(v) for k, v in INV[kind]["variables"].items()}
    mine = {k: tuple(v.shape) for k, v in P.items()}
    assert mine == ref
    assert O.param_count(P) == INV[kind]["total_parameters"]


def test_derived_parameter_counts():
    assert O.param_count(O.init_params(O.Spec("Transeiver"))) == 9524458
    P = O.init_params(O.Spec("Transeiver_GAN"))
    assert O.param_count(P) == 9532922 and len(P) == 120
    # utlis/gan_train.py:36 relies on G sitting at trainable_variables[104:108]
    names = [k for k in P if not k.startswith("generator/")]
    assert len(names) == 116


def test_positional_table_is_the_reference_formula():
    pe = O.positional_table(512, 128).numpy()
    pos, i = 7, 5
    ang = pos / np.power(10000, (2 * i) / np.float32(128))      # 2*i, not 2*(i//2)
    assert abs(pe[pos, i] - np.cos(ang)) < 1e-6
    assert abs(pe[pos, 4] - np.sin(pos / np.power(10000, 8 / np.float32(128)))) < 1e-6


def _dedup_cycles(P, pre, e, h2, cycle_num, relay):
    """Deduplicated star formulation (each node projected once, neighbours by index)."""
    b, l, d = e.shape
    H, dh = 8, 16
    Wq, Wk, Wv = (P[f"{pre}/multi_att_satellite/{w}/kernel"] for w in ("wq", "wk", "wv"))
    Wo, bo = P[f"{pre}/multi_att_satellite/dense/kernel"], P[f"{pre}/multi_att_satellite/dense/bias"]
    ke, ve = e @ Wk, e @ Wv
    h, s = e, e.mean(1)
    for _ in range(cycle_num):
        q, k, v = h @ Wq, h @ Wk, h @ Wv
        ks, vs = (s @ Wk)[:, None].expand(b, l, d), (s @ Wv)[:, None].expand(b, l, d)
        keys = torch.stack([k.roll(-1, 1), k, k.roll(1, 1), ke, ks], 2).reshape(b, l, 5, H, dh)
        vals = torch.stack([v.roll(-1, 1), v, v.roll(1, 1), ve, vs], 2).reshape(b, l, 5, H, dh)
        w = torch.softmax(torch.einsum("blhd,bljhd->blhj", q.reshape(b, l, H, dh), keys) / 4.0, -1)
        h = torch.relu(torch.einsum("blhj,bljhd->blhd", w, vals).reshape(b, l, d) @ Wo + bo)
        parts = [s[:, None], h] + ([h2] if h2 is not None else [])
        m = torch.cat(parts, 1)
        s = torch.relu(O.mha(P, f"{pre}/{relay}", s[:, None], m, m, None))[:, 0]
    return h, s


def test_literal_star_equals_deduplicated_form():
    P = O.to_dtype(_cases.params("Transeiver_Star"), torch.float64)
    g = torch.Generator().manual_seed(3)
    e = torch.randn(3, 31, 128, generator=g, dtype=torch.float64)
    h2 = torch.randn(3, 30, 128, generator=g, dtype=torch.float64)
    pre = "semantic_decoder/dec_layers"
    h_a, s_a = O._star_cycles(P, pre, e, h2, 8, "multi_att_relay")
    h_b, s_b = _dedup_cycles(P, pre, e, h2, 8, "multi_att_relay")
    assert torch.allclose(h_a, h_b, atol=1e-12) and torch.allclose(s_a, s_b, atol=1e-12)


def test_masks_and_loss():
    inp = torch.tensor([[1, 7, 9, 2, 0, 0]])
    enc, comb, dec = O.create_masks(inp, inp[:, :-1])
    assert enc.shape == (1, 1, 1, 6) and comb.shape == (1, 1, 5, 5)
    assert comb[0, 0, 0].tolist() == [0, 1, 1, 1, 1] and comb[0, 0, 4].tolist() == [0, 0, 0, 0, 1]
    logits = torch.zeros(1, 5, 11)
    # mean over ALL positions, PAD masked: targets [7,9,2,0,0] -> 3 real of 5 -> 3*log(11)/5
    loss = O.loss_function(inp[:, 1:], logits)
    assert abs(float(loss) - 3 * np.log(11) / 5) < 1e-6


def test_snr_and_fading_coefficient():
    assert abs(O.snr_to_noise(6) - 0.5011872336) < 1e-9
    h = O.fading_coeff(1, 0.0, 0.0)
    assert abs(h.real - 0.5) < 1e-12 and abs(h.imag - 0.5) < 1e-12
    assert O.fading_coeff(0, 1.0, -1.0) == complex(np.sqrt(0.5), -np.sqrt(0.5))


def test_fading_returns_unequalised_y_by_default():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 31, 16, generator=g)
    z = torch.zeros(2, 31, 16)
    y = O.fading(x, 0, 0.1, (1.0, 0.5), z)
    est = O.fading(x, 0, 0.1, (1.0, 0.5), z, detector="LS", apply_detector=True)
    assert not torch.allclose(y, x, atol=1e-3)      # reference returns y (transceiver.py:74-75)
    assert torch.allclose(est, x, atol=1e-5)        # the discarded LS estimate inverts the channel
    with pytest.raises(ValueError, match="detector must in LS and MMSE"):
        O.fading(x, 0, 0.1, (1.0, 0.5), z, detector="ZF")


def test_greedy_cached_last_position_equals_full_logits():
    P = _cases.params("Transeiver_Star")
    spec = O.Spec("Transeiver_Star")
    inp = _cases.synthetic_unit(1)[:4].long()
    z = _cases.draws()[0][:4]
    a = O.greedy_decode_noattack(P, spec, inp, 0.0, "AWGN", 0.5, z, max_length=4, last_only=True)
    b = O.greedy_decode_noattack(P, spec, inp, 0.0, "AWGN", 0.5, z, max_length=4, last_only=False)
    assert torch.equal(a, b)


@pytest.mark.parametrize("kind,channel", [("Transeiver_Star", "AWGN"), ("Transeiver", "AWGN")])
def test_oracle_reproduces_frozen_goldens(kind, channel):
    gold = np.load(_cases.golden_path(kind, channel))
    c = _cases.oracle_case(kind, channel, greedy=True)
    assert np.array_equal(c["inp"], gold["inp"])
    assert np.allclose(c["symbols"][:8], gold["symbols"], rtol=1e-4, atol=1e-5)
    assert np.allclose(c["lse"], gold["lse"], rtol=1e-4, atol=1e-4)
    assert (c["greedy_ids"] == gold["greedy_ids"]).mean() > 0.99
    assert np.array_equal(c["bleu_counts"][(c["greedy_ids"] == gold["greedy_ids"]).all(1)],
                          gold["bleu_counts"][(c["greedy_ids"] == gold["greedy_ids"]).all(1)])
