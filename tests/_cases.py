"""Seeded parity cases shared by the golden generator and the tests."""
from __future__ import annotations

import os

import numpy as np
import torch

import deepsc_gan_b200  # noqa: F401  (registers the package)
from deepsc_gan_b200.dataset.synthetic import synthetic_unit
from oracle import bleu_oracle, deepsc_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KINDS = ("Transeiver", "Transeiver_star", "Transeiver_Star", "Transeiver_GAN")
WEIGHT_SEED = 2024
NOISE_SEED = 7
SNR_DB = 6.0


def params(kind: str, gain: float = 1.0, emb_gain: float = 1.0):
    """Keras-default init (seed 2024) with perturbed biases / LN affine so every parameter matters.  ``gain`` scales the
    query / key projections (sharper attention), ``emb_gain`` the embedding tables (token identity dominates the
    positional code): with both > 1 the decoded ids depend on the input sentence the way a trained model's do."""
    P = O.init_params(O.Spec(kind), seed=WEIGHT_SEED, randomize_affine=True)
    if gain != 1.0 or emb_gain != 1.0:
        for k in P:
            if "/wq/" in k or "/wk/" in k:
                P[k] = P[k] * gain
            elif k.endswith("/embedding/embeddings"):
                P[k] = P[k] * emb_gain
    return P


def draws(seed: int = NOISE_SEED, n: int = 64):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(n, 31, 16, generator=g)
    z_r = torch.randn(n, 31, 16, generator=g)
    p = torch.randn(n, 31, 16, generator=g)
    p = p / torch.linalg.vector_norm(p)
    h = torch.randn(4, generator=g).tolist()
    return z, z_r, p, (h[0], h[1]), (h[2], h[3])


def oracle_case(kind: str, channel: str = "AWGN", unit_index: int = 0, greedy: bool = True):
    """Teacher-forced forward + greedy decode of one 64-sentence unit through the oracle."""
    spec = O.Spec(kind)
    P = params(kind)
    inp = synthetic_unit(unit_index).long()
    z, z_r, p, h_z, h_z_r = draws()
    n_std = O.snr_to_noise(SNR_DB)
    tar_inp = inp[:, :-1]
    masks = O.create_masks(inp, tar_inp)
    out = {"inp": inp.numpy().astype(np.int32)}
    with torch.no_grad():
        fw = O.transceiver_forward(P, spec, inp, tar_inp, p, 3.0, channel, n_std, *masks, z=z, h_z=h_z, z_r=z_r,
                                   h_z_r=h_z_r, traingan=(kind == "Transeiver_GAN"))
        if kind == "Transeiver_GAN":
            pred_p, pred_r, x, y_r = fw
            out["pred_r_lse"] = torch.logsumexp(pred_r, -1).numpy()
            out["pred_r_argmax"] = pred_r.argmax(-1).numpy().astype(np.int32)
            pred, y = pred_p, y_r
        else:
            pred, x, y, _ = fw
        tar_real = inp if spec.is_star else inp[:, 1:]
        out.update(symbols=x.numpy(), received=y.numpy(), lse=torch.logsumexp(pred, -1).numpy(),
                   tf_argmax=pred.argmax(-1).numpy().astype(np.int32),
                   logits_slice=pred[:4, :, :64].numpy().copy(), loss=np.float32(O.loss_function(tar_real, pred)))
        if greedy:
            ids = O.greedy_decode_noattack(P, spec, inp, 0.0, channel, n_std, z, h_z)
            out["greedy_ids"] = ids.numpy().astype(np.int32)
            out["bleu_counts"] = bleu_oracle.bleu_counts(out["inp"], out["greedy_ids"])
    return out


def golden_path(kind: str, channel: str) -> str:
    return os.path.join(GOLDEN_DIR, f"golden_{kind}_{channel}.npz")


def max_rel(a, b) -> float:
    """Element-wise relative error with the tensor's RMS as the absolute floor: max |a - b| / (|b| + rms(b)).  (A purely
    element-wise |a - b| / |b| is meaningless for logits and symbols that cross zero.)"""
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    rms = float(b.pow(2).mean().sqrt()) + 1e-30
    return float(((a - b).abs() / (b.abs() + rms)).max())


# --------------------------------------------------------------------------- margin-enforced greedy cases
# name -> (system, channel, first synthetic unit, SNR dB, PSR dB of the generator attack or None, weight variant).
# Weight variants: "keras" = the seeded Keras initialisation; "lively" = the same with the embeddings x4 and the
# query / key projections x3, which makes the decoded ids depend on the input (about 170 distinct ids per unit for the
# baseline decoder instead of about 10), so that an id comparison exercises more than one constant token.
MARGIN_CASES = {
    "Transeiver_AWGN": ("Transeiver", "AWGN", 40, 6.0, None, "keras"),                    # BASELINE.json configs[0]
    "Transeiver_AWGN_lively": ("Transeiver", "AWGN", 50, 6.0, None, "lively"),
    "Transeiver_Star_AWGN": ("Transeiver_Star", "AWGN", 41, 6.0, None, "keras"),          # configs[1]
    "Transeiver_Star_AWGN_0dB": ("Transeiver_Star", "AWGN", 45, 0.0, None, "keras"),      # configs[1], the noisiest sweep point
    "Transeiver_Star_AWGN_12dB": ("Transeiver_Star", "AWGN", 46, 12.0, None, "keras"),
    "Transeiver_Star_AWGN_18dB": ("Transeiver_Star", "AWGN", 47, 18.0, None, "keras"),    # configs[1], the cleanest sweep point
    "Transeiver_Star_AWGN_lively": ("Transeiver_Star", "AWGN", 51, 3.0, None, "lively"),
    "Transeiver_Star_Rayleigh": ("Transeiver_Star", "Rayleigh", 42, 12.0, None, "keras"), # configs[2]
    "Transeiver_Star_Rayleigh_6dB": ("Transeiver_Star", "Rayleigh", 48, 6.0, None, "keras"),
    "Transeiver_star_AWGN": ("Transeiver_star", "AWGN", 43, 6.0, None, "keras"),          # the 4-layer star codec
    "Transeiver_GAN_generator": ("Transeiver_GAN", "AWGN", 44, 9.0, -6.0, "keras"),       # configs[3]: generator attack, PSR -6 dB
    "Transeiver_GAN_generator_lively": ("Transeiver_GAN", "AWGN", 52, 9.0, -6.0, "lively"),
}


def margin_params(name: str):
    kind, variant = MARGIN_CASES[name][0], MARGIN_CASES[name][5]
    return params(kind, gain=3.0, emb_gain=4.0) if variant == "lively" else params(kind)


def margin_path(name: str) -> str:
    return os.path.join(GOLDEN_DIR, f"margin_{name}.npz")


def margin_noise(seeds) -> torch.Tensor:
    """Unit-normal channel draw [64,31,16] of a margin case: sentence slot b draws from its own generator seeds[b]."""
    return torch.stack([torch.randn(31, 16, generator=torch.Generator().manual_seed(int(sd))) for sd in seeds])


def greedy_with_margin(kind: str, P, inp: torch.Tensor, channel: str, snr_db: float, psr_db, seeds, dtype=torch.float32):
    """Oracle greedy decode -> (ids [n,31] int32, per-sentence minimum over the 30 steps of (top1 - top2) / max|logit|)."""
    spec = O.Spec(kind)
    P = O.to_dtype(P, dtype)
    h_z = draws()[3]
    z = margin_noise(seeds).to(dtype)
    with torch.no_grad():
        if psr_db is not None:
            ids, lg = O.greedy_generator_attack(P, spec, inp.long(), snr_db, psr_db, z, return_logits=True)
        else:
            ids, lg = O.greedy_decode_noattack(P, spec, inp.long(), 0.0, channel, O.snr_to_noise(snr_db), z, h_z,
                                               return_logits=True)
    top2 = lg.topk(2, -1).values
    rel = (top2[..., 0] - top2[..., 1]) / lg.abs().amax(-1)
    return ids, rel.min(1).values.double()


def margin_oracle(name: str, inp: torch.Tensor, seeds, dtype=torch.float32):
    """Oracle greedy decode of a margin case -> (ids, per-sentence minimum relative top-2 margin)."""
    kind, channel, _, snr_db, psr_db, _ = MARGIN_CASES[name]
    return greedy_with_margin(kind, margin_params(name), inp, channel, snr_db, psr_db, seeds, dtype)


# The margin-enforced SNR-sweep case (tests/golden/margin_sweep_<channel>.npz): SWEEP_UNITS 64-sentence units at every
# point of SWEEP_SNRS = SWEEP_UNITS * len(SWEEP_SNRS) work items in the order of sweep.work_items (SNR-major); item i has
# its own noise seeds [64]; the fading coefficient of every item is draws()[3].
SWEEP_SNRS = (0.0, 9.0, 18.0)
SWEEP_UNITS = 2
SWEEP_FIRST_UNIT = 60


def sweep_margin_path(channel: str) -> str:
    return os.path.join(GOLDEN_DIR, f"margin_sweep_{channel}.npz")


def max_rel(a, b) -> float:
    """Element-wise relative error with the tensor's RMS as the absolute floor: max |a - b| / (|b| + rms(b)).  (A purely
    element-wise |a - b| / |b| is meaningless for logits and symbols that cross zero.)"""
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    rms = float(b.pow(2).mean().sqrt()) + 1e-30
    return float(((a - b).abs() / (b.abs() + rms)).max())


# --------------------------------------------------------------------------- margin-enforced greedy cases
# name -> (system, channel, first synthetic unit, SNR dB, PSR dB of the generator attack or None, weight variant).
# Weight variants: "keras" = the seeded Keras initialisation; "lively" = the same with the embeddings x4 and the
# query / key projections x3, which makes the decoded ids depend on the input (about 170 distinct ids per unit for the
# baseline decoder instead of about 10), so that an id comparison exercises more than one constant token.
MARGIN_CASES = {
    "Transeiver_AWGN": ("Transeiver", "AWGN", 40, 6.0, None, "keras"),                    # BASELINE.json configs[0]
    "Transeiver_AWGN_lively": ("Transeiver", "AWGN", 50, 6.0, None, "lively"),
    "Transeiver_Star_AWGN": ("Transeiver_Star", "AWGN", 41, 6.0, None, "keras"),          # configs[1]
    "Transeiver_Star_AWGN_0dB": ("Transeiver_Star", "AWGN", 45, 0.0, None, "keras"),      # configs[1], the noisiest sweep point
    "Transeiver_Star_AWGN_12dB": ("Transeiver_Star", "AWGN", 46, 12.0, None, "keras"),
    "Transeiver_Star_AWGN_18dB": ("Transeiver_Star", "AWGN", 47, 18.0, None, "keras"),    # configs[1], the cleanest sweep point
    "Transeiver_Star_AWGN_lively": ("Transeiver_Star", "AWGN", 51, 3.0, None, "lively"),
    "Transeiver_Star_Rayleigh": ("Transeiver_Star", "Rayleigh", 42, 12.0, None, "keras"), # configs[2]
    "Transeiver_Star_Rayleigh_6dB": ("Transeiver_Star", "Rayleigh", 48, 6.0, None, "keras"),
    "Transeiver_star_AWGN": ("Transeiver_star", "AWGN", 43, 6.0, None, "keras"),          # the 4-layer star codec
    "Transeiver_GAN_generator": ("Transeiver_GAN", "AWGN", 44, 9.0, -6.0, "keras"),       # configs[3]: generator attack, PSR -6 dB
    "Transeiver_GAN_generator_lively": ("Transeiver_GAN", "AWGN", 52, 9.0, -6.0, "lively"),
}


def margin_params(name: str):
    kind, variant = MARGIN_CASES[name][0], MARGIN_CASES[name][5]
    return params(kind, gain=3.0, emb_gain=4.0) if variant == "lively" else params(kind)


def margin_path(name: str) -> str:
    return os.path.join(GOLDEN_DIR, f"margin_{name}.npz")


def margin_noise(seeds) -> torch.Tensor:
    """Unit-normal channel draw [64,31,16] of a margin case: sentence slot b draws from its own generator seeds[b]."""
    return torch.stack([torch.randn(31, 16, generator=torch.Generator().manual_seed(int(sd))) for sd in seeds])


def margin_oracle(name: str, inp: torch.Tensor, seeds, dtype=torch.float32):
    """Oracle greedy decode of a margin case -> (ids [64,31] int32, per-sentence minimum over the 30 steps of
    (top1 - top2) / max|logit|)."""
    kind, channel, _, snr_db, psr_db, _ = MARGIN_CASES[name]
    spec = O.Spec(kind)
    P = O.to_dtype(margin_params(name), dtype)
    h_z = draws()[3]
    z = margin_noise(seeds).to(dtype)
    with torch.no_grad():
        if psr_db is not None:
            ids, lg = O.greedy_generator_attack(P, spec, inp.long(), snr_db, psr_db, z, return_logits=True)
        else:
            ids, lg = O.greedy_decode_noattack(P, spec, inp.long(), 0.0, channel, O.snr_to_noise(snr_db), z, h_z,
                                               return_logits=True)
    top2 = lg.topk(2, -1).values
    rel = (top2[..., 0] - top2[..., 1]) / lg.abs().amax(-1)
    return ids, rel.min(1).values.double()


def europarl_test():
    """The reference's whole test set (tests/golden/make_europarl_fixture.py): (ids [7347,31] int32 padded post,
    token_to_idx dict of all 22,234 tokens)."""
    fx = np.load(os.path.join(GOLDEN_DIR, "europarl_test.npz"))
    tokens = bytes(fx["tokens"]).decode("utf-8").split("\n")
    return fx["ids"].astype(np.int32), {tok: i for i, tok in enumerate(tokens)}
