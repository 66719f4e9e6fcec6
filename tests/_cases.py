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


def params(kind: str, gain: float = 1.0):
    """Keras-default init (seed 2024) with perturbed biases / LN affine so every parameter matters."""
    P = O.init_params(O.Spec(kind), seed=WEIGHT_SEED, randomize_affine=True)
    if gain != 1.0:
        for k in P:
            if "/wq/" in k or "/wk/" in k:
                P[k] = P[k] * gain
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
