"""SNR-sweep evaluation: (SNR point, 64-sentence unit) work items -> greedy ids -> BLEU counts.

The reference's driver for this loop is missing; the loop body is
``greedy_decode_noattack`` + ``SeqtoText``/``BleuScore`` per batch per SNR (utlis/eval.py:78-117,
utlis/tools.py:10-43).  Work items are independent, so they are dealt to ranks in contiguous blocks and
never split below a unit (per-unit power norm / fading coefficient).  The only exchange is the final
gather of the int32 count table; float BLEU is formed on rank 0 in fp64 so the result does not depend
on the GPU count.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, engine
from .utlis.tools import score_from_counts

SNR_POINTS = tuple(range(0, 19))     # 0..18 dB, the 19 rows of the reference's log/eval-D-GAN-STAR/*.pkl


def snr_to_noise(snr_db: float) -> float:
    return float(1.0 / np.sqrt(10 ** (snr_db / 10)))


def shard_items(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of work items for ``rank``; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def work_items(n_units: int, snr_points: Sequence[float] = SNR_POINTS) -> List[Tuple[int, int]]:
    """(snr_index, unit_index) pairs, SNR-major like the reference's per-SNR evaluation."""
    return [(s, u) for s in range(len(snr_points)) for u in range(n_units)]


class SweepRunner:
    """Runs blocks of work items on one GPU.  ``units_per_launch`` units are stacked along the batch axis."""

    def __init__(self, net, units_per_launch: int, channel: str = "AWGN", detector: int = 0, max_length: int = 30,
                 seed: int = 0):
        self.net, self.U, self.channel, self.detector = net, units_per_launch, channel, detector
        self.S = units_per_launch * 64
        self.decoder = engine.make_decoder(net, self.S, max_length)
        self.dev = net.semantic_decoder.embedding.embeddings.device
        self.seed = seed
        self.launch = 0

    def run(self, inp: torch.Tensor, n_std: torch.Tensor, h: Optional[torch.Tensor] = None,
            noise: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """inp [S,31] int32 (device), n_std [U] (device) -> (ids [S,31] int32, counts [S,10] int32)."""
        self.launch += 1
        ids = engine.greedy_units(self.net, inp, self.U, n_std, channel=self.channel, noise=noise, seed=self.seed,
                                  offset=self.launch, h=h, detector=self.detector, decoder=self.decoder)
        return ids, _lib.bleu_counts(inp, ids)


def fading_coefficients(K: int, n_units: int, generator: torch.Generator) -> torch.Tensor:
    """h = N(mean,std) + j N(mean,std) per unit (models/transceiver.py:39-50) -> [n_units, 2] float32 (CPU)."""
    mean, std = math.sqrt(K / (2 * (K + 1))), math.sqrt(1 / (2 * (K + 1)))
    return mean + std * torch.randn(n_units, 2, generator=generator)


def gather_counts(counts: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather the per-rank count tables (equal shapes) -> [world*n, 10]."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return counts
    parts = [torch.empty_like(counts) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, counts.contiguous(), group=group)
    return torch.cat(parts, dim=0)


def bleu_table(counts: np.ndarray, snr_index: np.ndarray, n_snr: int,
               weight_sets=((1, 0, 0, 0), (0.25, 0.25, 0.25, 0.25))) -> List[List[float]]:
    """Rows [snr_idx, mean score per weight set...], the layout of the reference's result pickles."""
    rows = []
    for s in range(n_snr):
        sel = counts[snr_index == s]
        rows.append([s] + [float(score_from_counts(sel, w).mean()) if len(sel) else float("nan") for w in weight_sets])
    return rows
