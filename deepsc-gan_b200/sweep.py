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
                 seed: int = 0, attack: Optional[str] = None, psr_db: float = -6.0, graph: Optional[bool] = None):
        """``attack="generator"`` (config 4): the perturbation of ``net.generator`` at a fixed perturbation-to-signal
        ratio.  With ||p||_F = 1 the term sqrt(size)*p has unit mean square, so the perturbation power is
        n_std^2 * PNR and PSR_dB = PNR_dB - SNR_dB: at fixed PSR the scale n_std*sqrt(PNR)*sqrt(size)
        = sqrt(10^(PSR/10) * size) is the same at every SNR point."""
        self.net, self.U, self.channel, self.detector = net, units_per_launch, channel, detector
        self.attack, self.psr_db = attack, psr_db
        self.S = units_per_launch * 64
        # graph=True: the 30-step decode loop (static shapes, no random numbers) is replayed as ONE CUDA graph per call;
        # None keeps engine.make_decoder's default (graph for the launch-bound baseline decoder, eager star decoders)
        self.decoder = engine.make_decoder(net, self.S, max_length, graph=graph)
        self.dev = net.semantic_decoder.embedding.embeddings.device
        self.seed = seed
        self.launch = 0

    def run(self, inp: torch.Tensor, n_std: torch.Tensor, h: Optional[torch.Tensor] = None,
            noise: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """inp [S,31] int32 (device), n_std [U] (device) -> (ids [S,31] int32, counts [S,10] int32)."""
        self.launch += 1
        p_scale = None
        if self.attack == "generator":
            elems = 64 * 31 * 16
            p_scale = torch.full((self.U,), math.sqrt(10 ** (self.psr_db / 10) * elems), device=n_std.device,
                                 dtype=torch.float32)
        ids = engine.greedy_units(self.net, inp, self.U, n_std, channel=self.channel, noise=noise, seed=self.seed,
                                  offset=self.launch, h=h, detector=self.detector, decoder=self.decoder,
                                  attack=self.attack, p_scale=p_scale)
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


def evaluate_sweep(runner, units: torch.Tensor, snr_points: Sequence[float] = SNR_POINTS, *, channel: str = "AWGN", K: int = 0,
                   rank: int = 0, world: int = 1, group=None, h_seed: int = 7,
                   weight_sets=((1, 0, 0, 0), (0.25, 0.25, 0.25, 0.25)), noise_for_item=None, h_all=None):
    """The reference's missing SNR-sweep driver: every 64-sentence unit of ``units``
    [n_units*64, 31] at every SNR point -> greedy ids -> BLEU counts -> rows ``[snr_idx, mean score per weight set...]``
    in the layout of the reference's log/eval-D-GAN-STAR/*.pkl.

    Work items (snr, unit) are dealt to ranks in contiguous blocks (``shard_items``); ``runner`` (a ``SweepRunner`` or
    anything with ``U``, ``dev`` and ``run(inp, n_std, h=...) -> (ids, counts)``) evaluates ``runner.U`` items per call,
    the ragged tail being padded with repeats that are dropped again.  The only exchange is the final gather of the
    int32 count table; float BLEU is formed from the gathered table in fp64, so it does not depend on ``world``.
    Injected randomness (parity tests): ``noise_for_item(i)`` -> unit-normal [64,31,16] host tensor for global work item
    ``i`` (default: the runner's on-device Philox stream); ``h_all`` [n_items, 2] fading coefficients (default: drawn
    from ``h_seed``).
    Returns (rows, counts [n_items*64, 10] int32 on the host, snr index per sentence)."""
    import torch.distributed as dist
    n_units = units.shape[0] // 64
    items = work_items(n_units, snr_points)
    lo, hi = shard_items(len(items), rank, world)
    mine = items[lo:hi]
    U, dev = runner.U, runner.dev
    gen = torch.Generator().manual_seed(h_seed)
    if channel == "AWGN":
        h_all = None
    elif h_all is None:
        h_all = fading_coefficients(K, len(items), gen)                              # indexed by global item id
    out = torch.zeros((len(mine) * 64, 10), dtype=torch.int32)
    for b0 in range(0, len(mine), U):
        blk = mine[b0:b0 + U]
        pad = blk + [blk[-1]] * (U - len(blk))
        inp = torch.cat([units[64 * u:64 * u + 64] for _, u in pad], dim=0).to(dev)
        n_std = torch.tensor([snr_to_noise(snr_points[s]) for s, _ in pad], dtype=torch.float32, device=dev)
        h = None
        idx = [lo + b0 + min(j, len(blk) - 1) for j in range(U)]
        if h_all is not None:
            h = h_all[idx].to(dev).contiguous()
        if noise_for_item is not None:
            _, counts = runner.run(inp, n_std, h=h, noise=torch.cat([noise_for_item(i) for i in idx]).to(dev))
        else:
            _, counts = runner.run(inp, n_std, h=h)
        out[b0 * 64:(b0 + len(blk)) * 64] = counts[: len(blk) * 64].to("cpu", torch.int32)
    if world > 1 and dist.is_available() and dist.is_initialized():
        # ragged shards: pad to the largest block, gather, trim
        most = max(shard_items(len(items), r, world)[1] - shard_items(len(items), r, world)[0] for r in range(world))
        backend_dev = dev if dist.get_backend(group) == "nccl" else torch.device("cpu")
        buf = torch.zeros((most * 64, 10), dtype=torch.int32, device=backend_dev)
        buf[: out.shape[0]] = out.to(backend_dev)
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
        sizes = [shard_items(len(items), r, world) for r in range(world)]
        out = torch.cat([parts[r][: (b - a) * 64].cpu() for r, (a, b) in enumerate(sizes)], dim=0)
        mine_all = items
    else:
        mine_all = mine
    snr_index = np.repeat(np.array([s for s, _ in mine_all], dtype=np.int64), 64)
    rows = bleu_table(out.numpy(), snr_index, len(snr_points), weight_sets)
    return rows, out, snr_index


def bleu_table(counts: np.ndarray, snr_index: np.ndarray, n_snr: int,
               weight_sets=((1, 0, 0, 0), (0.25, 0.25, 0.25, 0.25))) -> List[List[float]]:
    """Rows [snr_idx, mean score per weight set...], the layout of the reference's result pickles."""
    rows = []
    for s in range(n_snr):
        sel = counts[snr_index == s]
        rows.append([s] + [float(score_from_counts(sel, w).mean()) if len(sel) else float("nan") for w in weight_sets])
    return rows
