"""Optimizer and gradient plumbing for the training steps (utlis/trainer.py, utlis/gan_train.py).

The reference builds its optimizer in the (missing) driver; upstream DeepSC uses
``tf.keras.optimizers.Adam(lr, beta_1=0.9, beta_2=0.98, epsilon=1e-8)`` and ``utlis/parameters.py:22`` fixes
``lr = 5e-4``; that is the default here (stated as an assumption).

All parameters of a model live in ONE flat fp32 buffer (``FlatParams``): every ``nn.Parameter`` is re-pointed to a
view of it, and gradients are accumulated by autograd into views of flat gradient buffers.  One Adam launch
(dsc_adam_step) then updates any contiguous range of parameters, and data-parallel training all-reduces one flat
bucket over NCCL (one collective per step) instead of one per variable.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from . import autograd as AG


class FlatParams:
    """Flat storage for the parameters of ``net`` in ``named_parameters()`` order."""

    def __init__(self, net: torch.nn.Module, n_grad_buffers: int = 1):
        named = list(net.named_parameters())
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        dev = self.params[0].device
        sizes = [p.numel() for p in self.params]
        self.offsets, off = [], 0
        for n in sizes:
            self.offsets.append(off)
            off += (n + 3) // 4 * 4                     # keep every view 16-byte aligned (float4 kernels)
        self.numel = off
        self.flat = torch.zeros((off,), device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                self.flat[o:o + p.numel()].view_as(p).copy_(p)
                p.data = self.flat[o:o + p.numel()].view_as(p)
        # gradient buffers live side by side so that ONE all-reduce covers all of them
        self.grad_bucket = torch.zeros((n_grad_buffers, off), device=dev, dtype=torch.float32)
        self.index: Dict[str, int] = {n: i for i, n in enumerate(self.names)}

    def grad_views(self, k: int) -> List[torch.Tensor]:
        g = self.grad_bucket[k]
        return [g[o:o + p.numel()].view_as(p) for p, o in zip(self.params, self.offsets)]

    def point_grads(self, k: int) -> None:
        """Make autograd accumulate into gradient buffer ``k``."""
        for p, v in zip(self.params, self.grad_views(k)):
            p.grad = v

    def ranges(self, select) -> List[Tuple[int, int]]:
        """Merged [begin, end) element ranges of the parameters whose name satisfies ``select``."""
        out: List[Tuple[int, int]] = []
        for i, n in enumerate(self.names):
            if not select(n):
                continue
            b = self.offsets[i]
            e = self.offsets[i + 1] if i + 1 < len(self.offsets) else self.numel
            if out and out[-1][1] == b:
                out[-1] = (out[-1][0], e)
            else:
                out.append((b, e))
        return out

    def select(self, select) -> List[torch.nn.Parameter]:
        return [p for n, p in zip(self.names, self.params) if select(n)]


class Adam:
    """tf.keras.optimizers.Adam on a ``FlatParams``: one ``iterations`` counter for the whole optimizer (incremented
    by every ``apply``, as ``apply_gradients`` does), per-element moments."""

    def __init__(self, flat: FlatParams, learning_rate: float = 5e-4, beta_1: float = 0.9, beta_2: float = 0.98,
                 epsilon: float = 1e-8):
        self.fp, self.lr, self.b1, self.b2, self.eps = flat, learning_rate, beta_1, beta_2, epsilon
        self.m = torch.zeros_like(flat.flat)
        self.v = torch.zeros_like(flat.flat)
        self.iterations = 0

    def apply(self, ranges: Sequence[Tuple[int, int]], grad: torch.Tensor, grad_scale: float = 1.0,
              grad2: Optional[torch.Tensor] = None, grad2_scale: float = 0.0, lr: Optional[float] = None) -> None:
        """One ``apply_gradients`` call over the given element ranges with g = grad*grad_scale + grad2*grad2_scale."""
        self.iterations += 1
        for b, e in ranges:
            AG.adam_step(self.fp.flat[b:e], grad[b:e], self.m[b:e], self.v[b:e], self.lr if lr is None else lr,
                         self.iterations, self.b1, self.b2, self.eps, grad_scale,
                         None if grad2 is None else grad2[b:e], grad2_scale)


def mean_scale(group=None) -> float:
    """1/world inside a process group, else 1: the factor Adam folds into its gradient read after a SUM all-reduce."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    return 1.0 / dist.get_world_size(group)


def all_reduce_mean_scale(bucket: torch.Tensor, group=None) -> float:
    """Sum the flat gradient bucket over the data-parallel ranks (NCCL) and return the 1/world factor that the Adam
    kernel folds into its gradient read.  No-op (factor 1) outside a process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    if world == 1:
        return 1.0
    dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world
