"""Synthetic Europarl-shape inputs: the shapes the reference's loader produces
(dataset/dataloader.py:11 pads post to 31, batch 64; dataset/preprocess_text.py:41-48 keeps sentences
of 5..29 words; real sentences are <START> words... '' <END>, i.e. ids 1, content, 4, 2)."""
from __future__ import annotations

import torch

UNIT = 64
SEQ = 31


def synthetic_unit(unit_index: int, vocab_size: int = 22234, unit: int = UNIT) -> torch.Tensor:
    """One 64-sentence unit, int32 [64, 31] on the CPU; seed 1234 + unit_index."""
    g = torch.Generator().manual_seed(1234 + unit_index)
    n_words = torch.randint(4, 29, (unit,), generator=g)            # content length 4..28
    content = torch.randint(5, vocab_size, (unit, SEQ), generator=g)
    out = torch.zeros((unit, SEQ), dtype=torch.int32)
    for b in range(unit):
        n = int(n_words[b])
        out[b, 0] = 1
        out[b, 1:1 + n] = content[b, :n].to(torch.int32)
        out[b, 1 + n] = 4
        out[b, 2 + n] = 2
    return out


def synthetic_units(first: int, count: int, vocab_size: int = 22234) -> torch.Tensor:
    return torch.cat([synthetic_unit(first + i, vocab_size) for i in range(count)], dim=0)
