"""Input pipeline: mirror of DeepSC-GAN/dataset/dataloader.py (return_dataset :4-17, return_loader :19-24).

The reference builds a ``tf.data`` pipeline: unpickle a list of id lists, ``pad_sequences(maxlen=31, padding='post')``
(which also TRUNCATES from the front when a sentence is longer than 31 ids, the Keras default ``truncating='pre'``),
``(inp, tar) = (x, x)``, ``shuffle(args.shuffle_size)``, ``batch(args.bs)`` without drop-remainder, prefetch.  Here the
padded matrix is one pinned int32 host tensor and a batch is a slice of a permutation, copied to the device
asynchronously; iteration yields ``(inp, tar)`` pairs of shape ``[<=bs, 31]`` like the reference's dataset.

``as_units()`` is the layout the SNR sweep consumes: whole 64-sentence units (the ragged tail is dropped there and
only there, SURVEY.md App. B Q16: power norm / fading / FGM are per-unit scalars).
"""
from __future__ import annotations

import pickle
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

MAXLEN = 31


def pad_sequences(sequences: Sequence[Sequence[int]], maxlen: int = MAXLEN, padding: str = "post",
                  truncating: str = "pre", value: int = 0) -> np.ndarray:
    """tf.keras.preprocessing.sequence.pad_sequences for int32 id lists (defaults as used at dataloader.py:11)."""
    out = np.full((len(sequences), maxlen), value, dtype=np.int32)
    for i, s in enumerate(sequences):
        s = list(s)
        if len(s) > maxlen:
            s = s[-maxlen:] if truncating == "pre" else s[:maxlen]
        if padding == "post":
            out[i, : len(s)] = s
        else:
            out[i, maxlen - len(s):] = s
    return out


class Dataset:
    """Iterable of ``(inp, tar)`` batches; ``len()`` = number of batches (the last one may be short)."""

    def __init__(self, data: np.ndarray, batch_size: int = 64, shuffle: bool = True, device=None, seed: int = 0):
        self.data = torch.from_numpy(np.ascontiguousarray(data, dtype=np.int32))
        if torch.cuda.is_available():
            self.data = self.data.pin_memory()
        self.bs, self.shuffle, self.device = int(batch_size), shuffle, device
        self.gen = torch.Generator().manual_seed(seed)

    def __len__(self) -> int:
        return (self.data.shape[0] + self.bs - 1) // self.bs

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        n = self.data.shape[0]
        # shuffle(buffer >= dataset size) is a uniform permutation, re-drawn every epoch (reshuffle_each_iteration)
        order = torch.randperm(n, generator=self.gen) if self.shuffle else torch.arange(n)
        for b0 in range(0, n, self.bs):
            batch = self.data[order[b0:b0 + self.bs]]
            if self.device is not None:
                batch = batch.pin_memory().to(self.device, non_blocking=True) if torch.cuda.is_available() else batch.to(self.device)
            yield batch, batch

    def as_units(self, unit: int = 64) -> torch.Tensor:
        """[n_units * 64, 31] int32 (host): the full units in file order, ragged tail dropped."""
        n = self.data.shape[0] // unit * unit
        return self.data[:n]


def return_dataset(args, path, length: int = -1, device=None, shuffle: bool = True, seed: int = 0) -> Dataset:
    """dataloader.py:4-17.  ``length=-1`` keeps the reference's ``raw_data[:-1]`` slice (its last sentence is dropped)."""
    with open(path, "rb") as f:
        raw_data = pickle.load(f)
    return Dataset(pad_sequences(raw_data[:length], maxlen=MAXLEN, padding="post"), getattr(args, "bs", 64), shuffle,
                   device, seed)


def return_loader(args, device=None):
    """dataloader.py:19-24."""
    return (return_dataset(args, args.train_save_path, -1, device), return_dataset(args, args.test_save_path, -1, device))
