"""Metrics helpers: mirror of DeepSC-GAN/utlis/tools.py (SeqtoText, BleuScore, SNR_to_noise).

``BleuScore`` keeps the reference's string interface (lists of sentences in, list of scores out) and
adds the id-domain path the eval loop uses on the GPU: ``counts_from_ids`` runs the integer n-gram
kernel (dsc_bleu_counts) and ``score_from_counts`` forms the float score on the host in fp64 with the
formula of nltk's ``sentence_bleu`` (method0 smoothing), see SURVEY.md App. D.  ``Similarity`` (BERT
cosine, tools.py:53-103) keeps the reference's scoring arithmetic (``similarity_from_features``); the BERT encoder it
scores (bert4keras + a checkpoint the reference does not ship) is injected by the caller.
"""
from __future__ import annotations

import math
import re
import sys
from typing import Dict, List, Sequence

import numpy as np
import torch

from .. import _lib

_RE_TAGS = re.compile(r"<[a-zA-Z\/!].*?>", re.DOTALL | re.IGNORECASE)


class SeqtoText:
    """utlis/tools.py:10-27."""

    def __init__(self, vocb_dictionary: Dict[str, int], end_idx: int):
        self.reverse_word_map = dict(zip(vocb_dictionary.values(), vocb_dictionary.keys()))
        self.vocb_dictionary = vocb_dictionary
        self.end_idx = end_idx

    def sequence_to_text(self, list_of_indices) -> str:
        words = []
        for idx in list_of_indices:
            idx = int(idx)
            if idx == self.end_idx:
                break
            words.append(self.reverse_word_map.get(idx))
        return ' '.join(words)

    def text_to_sequence(self, text: str) -> List[int]:
        """Inverse map used by ``BleuScore.compute_score`` to reach the id-domain kernel."""
        unk = self.vocb_dictionary.get('<UNK>', 3)
        return [self.vocb_dictionary.get(w, unk) for w in text.split()]

    def show(self):
        print(self.reverse_word_map)


def score_from_counts(counts, weights=(0.25, 0.25, 0.25, 0.25)) -> np.ndarray:
    """counts [n,10] int (match_1..4, total_1..4, hyp_len, ref_len) -> fp64 sentence BLEU, following nltk
    corpus_bleu for a single pair: 0 when there is no unigram match; brevity penalty; precisions with a
    zero numerator replaced by sys.float_info.min (SmoothingFunction.method0)."""
    c = np.asarray(counts.cpu() if isinstance(counts, torch.Tensor) else counts, dtype=np.int64).reshape(-1, 10)
    out = np.zeros((c.shape[0],), dtype=np.float64)
    for r, row in enumerate(c):
        match, total, hyp_len, ref_len = row[0:4], row[4:8], int(row[8]), int(row[9])
        if match[0] == 0:
            continue
        bp = 1.0 if hyp_len > ref_len else (0.0 if hyp_len == 0 else math.exp(1 - ref_len / hyp_len))
        terms = []
        for w, m, t in zip(weights, match, total):
            p = (int(m) / int(t)) if m != 0 else sys.float_info.min
            if p > 0:
                terms.append(w * math.log(p))
        out[r] = bp * math.exp(math.fsum(terms))
    return out


class BleuScore:
    """utlis/tools.py:30-43."""

    def __init__(self, w1, w2, w3, w4):
        self.w1, self.w2, self.w3, self.w4 = w1, w2, w3, w4

    @property
    def weights(self):
        return (self.w1, self.w2, self.w3, self.w4)

    @staticmethod
    def counts_from_ids(real_ids: torch.Tensor, predicted_ids: torch.Tensor) -> torch.Tensor:
        """[n, L] int32 CUDA tensors -> [n, 10] int32 counts (on device)."""
        return _lib.bleu_counts(real_ids.to(torch.int32).contiguous(), predicted_ids.to(torch.int32).contiguous())

    def score_from_ids(self, real_ids: torch.Tensor, predicted_ids: torch.Tensor) -> List[float]:
        return score_from_counts(self.counts_from_ids(real_ids, predicted_ids), self.weights).tolist()

    def compute_score(self, real: Sequence[str], predicted: Sequence[str], device="cuda") -> List[float]:
        """String interface of the reference: remove tags, split, score.  Tokens are mapped to a
        per-call integer alphabet (ids >= 5) so the n-gram kernel can be used for arbitrary words."""
        alphabet: Dict[str, int] = {}
        rows_r, rows_p = [], []
        for s1, s2 in zip(real, predicted):
            t1, t2 = _RE_TAGS.sub('', s1).split(), _RE_TAGS.sub('', s2).split()
            if len(t1) > 32 or len(t2) > 32:
                raise ValueError("BleuScore kernel handles sentences of at most 32 tokens")
            rows_r.append([alphabet.setdefault(w, 5 + len(alphabet)) for w in t1] + [0] * (32 - len(t1)))
            rows_p.append([alphabet.setdefault(w, 5 + len(alphabet)) for w in t2] + [0] * (32 - len(t2)))
        if not rows_r:
            return []
        r = torch.tensor(rows_r, dtype=torch.int32, device=device)
        p = torch.tensor(rows_p, dtype=torch.int32, device=device)
        return score_from_counts(_lib.bleu_counts(r, p), self.weights).tolist()


def SNR_to_noise(snr):
    """utlis/tools.py:46-50."""
    snr = 10 ** (snr / 10)
    noise_std = 1 / np.sqrt(snr)
    return noise_std


def similarity_from_features(vector1, vector2) -> List[float]:
    """The arithmetic of ``Similarity.compute_score`` after the two BERT forwards (utlis/tools.py:84-103):
    features [n, positions, hidden] are summed over positions, each feature COLUMN is divided by its maximum absolute
    value over the batch (sklearn ``normalize(axis=0, norm='max')``: the score of a sentence depends on the batch it is
    scored in), and the row-wise cosine is returned."""
    v1 = np.asarray(vector1, dtype=np.float64).sum(axis=1)
    v2 = np.asarray(vector2, dtype=np.float64).sum(axis=1)

    def _max_norm_columns(v):
        m = np.abs(v).max(axis=0)
        m[m == 0.0] = 1.0                                          # sklearn leaves all-zero columns untouched
        return v / m

    v1, v2 = _max_norm_columns(v1), _max_norm_columns(v2)
    dot = np.einsum("ij,ij->i", v1, v2)
    return (dot / (np.sqrt(np.einsum("ij,ij->i", v1, v1)) * np.sqrt(np.einsum("ij,ij->i", v2, v2)))).tolist()


class Similarity:
    """utlis/tools.py:53-103.  The reference builds a bert4keras model from (config_path, checkpoint_path, dict_path) and
    takes the output of layer 'Encoder-11-FeedForward-Norm'; neither bert4keras nor a checkpoint exists in this image or in
    the reference tree, so the encoder is a callable the caller supplies:

        encoder(list_of_sentences) -> float array [n, 32, hidden]    (tokenise, pad to 32 post, run BERT)

    ``compute_score`` then follows the reference: tags removed from both sentences, both batches encoded, scored with
    ``similarity_from_features``.  Without an encoder the constructor raises - there is no silent substitute metric."""

    def __init__(self, config_path=None, checkpoint_path=None, dict_path=None, *, encoder=None):
        if encoder is None:
            raise RuntimeError("Similarity needs a BERT sentence encoder (bert4keras + checkpoint are not available): "
                               "pass encoder=callable(sentences) -> [n, 32, hidden] features")
        self.config_path, self.checkpoint_path, self.dict_path, self.encoder = config_path, checkpoint_path, dict_path, encoder

    def compute_score(self, real: Sequence[str], predicted: Sequence[str]) -> List[float]:
        real = [_RE_TAGS.sub('', s) for s in real]
        predicted = [_RE_TAGS.sub('', s) for s in predicted]
        return similarity_from_features(self.encoder(real), self.encoder(predicted))
