"""CPU oracle for the text/BLEU tail of the eval loop.  TEST INFRASTRUCTURE ONLY.

Two independent restatements of utlis/tools.py:10-43:

* ``string_bleu`` walks the reference's *string* path: SeqtoText.sequence_to_text
  (:15-24) -> w3lib ``remove_tags`` -> ``str.split`` -> nltk ``sentence_bleu`` with the
  default (method0) smoothing.  nltk and w3lib are third-party, un-vendored and un-pinned
  by the reference; their published algorithms are restated below (nltk 3.x
  ``translate/bleu_score.py``: modified_precision, closest_ref_length, brevity_penalty,
  corpus_bleu; w3lib ``html.remove_tags`` regex ``<[a-zA-Z\\/!].*?>``).
* ``bleu_counts`` is the integer-domain contract the CUDA kernel implements
  (SURVEY.md App. D): int32[10] = [match_1..4, total_1..4, hyp_len, ref_len].

``tests/test_bleu_oracle.py`` checks that the two agree on real Europarl ids.
PARITY UNPINNED for the float score (older nltk releases differ on the zero-match branch);
the integer counts are the contract.
"""
from __future__ import annotations

import math
import re
import sys
from collections import Counter
from fractions import Fraction
from typing import Dict, List, Sequence, Tuple

import numpy as np

PAD, START, END, UNK, EMPTY = 0, 1, 2, 3, 4
_DROP = (PAD, START, UNK, EMPTY)

_RE_TAGS = re.compile(r"<[a-zA-Z\/!].*?>", re.DOTALL | re.IGNORECASE)


# ----------------------------- string domain -------------------------------- #
def sequence_to_text(ids: Sequence[int], reverse_word_map: Dict[int, str], end_idx: int = END) -> str:
    """SeqtoText.sequence_to_text, utlis/tools.py:15-24."""
    words = []
    for idx in ids:
        if idx == end_idx:
            break
        words.append(reverse_word_map.get(int(idx)))
    return " ".join(words)


def remove_tags(text: str) -> str:
    """w3lib.html.remove_tags with no which_ones/keep: drop everything shaped like a tag."""
    return _RE_TAGS.sub("", text)


def _ngrams(tokens: Sequence, n: int):
    return [tuple(tokens[i:i + n]) for i in range(len(tokens) - n + 1)]


def _modified_precision(reference: Sequence, hypothesis: Sequence, n: int) -> Tuple[int, int]:
    counts = Counter(_ngrams(hypothesis, n)) if len(hypothesis) >= n else Counter()
    ref_counts = Counter(_ngrams(reference, n)) if len(reference) >= n else Counter()
    clipped = {g: min(c, ref_counts[g]) for g, c in counts.items()}
    return sum(clipped.values()), max(1, sum(counts.values()))


def sentence_bleu_from_counts(counts: Sequence[int], weights=(0.25, 0.25, 0.25, 0.25)) -> float:
    """nltk corpus_bleu for one (reference, hypothesis) pair, from the integer counts."""
    match, total = counts[0:4], counts[4:8]
    hyp_len, ref_len = int(counts[8]), int(counts[9])
    if match[0] == 0:
        return 0.0
    if hyp_len > ref_len:
        bp = 1.0
    elif hyp_len == 0:
        bp = 0.0
    else:
        bp = math.exp(1 - ref_len / hyp_len)
    p_n = []
    for m, t in zip(match, total):
        p_n.append(Fraction(int(m), int(t)) if m != 0 else sys.float_info.min)   # method0
    s = (w * math.log(p) for w, p in zip(weights, p_n) if p > 0)
    return bp * math.exp(math.fsum(s))


def string_bleu(real_ids: Sequence[int], pred_ids: Sequence[int], reverse_word_map: Dict[int, str],
                weights=(0.25, 0.25, 0.25, 0.25)) -> Tuple[float, List[int]]:
    """BleuScore.compute_score for one pair, utlis/tools.py:37-43, through the string path.
    Returns (score, counts[10]) where the counts are taken from the token lists."""
    ref = remove_tags(sequence_to_text(real_ids, reverse_word_map)).split()
    hyp = remove_tags(sequence_to_text(pred_ids, reverse_word_map)).split()
    counts = []
    totals = []
    for n in range(1, 5):
        m, t = _modified_precision(ref, hyp, n)
        counts.append(m)
        totals.append(t)
    c = counts + totals + [len(hyp), len(ref)]
    return sentence_bleu_from_counts(c, weights), c


# ----------------------------- integer domain ------------------------------- #
def clean_ids(seq: Sequence[int]) -> List[int]:
    out = []
    for t in seq:
        t = int(t)
        if t == END:
            break
        if t in _DROP:
            continue
        out.append(t)
    return out


def bleu_counts_one(ref_ids: Sequence[int], hyp_ids: Sequence[int]) -> List[int]:
    ref, hyp = clean_ids(ref_ids), clean_ids(hyp_ids)
    match, total = [], []
    for n in range(1, 5):
        m, t = _modified_precision(ref, hyp, n)
        match.append(m)
        total.append(t)
    return match + total + [len(hyp), len(ref)]


def bleu_counts(ref_ids: np.ndarray, hyp_ids: np.ndarray) -> np.ndarray:
    """[N, Lr] , [N, Lh] int -> [N, 10] int32."""
    return np.asarray([bleu_counts_one(r, h) for r, h in zip(ref_ids, hyp_ids)], dtype=np.int32)


def bleu_scores(counts: np.ndarray, weights=(0.25, 0.25, 0.25, 0.25)) -> np.ndarray:
    return np.asarray([sentence_bleu_from_counts(c, weights) for c in counts], dtype=np.float64)
