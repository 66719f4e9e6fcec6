"""The integer-domain BLEU contract against the string-domain restatement of the reference's text path
(SeqtoText -> remove_tags -> split -> nltk sentence_bleu), on real Europarl ids from the reference's
own test_data.pkl / vocab.json (fixture: tests/golden/europarl_sample.json)."""
import json
import os
import random

import numpy as np

import _cases
from oracle import bleu_oracle as B

FIX = json.load(open(os.path.join(_cases.GOLDEN_DIR, "europarl_sample.json")))
REV = {v: k for k, v in FIX["token_to_idx"].items()}
PRESENT = sorted(i for i in REV if i >= 5)


def corrupt(seq, rng):
    """A noisy 'decoded' sentence: substitutions, deletions, repeats, early END, PAD/UNK intrusions."""
    out = []
    for t in seq:
        r = rng.random()
        if r < 0.12:
            out.append(rng.choice(PRESENT))
        elif r < 0.18:
            continue
        elif r < 0.24:
            out.extend([t, t])
        elif r < 0.27:
            out.append(rng.choice([0, 3, 4]))
        else:
            out.append(t)
    out = out[:31]
    return out + [0] * (31 - len(out))


def pad(seq):
    return list(seq) + [0] * (31 - len(seq))


def test_fixture_shape():
    assert len(FIX["sentences"]) == 192
    assert all(7 <= len(s) <= 31 and s[0] == 1 and s[-1] == 2 for s in FIX["sentences"])
    assert REV[4] == "" and REV[1] == "<START>"


def test_integer_counts_equal_string_path_on_real_sentences():
    rng = random.Random(5)
    for s in FIX["sentences"]:
        hyp = corrupt(s, rng)
        score_s, counts_s = B.string_bleu(pad(s), hyp, REV)
        counts_i = B.bleu_counts_one(pad(s), hyp)
        assert counts_i == counts_s
        assert abs(B.sentence_bleu_from_counts(counts_i) - score_s) < 1e-15


def test_known_answers():
    ref = [1, 10, 11, 12, 13, 14, 4, 2]
    assert B.bleu_counts_one(ref, ref) == [5, 4, 3, 2, 5, 4, 3, 2, 5, 5]
    assert B.sentence_bleu_from_counts(B.bleu_counts_one(ref, ref)) == 1.0
    # clipping: hypothesis repeats a word that the reference has once
    assert B.bleu_counts_one([1, 10, 11, 2], [1, 10, 10, 10, 2])[:8] == [1, 0, 0, 0, 3, 2, 1, 1]
    # empty hypothesis (END first) and hypothesis shorter than n
    assert B.bleu_counts_one(ref, [2, 10, 11]) == [0, 0, 0, 0, 1, 1, 1, 1, 0, 5]
    assert B.sentence_bleu_from_counts([0, 0, 0, 0, 1, 1, 1, 1, 0, 5]) == 0.0
    c = B.bleu_counts_one(ref, [1, 10, 11, 2])
    assert c == [2, 1, 0, 0, 2, 1, 1, 1, 2, 5]
    bp = np.exp(1 - 5 / 2)
    assert abs(B.sentence_bleu_from_counts(c, (1, 0, 0, 0)) - bp * 1.0) < 1e-12
