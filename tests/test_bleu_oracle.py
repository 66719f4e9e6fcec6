"""The integer-domain BLEU contract against the string-domain restatement of the reference's text path
(SeqtoText -> remove_tags -> split -> nltk sentence_bleu), on real Europarl ids from the reference's
own test_data.pkl / vocab.json (fixture: tests/golden/europarl_sample.json)."""
import json
import os
import random

import numpy as np
import pytest

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


# ------------------------------------------------------------------ the whole reference test set and vocabulary
def test_whole_test_set_integer_counts_equal_string_path():
    """All 7,347 sentences of the reference's test_data.pkl with the full 22,234-token vocab.json: the integer-domain
    counts (the kernel's contract) equal the counts of the string path SeqtoText -> remove_tags -> split -> nltk
    modified precision, for a corrupted hypothesis of every sentence and for the sentence against itself."""
    ids, vocab = _cases.europarl_test()
    assert ids.shape == (7347, 31) and len(vocab) == 22234
    rev = {i: t for t, i in vocab.items()}
    present = sorted(set(int(t) for t in ids.ravel()) - {0, 1, 2, 3, 4})
    rng = random.Random(11)

    def corrupt_full(seq):
        out = []
        for t in seq:
            r = rng.random()
            if r < 0.12:
                out.append(rng.choice(present))
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

    for row in ids:
        s = [int(t) for t in row]
        hyp = corrupt_full(s)
        score_s, counts_s = B.string_bleu(s, hyp, rev)
        counts_i = B.bleu_counts_one(s, hyp)
        assert counts_i == counts_s
        assert abs(B.sentence_bleu_from_counts(counts_i) - score_s) < 1e-15
        self_counts = B.bleu_counts_one(s, s)
        assert self_counts == B.string_bleu(s, s, rev)[1] and self_counts[8] == self_counts[9]


def test_remove_tags_never_spans_two_tokens_on_the_reference_vocabulary():
    """w3lib's remove_tags regex works on the JOINED sentence; it could swallow real words only if some token other than
    the four specials held a '<' or '>'.  None of the 22,234 does, so on this vocabulary the string path drops exactly
    <PAD>/<START>/<UNK> (and ``split`` drops the empty token, id 4): the id-domain clean() is the same map."""
    ids, vocab = _cases.europarl_test()
    rev = {i: t for t, i in vocab.items()}
    assert sorted(t for t in vocab if "<" in t or ">" in t) == ["<END>", "<PAD>", "<START>", "<UNK>"]
    assert not any(any(ch.isspace() for ch in t) for t in vocab) and rev[4] == ""
    for row in ids[::7]:
        s = [int(t) for t in row]
        text = B.sequence_to_text(s, rev)
        assert B.remove_tags(text).split() == [rev[t] for t in B.clean_ids(s)]


def test_product_seqtotext_equals_the_oracle_and_round_trips():
    """deepsc_gan_b200.utlis.tools.SeqtoText (the mirror of utlis/tools.py:10-27) on every sentence of the test set."""
    from deepsc_gan_b200.utlis.tools import SeqtoText
    ids, vocab = _cases.europarl_test()
    rev = {i: t for t, i in vocab.items()}
    st = SeqtoText(vocab, 2)
    assert st.reverse_word_map[1] == "<START>" and st.end_idx == 2
    for row in ids:
        s = [int(t) for t in row]
        text = st.sequence_to_text(s)
        assert text == B.sequence_to_text(s, rev)
        assert "<END>" not in text and text.startswith("<START>")
        back = st.text_to_sequence(text)                      # split() drops the empty token (id 4), nothing else
        assert back == [t for t in s[: s.index(2)] if t != 4]
    assert st.sequence_to_text([1, 2, 17]) == "<START>" and st.sequence_to_text([]) == ""
    with pytest.raises(TypeError):                            # an id outside the vocabulary: ' '.join([..., None]), as in the reference
        st.sequence_to_text([1, 10 ** 6, 2])
