"""the tokeniser / vocabulary builder (mirror of DeepSC-GAN/dataset/preprocess_text.py) and the
Similarity scoring arithmetic (utlis/tools.py:84-103), pinned against the reference's own artefacts: its vocab.json and
test_data.pkl (tests/golden/europarl_test.npz, made by make_europarl_fixture.py)."""
import json
import os
import pickle
import types

import numpy as np
import pytest

import deepsc_gan_b200  # noqa: F401
from deepsc_gan_b200.dataset import preprocess_text as PT
from deepsc_gan_b200.utlis import tools

import _cases


def test_normalize_string_known_answers():
    assert PT.normalize_string("Résumé of the <b>Session</b>, 1999!") == "resume of the session !"
    assert PT.normalize_string("Is it so? Yes.  It is.") == "is it so ? yes . it is ."
    assert PT.normalize_string("a-b  c\td") == "a b c d"
    assert PT.normalize_string("<!-- note -->kept</p>") == "kept"
    assert PT.unicode_to_ascii("Ångström façade") == "Angstrom facade"


def test_cutted_data_bounds_are_strict():
    lines = [" ".join(["w"] * n) for n in (4, 5, 29, 30)]
    assert [len(s.split()) for s in PT.cutted_data(lines)] == [5, 29]
    assert PT.cutted_data(["  a   b c  d e  "]) == ["a b c d e"]


def test_tokenize_splits_on_the_delimiter_not_on_blank_runs():
    # the full stop is removed after normalisation put a blank in front of it: the sentence ends in the EMPTY token
    assert PT.tokenize("the house rose .", punct_to_keep=[";", ","], punct_to_remove=["?", "."]) == \
        ["<START>", "the", "house", "rose", "", "<END>"]
    assert PT.tokenize("why not ?", add_start_token=False, add_end_token=False, punct_to_remove=["?", "."]) == ["why", "not", ""]
    assert PT.tokenize("go !") == ["<START>", "go", "!", "<END>"]
    assert PT.tokenize("a,b", punct_to_keep=[","], add_start_token=False, add_end_token=False) == ["a", ",b"]


def test_build_vocab_sorted_ids_and_min_count():
    v = PT.build_vocab(["b a .", "c a ."], dict(PT.SPECIAL_TOKENS), punct_to_remove=["?", "."])
    assert v == {"<PAD>": 0, "<START>": 1, "<END>": 2, "<UNK>": 3, "": 4, "a": 5, "b": 6, "c": 7}
    v2 = PT.build_vocab(["b a", "c a"], min_token_count=2)
    assert v2 == {"a": 0}
    assert PT.build_vocab(["x"]) == {"x": 0}, "no state may leak between calls (the reference's mutable default)"


def test_encode_decode():
    v = {"<PAD>": 0, "<START>": 1, "<END>": 2, "<UNK>": 3, "a": 4}
    assert PT.encode(["<START>", "a", "zzz", "<END>"], v, allow_unk=True) == [1, 4, 3, 2]
    with pytest.raises(KeyError):
        PT.encode(["zzz"], v)
    inv = {i: t for t, i in v.items()}
    assert PT.decode([1, 4, 2, 4], inv) == ["<START>", "a", "<END>"]
    assert PT.decode([1, 4, 2, 4], inv, delim=" ", stop_at_end=False) == "<START> a <END> a"


def test_reference_vocabulary_has_the_builder_s_order():
    """The reference's vocab.json is what build_vocab produces: special tokens 0..3, then every token in sorted order,
    the empty token (the deleted full stop) first."""
    _, vocab = _cases.europarl_test()
    tokens = sorted(vocab, key=vocab.get)
    assert tokens[:4] == ["<PAD>", "<START>", "<END>", "<UNK>"] and tokens[4] == "" and tokens[5] == "!"
    assert tokens[4:] == sorted(tokens[4:])
    assert all(PT.normalize_string(t) == t for t in tokens[5:] if t != "!"), "every token is a fixed point of the normaliser"


def test_round_trip_on_the_reference_test_set():
    """ids of test_data.pkl -> tokens -> sentence text -> tokenize + encode gives the ids back for all 7,347 sentences, and
    a vocabulary rebuilt from those sentences maps onto the reference's ids monotonically (same sorted-order rule)."""
    ids, vocab = _cases.europarl_test()
    inv = {i: t for t, i in vocab.items()}
    sentences = []
    for row in ids:
        toks = PT.decode([int(i) for i in row], inv)
        assert toks[0] == "<START>" and toks[-1] == "<END>"
        # cutted_data's bounds, counted on the normalised line where the (later deleted) '.' / '?' are words of their own
        assert 4 < len(toks) - 2 < 30
        sentences.append(" ".join(toks[1:-1]))
    enc = PT.encode_corpus(sentences, vocab)
    for row, e in zip(ids, enc):
        assert e == [int(i) for i in row[: len(e)]] and not row[len(e):].any()
    rebuilt = PT.build_vocab(sentences, dict(PT.SPECIAL_TOKENS), punct_to_keep=PT.PUNCT_TO_KEEP, punct_to_remove=PT.PUNCT_TO_REMOVE)
    common = [t for t in sorted(rebuilt, key=rebuilt.get)]
    ref_ids = [vocab[t] for t in common]
    assert ref_ids == sorted(ref_ids), "sorted-order rule: the rebuilt ids are a monotone relabelling of the reference's"


def test_main_writes_the_three_artefacts(tmp_path):
    src = tmp_path / "txt" / "en"
    src.mkdir(parents=True)
    lines = ["The <i>quick</i> brown fox jumps over it.", "Too short.", "The quick brown fox jumps over it.",
             "Déjà vu: we have seen all of this before!", "What do we do about the year two thousand?",
             "One two three four five six seven.", "Alpha beta gamma delta epsilon zeta.", "Nine lives has the cat in the hat.",
             "Every good boy deserves fun and more.", "All cows eat grass in the green field.", "My very eager mother just served us."]
    (src / "ep.txt").write_text("\n".join(lines), encoding="utf8")
    (src / "ignored.dat").write_text("not read")
    args = PT.build_parser().parse_args(["--data-dir", str(tmp_path) + os.sep])
    vocab = PT.main(args)
    on_disk = json.load(open(tmp_path / "txt" / "vocab.json"))["token_to_idx"]
    assert on_disk == vocab and list(vocab)[:5] == ["<PAD>", "<START>", "<END>", "<UNK>", ""]
    train = pickle.load(open(tmp_path / "txt" / "train_data.pkl", "rb"))
    test = pickle.load(open(tmp_path / "txt" / "test_data.pkl", "rb"))
    assert len(train) + len(test) == 9 and len(train) == round(9 * 0.9)          # duplicate and short line dropped
    inv = {i: t for t, i in vocab.items()}
    assert PT.decode(train[0], inv, delim=" ") == "<START> the quick brown fox jumps over it  <END>"
    assert PT.decode(train[1], inv, delim=" ") == "<START> deja vu we have seen all of this before ! <END>"
    # the data loader reads what the tokeniser wrote
    from deepsc_gan_b200.dataset import dataloader
    ds = dataloader.return_dataset(types.SimpleNamespace(bs=4), str(tmp_path / "txt" / "train_data.pkl"), -1, shuffle=False)
    assert ds.data.shape == (len(train) - 1, 31) and int(ds.data[0, 0]) == 1


def test_similarity_arithmetic_matches_the_reference_formula():
    rng = np.random.default_rng(3)
    f1, f2 = rng.standard_normal((5, 32, 24)), rng.standard_normal((5, 32, 24))
    got = tools.similarity_from_features(f1, f2)
    # the reference's lines, literally: sum over axis 1, max-normalise columns, diag of the Gram matrices
    v1, v2 = f1.sum(1), f2.sum(1)
    v1, v2 = v1 / np.abs(v1).max(0), v2 / np.abs(v2).max(0)
    want = np.diag(v1 @ v2.T) / (np.sqrt(np.diag(v1 @ v1.T)) * np.sqrt(np.diag(v2 @ v2.T)))
    np.testing.assert_allclose(got, want, rtol=1e-12)
    assert tools.similarity_from_features(f1, f1) == pytest.approx([1.0] * 5)


def test_similarity_needs_an_encoder_and_strips_tags():
    with pytest.raises(RuntimeError):
        tools.Similarity("cfg", "ckpt", "dict")
    seen = []

    def encoder(sentences):
        seen.append(list(sentences))
        return np.stack([np.full((32, 4), float(len(s))) + np.arange(4) for s in sentences])

    sim = tools.Similarity(encoder=encoder)
    out = sim.compute_score(["<START> a b", "c"], ["a b", "<b>c</b> d"])
    assert seen == [[" a b", "c"], ["a b", "c d"]] and len(out) == 2 and all(0 < v <= 1 for v in out)
