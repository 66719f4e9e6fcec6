"""Text preprocessing: the tokeniser / vocabulary builder that produces the ``train_data.pkl`` / ``test_data.pkl`` /
``vocab.json`` the data loader and the BLEU path consume.  Mirror of DeepSC-GAN/dataset/preprocess_text.py (same
function names, argument meaning and file formats); host-side Python, no device work.

    python -m deepsc_gan_b200.dataset.preprocess_text --data-dir data/ --input-data-dir txt/en

Behaviour that downstream code depends on and that is pinned by tests/test_preprocess.py against the reference's own
``vocab.json`` / ``test_data.pkl``:

* ids 0..3 are the special tokens; every other token gets its id in SORTED token order (build_vocab, ref :90-106), so the
  reference vocabulary satisfies ``tokens[4:] == sorted(tokens[4:])``;
* the sentence-final ``.`` / ``?`` are deleted AFTER normalisation put a blank in front of them (ref :31, :140), so a
  sentence that ended in a full stop ends in the EMPTY token ``''`` (id 4 in the reference vocabulary); ``!`` survives as
  a token of its own (id 5);
* ``cutted_data`` keeps sentences of 5..29 blank-separated words (strict inequalities, ref :42-46), counted before the
  start / end tokens are added: encoded sentences have 7..31 ids, the ``maxlen = 31`` of dataset/dataloader.py:11.
"""
from __future__ import annotations

import argparse
import json
import os
import pickle
import re
import unicodedata
from typing import Dict, Iterable, List, Optional, Sequence

SPECIAL_TOKENS = {'<PAD>': 0, '<START>': 1, '<END>': 2, '<UNK>': 3}

# w3lib.html.remove_tags with its default arguments (w3lib is not a dependency here): every <tag ...>, </tag> and
# <!...> is deleted, text between tags stays
_TAG = re.compile(r"<[a-zA-Z\/!].*?>", re.DOTALL | re.IGNORECASE)
_PUNCT = re.compile(r"([!.?])")
_NOT_KEPT = re.compile(r"[^a-zA-Z.!?]+")
_BLANKS = re.compile(r"\s+")


def remove_tags(text: str) -> str:
    return _TAG.sub("", text)


def unicode_to_ascii(s: str) -> str:
    """ref :23-25: NFD-decompose and drop the combining marks (category Mn)."""
    return "".join(ch for ch in unicodedata.normalize("NFD", s) if unicodedata.category(ch) != "Mn")


def normalize_string(s: str) -> str:
    """ref :27-38: strip accents and XML tags, put a blank before ``!`` ``.`` ``?``, replace every run of other
    non-letters by one blank, squeeze blanks, lower-case."""
    s = remove_tags(unicode_to_ascii(s))
    s = _PUNCT.sub(r" \1", s)
    s = _NOT_KEPT.sub(" ", s)
    return _BLANKS.sub(" ", s).lower()


def cutted_data(cleaned: Iterable[str], MIN_LENGTH: int = 4, MAX_LENGTH: int = 30) -> List[str]:
    """ref :40-47: keep lines with MIN_LENGTH < words < MAX_LENGTH, re-joined with single blanks."""
    kept = []
    for line in cleaned:
        words = line.split()
        if MIN_LENGTH < len(words) < MAX_LENGTH:
            kept.append(" ".join(words))
    return kept


def save_clean_sentences(sentence, save_path: str) -> None:
    with open(save_path, "wb") as f:
        pickle.dump(sentence, f)
    print("Saved: %s" % save_path)


def process(text_path: str) -> List[str]:
    """ref :53-61: one sentence per line of a UTF-8 text file -> normalised, length-filtered sentences."""
    with open(text_path, "r", encoding="utf8") as f:
        lines = f.read().strip().split("\n")
    return cutted_data(normalize_string(line) for line in lines)


def tokenize(s: str, delim: str = " ", add_start_token: bool = True, add_end_token: bool = True,
             punct_to_keep: Optional[Sequence[str]] = None, punct_to_remove: Optional[Sequence[str]] = None) -> List[str]:
    """ref :64-85: split on ``delim`` (NOT on runs of blanks: consecutive delimiters yield empty tokens), after giving
    every ``punct_to_keep`` mark a delimiter in front and deleting every ``punct_to_remove`` mark."""
    for mark in punct_to_keep or ():
        s = s.replace(mark, delim + mark)
    for mark in punct_to_remove or ():
        s = s.replace(mark, "")
    tokens = s.split(delim)
    return (["<START>"] if add_start_token else []) + tokens + (["<END>"] if add_end_token else [])


def build_vocab(sequences: Iterable[str], token_to_idx: Optional[Dict[str, int]] = None, min_token_count: int = 1,
                delim: str = " ", punct_to_keep=None, punct_to_remove=None) -> Dict[str, int]:
    """ref :88-106: count tokens over all sequences; tokens seen at least ``min_token_count`` times are appended to
    ``token_to_idx`` in sorted order.  The reference's mutable default ``{}`` is replaced by a fresh dict per call; a
    dict that is passed in is extended in place and returned, as there."""
    if token_to_idx is None:
        token_to_idx = {}
    counts: Dict[str, int] = {}
    for seq in sequences:
        for tok in tokenize(seq, delim=delim, punct_to_keep=punct_to_keep, punct_to_remove=punct_to_remove,
                            add_start_token=False, add_end_token=False):
            counts[tok] = counts.get(tok, 0) + 1
    for tok in sorted(counts):
        # a token already present keeps being re-assigned len(token_to_idx) in the reference (:104), which for the
        # special tokens never happens (they contain '<'); first assignment wins here
        if counts[tok] >= min_token_count and tok not in token_to_idx:
            token_to_idx[tok] = len(token_to_idx)
    return token_to_idx


def encode(seq_tokens: Iterable[str], token_to_idx: Dict[str, int], allow_unk: bool = False) -> List[int]:
    """ref :109-118: KeyError on an unknown token unless ``allow_unk``."""
    out = []
    for tok in seq_tokens:
        if tok not in token_to_idx:
            if not allow_unk:
                raise KeyError('Token "%s" not in vocab' % tok)
            tok = "<UNK>"
        out.append(token_to_idx[tok])
    return out


def decode(seq_idx: Iterable[int], idx_to_token, delim: Optional[str] = None, stop_at_end: bool = True):
    """ref :121-130: ids -> tokens, the ``<END>`` token included when ``stop_at_end`` cuts there."""
    tokens = []
    for idx in seq_idx:
        tokens.append(idx_to_token[idx])
        if stop_at_end and tokens[-1] == "<END>":
            break
    return tokens if delim is None else delim.join(tokens)


PUNCT_TO_KEEP, PUNCT_TO_REMOVE = (";", ","), ("?", ".")


def encode_corpus(sentences: Sequence[str], token_to_idx: Dict[str, int]) -> List[List[int]]:
    """ref :167-172: start / end tokens added, the same punctuation rules as the vocabulary."""
    return [encode(tokenize(s, punct_to_keep=PUNCT_TO_KEEP, punct_to_remove=PUNCT_TO_REMOVE), token_to_idx) for s in sentences]


def split_train_test(results: Sequence, train_fraction: float = 0.9):
    """ref :176-177: the first round(0.9 n) sentences train, the rest test (file order, no shuffle)."""
    cut = round(len(results) * train_fraction)
    return list(results[:cut]), list(results[cut:])


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser()
    p.add_argument("--input-data-dir", default="txt/en", type=str)
    p.add_argument("--output-train-dir", default="txt/train_data.pkl", type=str)
    p.add_argument("--output-test-dir", default="txt/test_data.pkl", type=str)
    p.add_argument("--output-vocab", default="txt/vocab.json", type=str)
    p.add_argument("--data-dir", default="data/", type=str,
                   help="prefix of the four paths (the reference hard-codes a Windows path at :134)")
    return p


def main(args) -> Dict[str, int]:
    """ref :133-182: every ``*.txt`` under the input directory -> de-duplicated sentences (first occurrence order) ->
    vocab.json ({'token_to_idx': ...}) and the two pickles of id lists."""
    root = getattr(args, "data_dir", "")
    in_dir = os.path.join(root, args.input_data_dir)
    sentences: List[str] = []
    for fn in os.listdir(in_dir):
        if fn.endswith(".txt"):
            sentences += process(os.path.join(in_dir, fn))
    sentences = list(dict.fromkeys(sentences))                     # remove repeated sentences, keep first-seen order
    print("Number of sentences: {}".format(len(sentences)))
    token_to_idx = build_vocab(sentences, dict(SPECIAL_TOKENS), punct_to_keep=PUNCT_TO_KEEP, punct_to_remove=PUNCT_TO_REMOVE)
    print("Number of words in Vocab: {}".format(len(token_to_idx)))
    if args.output_vocab != "":
        with open(os.path.join(root, args.output_vocab), "w") as f:
            json.dump({"token_to_idx": token_to_idx}, f)
    train, test = split_train_test(encode_corpus(sentences, token_to_idx))
    with open(os.path.join(root, args.output_train_dir), "wb") as f:
        pickle.dump(train, f)
    with open(os.path.join(root, args.output_test_dir), "wb") as f:
        pickle.dump(test, f)
    return token_to_idx


if __name__ == "__main__":
    main(build_parser().parse_args())
