"""Real-data fixture: the reference's whole test set and vocabulary, for the tests and the real-data SNR sweep.

Run in the build container (reads /root/reference):  python tests/golden/make_europarl_fixture.py
Writes tests/golden/europarl_test.npz: ``ids`` uint16 [7347, 31] = DeepSC-GAN/data/txt/test_data.pkl padded post with 0
to 31 as dataset/dataloader.py:11 does, ``lengths`` uint8 [7347], and ``tokens`` = the 22,234 tokens of
DeepSC-GAN/data/txt/vocab.json in id order (token_to_idx[tokens[i]] == i), newline-joined UTF-8.
Data fixture only - no reference source is copied.
"""
import json
import os
import pickle

import numpy as np

REF = "/root/reference/DeepSC-GAN/data/txt"


def main():
    data = pickle.load(open(os.path.join(REF, "test_data.pkl"), "rb"))
    vocab = json.load(open(os.path.join(REF, "vocab.json")))["token_to_idx"]
    ids = np.zeros((len(data), 31), dtype=np.uint16)
    for r, s in enumerate(data):
        assert len(s) <= 31 and max(s) < 65536
        ids[r, : len(s)] = s
    tokens = [None] * len(vocab)
    for tok, i in vocab.items():
        assert tokens[i] is None and "\n" not in tok
        tokens[i] = tok
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "europarl_test.npz")
    np.savez_compressed(dst, ids=ids, lengths=np.array([len(s) for s in data], dtype=np.uint8),
                        tokens=np.frombuffer("\n".join(tokens).encode("utf-8"), dtype=np.uint8))
    print("wrote", dst, os.path.getsize(dst), "bytes:", ids.shape, len(tokens), "tokens")


if __name__ == "__main__":
    main()
