"""Real-data fixture for the BLEU/text tail: the first 192 sentences of the reference's
data/txt/test_data.pkl plus the slice of data/txt/vocab.json they use (and the special tokens).

Run in the build container (reads /root/reference):  python tests/golden/make_bleu_fixture.py
Writes tests/golden/europarl_sample.json.  Data fixture only - no reference source is copied.
"""
import json
import os
import pickle

REF = "/root/reference/DeepSC-GAN/data/txt"
N = 192


def main():
    data = pickle.load(open(os.path.join(REF, "test_data.pkl"), "rb"))[:N]
    vocab = json.load(open(os.path.join(REF, "vocab.json")))["token_to_idx"]
    used = {0, 1, 2, 3, 4}
    for s in data:
        used.update(int(t) for t in s)
    sub = {tok: idx for tok, idx in vocab.items() if idx in used}
    out = {"source": "DeepSC-GAN/data/txt/test_data.pkl[:192] + vocab.json subset", "vocab_size": len(vocab),
           "sentences": [[int(t) for t in s] for s in data], "token_to_idx": sub}
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "europarl_sample.json")
    json.dump(out, open(dst, "w"))
    print("wrote", dst, len(data), "sentences", len(sub), "tokens")


if __name__ == "__main__":
    main()
