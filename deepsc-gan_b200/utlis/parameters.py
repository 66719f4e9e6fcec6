"""Flag set of the reference (DeepSC-GAN/utlis/parameters.py:5-65): same flag names and defaults.

The reference's missing driver also sets ``vocab_size``, ``start_idx`` and ``end_idx`` on the namespace
(they are read at models/transceiver.py:124 and utlis/eval.py:13 but never parsed); they are flags here
with the values fixed by dataset/preprocess_text.py:17-22 and data/txt/vocab.json (22,234 tokens).
The Windows path defaults of the reference are replaced by relative ones.
"""
import argparse


def para_config(argv=None):
    parser = argparse.ArgumentParser()
    # preprocessing parameters
    parser.add_argument('--input-data-dir', default='txt/en', type=str)
    parser.add_argument('--output-train-dir', default='txt/train_data.pkl', type=str)
    parser.add_argument('--output-test-dir', default='txt/test_data.pkl', type=str)
    parser.add_argument('--output-vocab', default='txt/vocab.json', type=str)
    parser.add_argument('--log-save-path', default='log', type=str)
    parser.add_argument('--train-save-path', default='data/txt/train_data.pkl', type=str)
    parser.add_argument('--test-save-path', default='data/txt/test_data.pkl', type=str)
    parser.add_argument('--vocab-path', default='data/txt/vocab.json', type=str)
    # training parameters
    parser.add_argument('--bs', default=64, type=int, help='The training batch size')
    parser.add_argument('--shuffle-size', default=22234, type=int, help='The training shuffle size')
    parser.add_argument('--lr', default=5e-4, type=float, help='The training learning rate')
    parser.add_argument('--epochs', default=60, type=int, help='The training number of epochs')
    parser.add_argument('--train-with-mine', action='store_true')
    parser.add_argument('--checkpoint-path', default='checkpoint', type=str)
    parser.add_argument('--max-length', default=30, type=int)
    parser.add_argument('--channel', default='AWGN', type=str, help='Choose the channel to simulate')
    # model parameters
    parser.add_argument('--encoder-num-layer', default=4, type=int)
    parser.add_argument('--encoder-d-model', default=128, type=int)
    parser.add_argument('--encoder-d-ff', default=512, type=int)
    parser.add_argument('--encoder-num-heads', default=8, type=int)
    parser.add_argument('--encoder-dropout', default=0.1, type=float)
    parser.add_argument('--decoder-num-layer', default=4, type=int)
    parser.add_argument('--decoder-d-model', default=128, type=int)
    parser.add_argument('--decoder-d-ff', default=512, type=int)
    parser.add_argument('--decoder-num-heads', default=8, type=int)
    parser.add_argument('--decoder-dropout', default=0.1, type=float)
    # star-transformer
    parser.add_argument('--cycle-num', default=8, type=int, help='Number of inner cycles')
    parser.add_argument('--cycle-layers', default=8, type=int, help='Number of outer cycles')
    # other parameter settings
    parser.add_argument('--train-snr', default=3, type=int, help='The train SNR')
    parser.add_argument('--test-snr', default=6, type=int, help='The test SNR')
    # set by the reference's (missing) driver, not by its parser
    parser.add_argument('--vocab-size', default=22234, type=int)
    parser.add_argument('--start-idx', default=1, type=int)
    parser.add_argument('--end-idx', default=2, type=int)
    parser.add_argument('--pad-idx', default=0, type=int)
    args = parser.parse_known_args(argv)[0]   # argv=None parses sys.argv like the reference (:63)
    return args
