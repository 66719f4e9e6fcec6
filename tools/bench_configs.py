"""Device-timed throughput of the five BASELINE.json configs on one GPU (one JSON line each).  Not the driver's
bench (that is bench.py = configs[1]); these are the numbers quoted in DESIGN.md / BASELINE.md for the other configs.

    python tools/bench_configs.py [--units U] [--steps K] [--only 1,3,4,5]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import deepsc_gan_b200  # noqa: F401
from deepsc_gan_b200 import _lib, sweep
from deepsc_gan_b200.dataset.synthetic import synthetic_units
from deepsc_gan_b200 import models
from deepsc_gan_b200.models import modules
from deepsc_gan_b200.utlis.parameters import para_config


def timed(fn, steps, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--units", type=int, default=37)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--prec", type=int, default=1)
    ap.add_argument("--only", default="1,2,3,4,5")
    args = ap.parse_args()
    only = {int(x) for x in args.only.split(",")}
    dev = torch.device("cuda:0")
    modules.set_precision(args.prec)
    cfg = para_config([])
    U, S = args.units, args.units * 64
    inp = synthetic_units(0, U).to(dev)
    snrs = [float(u % 19) for u in range(U)]
    n_std = torch.tensor([sweep.snr_to_noise(s) for s in snrs], dtype=torch.float32, device=dev)

    def emit(config, what, ms, sentences, **extra):
        print(json.dumps({"config": config, "workload": what, "ms_per_step": ms, "sentences_per_step": sentences,
                          "value": sentences / (ms * 1e-3), "unit": "sentences/s", "prec": args.prec, **extra}), flush=True)

    if 1 in only:
        torch.manual_seed(2024)
        net = models.Transeiver(cfg).to(dev).eval()
        r = sweep.SweepRunner(net, U, channel="AWGN", seed=1)
        ns6 = torch.full((U,), sweep.snr_to_noise(6.0), device=dev)
        emit(1, "Transeiver (4+4 layers) AWGN SNR 6 dB, greedy 30 steps + BLEU counts", timed(lambda: r.run(inp, ns6), args.steps), S)
    if 2 in only or 3 in only:
        torch.manual_seed(2024)
        net = models.Transeiver_Star(cfg).to(dev).eval()
        if 2 in only:
            r = sweep.SweepRunner(net, U, channel="AWGN", seed=1)
            emit(2, "Transeiver_Star AWGN SNR 0..18 dB, greedy + BLEU counts", timed(lambda: r.run(inp, n_std), args.steps), S)
        if 3 in only:
            g = torch.Generator().manual_seed(7)
            h = sweep.fading_coefficients(0, U, g).to(dev).contiguous()
            for det, name in ((0, "reference parity (unequalised y, D6)"), (2, "MMSE equaliser applied")):
                r = sweep.SweepRunner(net, U, channel="Rayleigh", detector=det, seed=1)
                emit(3, f"Transeiver_Star Rayleigh, {name}, SNR 0..18 dB, greedy + BLEU counts",
                     timed(lambda: r.run(inp, n_std, h=h), args.steps), S)
    if 4 in only:
        torch.manual_seed(2024)
        net = models.Transeiver_GAN(cfg).to(dev).eval()
        r = sweep.SweepRunner(net, U, channel="AWGN", seed=1, attack="generator", psr_db=-6.0)
        emit(4, "Transeiver_GAN, generator perturbation at PSR -6 dB, SNR 0..18 dB, greedy + BLEU counts",
             timed(lambda: r.run(inp, n_std), args.steps), S)
        from deepsc_gan_b200.utlis import eval as E
        one = inp[:64]
        emit(4, "Transeiver_GAN eval_step_FGM (clean fwd + d loss/d symbols + attacked fwd), one 64-unit per call",
             timed(lambda: E.eval_step_FGM(one, one, net, 0.0, channel="AWGN", n_std=float(n_std[6])), max(args.steps, 5)), 64)
    if 5 in only:
        from deepsc_gan_b200.utlis import gan_train as GT
        torch.manual_seed(2024)
        net = models.Transeiver_GAN(cfg).to(dev).train()
        opt = GT.make_optimizer(net, learning_rate=cfg.lr)
        one = inp[:64]
        ns3 = float(sweep.snr_to_noise(3.0))
        emit(5, "gan_train_step eager (Transeiver_GAN, traingan=True, lambda 0.5, dropout 0.1, Adam), one 64-unit per step",
             timed(lambda: GT.gan_train_step(one, one, None, net, opt, 0.5, channel="AWGN", n_std=ns3, training=True,
                                             traingan=True), max(args.steps, 10), warmup=5), 64)
        step = GT.GraphedGanTrainStep(net, opt, 0.5, n_std=ns3, traingan=True)
        emit(5, "gan_train_step as one CUDA-graph launch per step (GraphedGanTrainStep), one 64-unit per step",
             timed(lambda: step(one, one), max(args.steps, 20), warmup=3), 64)


if __name__ == "__main__":
    main()
