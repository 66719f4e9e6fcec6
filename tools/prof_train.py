"""Profiling driver: a few gan_train_step calls on one 64-sentence unit (used under ncu; not a benchmark)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import deepsc_gan_b200  # noqa: F401
from deepsc_gan_b200 import models, sweep
from deepsc_gan_b200.dataset.synthetic import synthetic_units
from deepsc_gan_b200.models import modules
from deepsc_gan_b200.utlis import gan_train as GT
from deepsc_gan_b200.utlis.parameters import para_config

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
modules.set_precision(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
dev = torch.device("cuda:0")
cfg = para_config([])
torch.manual_seed(2024)
net = models.Transeiver_GAN(cfg).to(dev).train()
opt = GT.make_optimizer(net, learning_rate=cfg.lr)
one = synthetic_units(0, 1).to(dev)
ns3 = float(sweep.snr_to_noise(3.0))
for i in range(steps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = GT.gan_train_step(one, one, None, net, opt, 0.5, channel="AWGN", n_std=ns3, training=True, traingan=True)
    b.record()
    torch.cuda.synchronize()
    print(f"step {i}: {a.elapsed_time(b):.2f} ms  loss {float(out[0]):.4f} g_loss {float(out[1]):.4f} d_loss {float(out[2]):.4f}")
