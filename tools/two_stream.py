"""Experiment: does splitting a bench step into independent sub-batches on separate CUDA streams let one sub-batch's small
per-step kernels (19-57 CTAs on 148 SMs) run under another's star kernel?   python tools/two_stream.py [parts ...]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepsc_gan_b200  # noqa
from deepsc_gan_b200 import _lib, sweep
from deepsc_gan_b200.dataset.synthetic import synthetic_units
from deepsc_gan_b200.models import Transeiver_Star, modules
from deepsc_gan_b200.utlis.parameters import para_config

dev = torch.device("cuda:0")
modules.set_precision(1)
torch.manual_seed(2024)
net = Transeiver_Star(para_config([])).to(dev).eval()
U = 37
steps = 8
for parts in [int(a) for a in sys.argv[1:]] or [1, 2, 3]:
    sizes = [U // parts + (1 if i < U % parts else 0) for i in range(parts)]
    runners = [sweep.SweepRunner(net, u, channel="AWGN", seed=7, graph=True) for u in sizes]
    streams = [torch.cuda.Stream() for _ in sizes]
    inputs = [[synthetic_units(s * U + sum(sizes[:i]), u).to(dev) for i, u in enumerate(sizes)] for s in range(steps + 2)]
    n_std = [torch.full((u,), 0.3, device=dev) for u in sizes]

    def step(s):
        cur = torch.cuda.current_stream()
        for i, r in enumerate(runners):
            streams[i].wait_stream(cur)
            with torch.cuda.stream(streams[i]):
                r.run(inputs[s][i], n_std[i])
        for st in streams:
            cur.wait_stream(st)

    for s in range(2):
        step(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(2, steps + 2):
        step(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{parts} sub-batch(es) {sizes}: {ms:.2f} ms per step, {U * 64 / ms * 1e3:.0f} sentences/s", flush=True)
    del runners
