"""Data-parallel gan_train_step under torchrun (one rank per GPU, NCCL): every rank trains on its own 64-sentence unit,
the flat gradient bucket is all-reduced (one collective per step), and all replicas must stay bit-identical.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_dp.py [--steps K] [--graph]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import deepsc_gan_b200  # noqa: F401
from deepsc_gan_b200 import models, sweep
from deepsc_gan_b200.dataset.synthetic import synthetic_units
from deepsc_gan_b200.models import modules
from deepsc_gan_b200.utlis import gan_train as GT
from deepsc_gan_b200.utlis.parameters import para_config


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--graph", action="store_true")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    modules.set_precision(1)
    cfg = para_config([])
    torch.manual_seed(2024)                                       # same initial weights on every rank
    net = models.Transeiver_GAN(cfg).to(dev).train()
    opt = GT.make_optimizer(net, learning_rate=cfg.lr)
    torch.manual_seed(100 + rank)                                 # different noise / dropout per rank
    modules.set_dropout_seed(7000 + rank)
    units = [synthetic_units(rank * 1000 + s, 1).to(dev) for s in range(4)]
    ns3 = float(sweep.snr_to_noise(3.0))
    step = GT.GraphedGanTrainStep(net, opt, 0.5, n_std=ns3, traingan=True) if a.graph else None
    run = (lambda u: step(u, u)) if a.graph else (lambda u: GT.gan_train_step(u, u, None, net, opt, 0.5, channel="AWGN", n_std=ns3,
                                                                                training=True, traingan=True))
    for s in range(3):
        run(units[s % 4])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    first = last = None
    for s in range(a.steps):
        out = run(units[s % 4])
        if s == 0:
            first = float(out[0])
    e1.record()
    torch.cuda.synchronize()
    last = float(out[0])
    ms = e0.elapsed_time(e1) / a.steps
    # replicas identical?
    digest = opt.fp.flat.double().sum().reshape(1)
    same = True
    if world > 1:
        parts = [torch.empty_like(digest) for _ in range(world)]
        dist.all_gather(parts, digest)
        same = all(bool(torch.equal(p, parts[0])) for p in parts)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    if rank == 0:
        print(json.dumps({"workload": "gan_train_step data-parallel, one 64-sentence unit per rank per step",
                          "graph": a.graph, "n_gpus": world, "ms_per_step": ms, "value": 64 * world / (ms * 1e-3),
                          "unit": "training sentences/s", "replicas_identical": same, "loss_first": first, "loss_last": last,
                          "optimizer_iterations": opt.iterations}))
    # GraphedGanTrainStep keeps NCCL out of its graphs (two graphs, all-reduce on the stream between them), so the
    # communicator is torn down the ordinary way
    sys.stdout.flush()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
