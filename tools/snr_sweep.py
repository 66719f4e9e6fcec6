"""The reference's missing top level: an SNR sweep of one system over a test set, written as a
result pickle in the layout of the reference's ``log/eval-D-GAN-STAR/*.pkl`` - a list of rows
``[snr_idx, BLEU-1, BLEU(1/4,1/4,1/4,1/4)]`` - next to the per-sentence int32 count table.

    python tools/snr_sweep.py --system Transeiver_Star --channel AWGN --units 38 --out log/eval-star.pkl
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/snr_sweep.py --data data/txt/test_data.pkl --channel Rayleigh

Weights are random-initialised (the reference ships checkpoint indexes without data) unless ``--weights``
names a ``torch.save``d state dict with the TF variable names.  (SNR point, 64-sentence unit) items are dealt to the
ranks; the only exchange is the final gather of the count table.
"""
import argparse
import json
import os
import pickle
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import deepsc_gan_b200  # noqa: F401
from deepsc_gan_b200 import models, sweep
from deepsc_gan_b200.dataset import dataloader
from deepsc_gan_b200.dataset.synthetic import synthetic_units
from deepsc_gan_b200.models import modules
from deepsc_gan_b200.utlis.parameters import para_config


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--system", default="Transeiver_Star", choices=["Transeiver", "Transeiver_Star", "Transeiver_star", "Transeiver_GAN"])
    ap.add_argument("--channel", default="AWGN", choices=["AWGN", "Rayleigh", "Rician"])
    ap.add_argument("--detector", type=int, default=0, help="0 reference parity (unequalised), 1 LS, 2 MMSE")
    ap.add_argument("--attack", default=None, choices=[None, "generator"])
    ap.add_argument("--psr-db", type=float, default=-6.0)
    ap.add_argument("--data", default=None, help="test_data.pkl of the reference (default: synthetic Europarl-shape units)")
    ap.add_argument("--units", type=int, default=38, help="synthetic 64-sentence units when --data is not given")
    ap.add_argument("--units-per-launch", type=int, default=37)
    ap.add_argument("--weights", default=None)
    ap.add_argument("--prec", type=int, default=1)
    ap.add_argument("--out", default="log/eval.pkl")
    a = ap.parse_args()

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    modules.set_precision(a.prec)
    cfg = para_config([])
    torch.manual_seed(2024)
    net = getattr(models, a.system)(cfg).to(dev).eval()
    if a.weights:
        net.load_tf_state_dict(torch.load(a.weights, map_location=dev))
    if a.data and a.data.endswith(".npz"):
        # tests/golden/europarl_test.npz: the reference's test_data.pkl already padded (make_europarl_fixture.py); the
        # [:-1] slice and the full-unit cut are dataset/dataloader.py:11-14's
        import numpy as np
        ids = torch.from_numpy(np.load(a.data)["ids"].astype(np.int32))[:-1]
        units = ids[: ids.shape[0] // 64 * 64]
    else:
        units = dataloader.return_dataset(cfg, a.data, -1, shuffle=False).as_units() if a.data else synthetic_units(0, a.units)
    K = 1 if a.channel == "Rician" else 0
    runner = sweep.SweepRunner(net, a.units_per_launch, channel="AWGN" if a.channel == "AWGN" else "Rayleigh", detector=a.detector,
                               seed=1, attack=a.attack, psr_db=a.psr_db)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    rows, counts, snr_index = sweep.evaluate_sweep(runner, units, channel=a.channel, K=K, rank=rank, world=world)
    t1.record()
    torch.cuda.synchronize()
    if rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "wb") as f:
            pickle.dump(rows, f)
        with open(os.path.splitext(a.out)[0] + "-counts.pkl", "wb") as f:
            pickle.dump({"counts": counts.numpy(), "snr_index": snr_index}, f)
        with open(os.path.splitext(a.out)[0] + "-rows.json", "w") as f:
            json.dump([[float(v) for v in r] for r in rows], f)
        n_eval = len(snr_index)
        print(json.dumps({"system": a.system, "channel": a.channel, "sentences": int(units.shape[0]), "snr_points": len(rows),
                          "sentence_evaluations": n_eval, "n_gpus": world, "seconds": t0.elapsed_time(t1) * 1e-3,
                          "sentences_per_s": n_eval / (t0.elapsed_time(t1) * 1e-3), "out": a.out,
                          "bleu1_at_0_9_18_dB": [rows[i][1] for i in (0, 9, 18)]}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
