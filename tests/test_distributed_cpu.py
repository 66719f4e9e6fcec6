"""Host-side multi-rank logic on CPU with the gloo backend (world_size 2): work-item sharding, the count-table gather
of the SNR sweep, and the flat gradient-bucket all-reduce.  No kernel is launched here; the per-item "runner" is a
CPU stand-in that scores with the BLEU oracle (test infrastructure), so the test pins the plumbing: results must be
identical to a single-rank run."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _cases  # noqa: F401  (registers the package)
from deepsc_gan_b200 import optim, sweep
from deepsc_gan_b200.dataset.synthetic import synthetic_units
from oracle import bleu_oracle as B


class FakeRunner:
    """Corrupts a sentence-dependent number of tokens (more at low SNR) and counts n-grams on the CPU."""

    def __init__(self, U):
        self.U, self.dev = U, torch.device("cpu")

    def run(self, inp, n_std, h=None):
        ids = inp.clone()
        for u in range(self.U):
            k = int(round(float(n_std[u]) * 10)) + (0 if h is None else int(abs(float(h[u, 0])) * 3))
            ids[64 * u:64 * u + 64, 1:1 + k] = 5
        return ids, torch.from_numpy(B.bleu_counts(inp.numpy(), ids.numpy()).astype(np.int32))


def test_shard_items_partitions_everything():
    for n in (0, 1, 7, 38, 2166):
        for world in (1, 2, 3, 8):
            spans = [sweep.shard_items(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    items = sweep.work_items(3, (0, 6, 12))
    assert len(items) == 9 and items[0] == (0, 0) and items[-1] == (2, 2)


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        units = synthetic_units(0, 5)                                   # 5 units x 3 SNR points = 15 items, ragged over 2 ranks
        for channel in ("AWGN", "Rayleigh"):
            rows, counts, snr_idx = sweep.evaluate_sweep(FakeRunner(2), units, (0, 6, 12), channel=channel, rank=rank,
                                                         world=world)
            if rank == 0:
                np.save(os.path.join(tmp, f"counts_{channel}.npy"), counts.numpy())
                np.save(os.path.join(tmp, f"rows_{channel}.npy"), np.array(rows))
        # gather_counts: equal-shaped tables come back in rank order
        t = torch.full((4, 10), rank, dtype=torch.int32)
        g = sweep.gather_counts(t)
        assert g.shape == (8, 10) and g[:4].eq(0).all() and g[4:].eq(1).all()
        # flat gradient bucket: summed over ranks, factor 1/world handed to the optimizer kernel
        bucket = torch.arange(12, dtype=torch.float32).reshape(2, 6) * (rank + 1)
        scale = optim.all_reduce_mean_scale(bucket)
        assert scale == 0.5 and torch.equal(bucket, torch.arange(12, dtype=torch.float32).reshape(2, 6) * 3)
    finally:
        dist.destroy_process_group()


def test_sweep_and_gradient_bucket_over_two_gloo_ranks(tmp_path):
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    units = synthetic_units(0, 5)
    for channel in ("AWGN", "Rayleigh"):
        rows, counts, snr_idx = sweep.evaluate_sweep(FakeRunner(2), units, (0, 6, 12), channel=channel)
        assert np.array_equal(counts.numpy(), np.load(tmp_path / f"counts_{channel}.npy"))
        assert np.allclose(np.array(rows), np.load(tmp_path / f"rows_{channel}.npy"))
        assert len(rows) == 3 and rows[0][0] == 0 and rows[0][1] < rows[2][1]      # BLEU rises with SNR
        assert counts.shape == (15 * 64, 10) and snr_idx.shape == (15 * 64,)


def test_flat_params_views_and_ranges():
    """FlatParams re-points parameters to one buffer (CPU tensors suffice: no kernel is involved)."""
    lin = torch.nn.Sequential(torch.nn.Linear(3, 5), torch.nn.Linear(5, 2))
    before = [p.detach().clone() for p in lin.parameters()]
    fp = optim.FlatParams(lin, n_grad_buffers=2)
    for p, b in zip(lin.parameters(), before):
        assert torch.equal(p, b) and p.data_ptr() >= fp.flat.data_ptr()
    assert all(o % 4 == 0 for o in fp.offsets) and fp.grad_bucket.shape == (2, fp.numel)
    fp.point_grads(1)
    lin(torch.ones(1, 3)).sum().backward()
    assert fp.grad_bucket[1].abs().sum() > 0 and fp.grad_bucket[0].abs().sum() == 0
    r = fp.ranges(lambda n: n.startswith("1."))
    assert len(r) == 1 and r[0][1] == fp.numel
    fp.flat.zero_()
    assert all(float(p.abs().sum()) == 0 for p in lin.parameters())


def test_dataloader_mirror(tmp_path):
    """dataset/dataloader.py: post-padding to 31, pre-truncation, (x, x) pairs, batch 64 with a short last batch,
    a fresh permutation per epoch, raw_data[:-1] (the reference's length=-1 slice)."""
    import pickle
    from types import SimpleNamespace
    from deepsc_gan_b200.dataset import dataloader as D
    rng = np.random.RandomState(0)
    raw = [[1] + rng.randint(5, 22234, size=n).tolist() + [4, 2] for n in rng.randint(4, 29, size=150)]
    raw.append(list(range(1, 41)))                                   # longer than 31: truncated from the front
    raw.append([1, 9, 2])                                            # dropped by raw_data[:-1]
    path = tmp_path / "data.pkl"
    pickle.dump(raw, open(path, "wb"))
    pad = D.pad_sequences(raw[:-1])
    assert pad.shape == (151, 31) and pad.dtype == np.int32
    assert pad[0, : len(raw[0])].tolist() == raw[0] and not pad[0, len(raw[0]):].any()
    assert pad[150].tolist() == list(range(10, 41))
    ds = D.return_dataset(SimpleNamespace(bs=64), str(path), -1, seed=3)
    assert len(ds) == 3
    e1 = [b for b in ds]
    e2 = [b for b in ds]
    assert [tuple(x.shape) for x, _ in e1] == [(64, 31), (64, 31), (23, 31)]
    assert all(torch.equal(x, y) for x, y in e1)
    seen = torch.cat([x for x, _ in e1])
    assert sorted(map(tuple, seen.tolist())) == sorted(map(tuple, pad.tolist()))
    assert not torch.equal(seen, torch.cat([x for x, _ in e2]))      # reshuffled each epoch
    assert tuple(ds.as_units().shape) == (128, 31)
