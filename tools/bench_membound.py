"""Achieved HBM bandwidth of the memory-bound kernels of the transmit path at sizes far beyond L2 (CUDA events, warm,
median of 10).  Algorithmic bytes = the tensors a launch must read + write once, against the measured copy
bandwidth in MEASURED_PEAKS.json.  One JSON line per kernel.

    python tools/bench_membound.py [units]      # default 8192 units of 64 sentences
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import deepsc_gan_b200  # noqa: F401
from deepsc_gan_b200 import _lib as L

U = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
S = U * 64
dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
torch.manual_seed(0)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def emit(kernel, what, nbytes, ms):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": kernel, "workload": what, "algorithmic_bytes": nbytes, "ms": round(ms, 4),
                      "achieved_gbs": round(gbs, 1), "peak_gbs": peak, "frac": round(gbs / peak, 3)}), flush=True)


E = S * 31 * 16                                   # channel symbols (floats) of S sentences
x = torch.randn(E, device=dev)
n_std = torch.full((U,), 0.3, device=dev)
noise = torch.randn(E, device=dev)
ssq = L.unit_sumsq(x, U)
emit("unit_sumsq_kernel", f"{U} units x 31,744 floats", 4 * E, timed(lambda: L.unit_sumsq(x, U)))
emit("power_normalize_kernel", "x -> x / rms(unit)", 8 * E, timed(lambda: L.power_normalize(x, U, 1.0, ssq)))
emit("channel_kernel", "AWGN, Philox noise generated in the kernel (read x, write y)", 8 * E,
     timed(lambda: L.channel(x, U, n_std, x_sumsq=ssq, seed=1)))
emit("channel_kernel", "AWGN, injected noise tensor (read x, z; write y)", 12 * E,
     timed(lambda: L.channel(x, U, n_std, x_sumsq=ssq, noise=noise)))
h = torch.randn(U, 2, device=dev).contiguous()
emit("channel_kernel", "Rayleigh + MMSE equaliser, Philox noise (read x, write y)", 8 * E,
     timed(lambda: L.channel(x, U, n_std, x_sumsq=ssq, seed=1, h=h, detector=2)))
p = torch.randn(E, device=dev)
psq = L.unit_sumsq(p, U)
emit("channel_kernel", "AWGN + generator perturbation at a power budget (read x, p; write y)", 12 * E,
     timed(lambda: L.channel(x, U, n_std, x_sumsq=ssq, seed=1, p=p, p_sumsq=psq, p_factor=0.5)))
del noise, p

R = S * 31 // 8                                    # rows of 128 floats
a = torch.randn(R // 31, 31, 128, device=dev); r = torch.randn_like(a); o = torch.empty_like(a)
g, b = torch.ones(128, device=dev), torch.zeros(128, device=dev)
emit("add_layernorm_kernel", "LayerNorm(x + res) on rows of 128", 12 * a.numel(), timed(lambda: L.add_layernorm(a, r, g, b, out=o)))
del a, r, o

V = 22234
ids = torch.randint(0, V, (S // 8, 31), device=dev, dtype=torch.int32)
table = torch.randn(V, 128, device=dev); pos = torch.randn(64, 128, device=dev)
emit("embed_kernel", "embedding * sqrt(128) + positional rows (ids in, 512 B rows out, the 11 MB table once: gathers hit L2)",
     4 * ids.numel() * 128 + 4 * ids.numel() + 4 * table.numel(),
     timed(lambda: L.embed(ids, table, pos)))

rows = 16384
logits = torch.randn(rows, V, device=dev)
emit("argmax_rows_kernel", f"{rows} rows x {V} logits", 4 * rows * V, timed(lambda: L.argmax_rows(logits)))
tgt = torch.randint(1, V, (rows,), device=dev, dtype=torch.int32)
emit("masked_ce_rows_kernel", f"{rows} rows x {V} logits", 4 * rows * V, timed(lambda: L.masked_ce_rows(logits, tgt)))
del logits

ref = torch.randint(1, V, (S, 31), device=dev, dtype=torch.int32); hyp = ref.clone()
emit("bleu_counts_kernel", "31-token pairs -> 10 int32 counts", S * (2 * 31 * 4 + 40), timed(lambda: L.bleu_counts(ref, hyp)))
