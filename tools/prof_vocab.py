"""Profiling driver: the fused vocabulary projection + argmax on one super-batch (used under ncu; not a benchmark)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import deepsc_gan_b200  # noqa: F401
from deepsc_gan_b200 import _lib as L

M = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
V = 22234
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(M, 128, device=dev)
w = torch.zeros(128, (V + 127) // 128 * 128, device=dev)
w[:, :V] = torch.randn(128, V, device=dev) * 0.02
b = torch.randn(V, device=dev) * 0.01
ids = torch.zeros(M, dtype=torch.int32, device=dev)
for _ in range(3):
    L.vocab_argmax(x, w, b, V, ids, prec=1)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); L.vocab_argmax(x, w, b, V, ids, prec=1); e.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(e) * 1e3)
ts.sort()
print(f"vocab_argmax_tc M={M}: median {ts[5]:.1f} us, min {ts[0]:.1f} us; {2*M*128*22272*3/ts[5]/1e6:.0f} TFLOP/s bf16 executed")
