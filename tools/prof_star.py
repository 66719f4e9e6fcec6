"""Profiling driver: a few star cycles on one super-batch (used under ncu; not a benchmark)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepsc_gan_b200  # noqa
from deepsc_gan_b200 import _lib as L
import deepsc_gan_b200.models.modules as M

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 1
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2368
cycles = int(sys.argv[3]) if len(sys.argv) > 3 else 3
M.set_precision(prec)
dev = torch.device("cuda:0")
torch.manual_seed(0)
sat, relay = M.sublayer1(128, 8).to(dev), M.sublayer1(128, 8).to(dev)
e = torch.randn(S, 31, 128, device=dev)
kv2 = torch.randn(S, 30, 256, device=dev)
tile = L.star_pack(e)
ws = M.StarWorkspace(S, dev)
for rep in range(3):
    M.star_cycles(tile, sat, relay, cycles, kv2, 17, ws)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
M.star_cycles(tile, sat, relay, 8, kv2, 17, ws)
t1.record()
torch.cuda.synchronize()
print("8 cycles on", S, "sentences:", t0.elapsed_time(t1), "ms")
