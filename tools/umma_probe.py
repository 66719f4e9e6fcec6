"""tcgen05.mma issue-rate probe: cycles per UMMA (M=128, K=16) for TS vs SS A operands and N = 32..256."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepsc_gan_b200  # noqa
from deepsc_gan_b200 import _lib as L
L.use_debug_library()      # python deepsc-gan_b200/build.py --debug
lib = L.load()
out = torch.zeros(1, dtype=torch.int64, device="cuda:0")
iters = 2000
for ts in (1, 0, 3, 2):
    for n in (32, 64, 96, 128, 192, 256):
        for _ in range(2):
            rc = lib.dsc_umma_probe(ts, n, iters, out.data_ptr(), None)
            assert rc == 0, lib.dsc_last_error()
            torch.cuda.synchronize()
        cyc = int(out[0]) / (iters * 8)
        print(f"{'TS' if ts & 1 else 'SS'} {'elect.sync, converged warp' if ts & 2 else 'if (tid == 0)           '} N={n:3d}: {cyc:6.1f} cycles per UMMA   (math floor 128*N/256 = {128*n/256:.0f})")
