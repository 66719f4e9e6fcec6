"""CUDA-event timing of the vocabulary-sized backward GEMMs of a training step (tensor-core path vs the fp32 FFMA kernel)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import deepsc_gan_b200  # noqa: F401
from deepsc_gan_b200 import autograd as AG

dev = torch.device("cuda:0")
torch.manual_seed(0)
V, R = 22234, 1920
dz = torch.randn(R, V, device=dev)
w = torch.randn(128, V, device=dev)
x = torch.randn(R, 128, device=dev)


def timed(name, fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    print(f"{name:60s} median {ts[len(ts)//2]:8.1f} us  ({2*R*128*V*1e-6/ts[len(ts)//2]:.1f} TFLOP/s fp32-class)")


timed("dX = dZ[1920,V] @ W[128,V]^T (dsc_gemm -> tensor cores)", lambda: AG.gemm(dz, w, trans_b=True))
timed("dW = X[1920,128]^T @ dZ[1920,V] (transposes + tensor cores)", lambda: AG.gemm(x, dz, trans_a=True))
AG.TC_GEMM_MIN_MACS = 10 ** 18
timed("dW, fp32 FFMA kernel (dsc_gemm trans_a)", lambda: AG.gemm(x, dz, trans_a=True))
at, bt = torch.empty(128, R, device=dev), torch.empty(V, R, device=dev)
from deepsc_gan_b200._lib import load, _stream, _check
timed("dsc_transpose dZ -> [V,1920]", lambda: _check(load().dsc_transpose(dz.data_ptr(), V, bt.data_ptr(), R, R, V, _stream()), "t"))
out = torch.empty(128, V, device=dev)
timed("dsc_gemm_nt_tc alone [128,1920] x [V,1920]^T", lambda: _check(load().dsc_gemm_nt_tc(at.data_ptr(), R, bt.data_ptr(), R, out.data_ptr(), V, 128, V, R, 0, _stream()), "g"))
