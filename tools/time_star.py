"""Per-kernel CUDA-event timing of the fused star-cycle kernels on one super-batch (not a benchmark)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepsc_gan_b200  # noqa
from deepsc_gan_b200 import _lib as L
import deepsc_gan_b200.models.modules as M

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 1
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2368
n2 = int(sys.argv[3]) if len(sys.argv) > 3 else 17
M.set_precision(prec)
dev = torch.device("cuda:0")
torch.manual_seed(0)
sat, relay = M.sublayer1(128, 8).to(dev), M.sublayer1(128, 8).to(dev)
xi = torch.randn(S * 4096, device=dev); atti = torch.empty(S * 4096, device=dev)
kvei = torch.randn(S * 8192, device=dev); kv2i = torch.randn(S * 8192, device=dev)
s_buf = torch.randn(S, 128, device=dev); q_r = torch.randn(S, 128, device=dev); att_r = torch.empty(S, 128, device=dev)
w_g, wo, bo = sat._packed("qkv_grouped"), sat.dense.kernel.detach(), sat.dense.bias.detach()
wkv_r, wq_r = relay._packed("kv"), relay.wq.kernel.detach()
wo_r, bo_r = relay.dense.kernel.detach(), relay.dense.bias.detach()
flush = torch.empty(64 * 1024 * 1024, device=dev)   # 256 MB > L2


def timeit(name, fn, reps=10, cold=False):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        if cold:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    print(f"{name:28s} {'cold' if cold else 'warm'} median {ts[len(ts)//2]:8.1f} us  min {ts[0]:8.1f} us")


xrow = torch.empty(S * 4096, device=dev)
for cyc in (1, 8):
    timeit(f"star_cycles_tc x{cyc} n2={n2}", lambda: L.star_cycles_tc(xi, s_buf, q_r, kvei, kv2i, n2, w_g, wo, wkv_r, wo_r, wq_r,
                                                                      bo, bo_r, xrow, S, cyc, prec), cold=True)
if "--dbg" in sys.argv:
    for dbg in (0, 1, 2, 4, 1 | 2, 1 | 4, 2 | 4, 7):
        timeit(f"star_sat_tc dbg={dbg}", lambda: L.star_sat_tc(xi, s_buf, kvei, w_g, atti, S, prec | (dbg << 8)))
for cold in (False, True):
    timeit("star_sat_tc", lambda: L.star_sat_tc(xi, s_buf, kvei, w_g, atti, S, prec), cold=cold)
    timeit("star_mix_tc", lambda: L.star_mix_tc(atti, xi, None, s_buf, wo, bo, wkv_r, q_r, kv2i, n2, att_r, S, prec), cold=cold)
    timeit("star_relay_update", lambda: L.star_relay_update(att_r, wo_r, bo_r, wq_r, s_buf, q_r), cold=cold)
