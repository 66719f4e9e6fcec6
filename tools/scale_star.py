"""Per-launch fixed cost vs per-tile cost of the fused star kernels: time vs tiles per CTA."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepsc_gan_b200  # noqa
from deepsc_gan_b200 import _lib as L
import deepsc_gan_b200.models.modules as M

dev = torch.device("cuda:0")
torch.manual_seed(0)
sat, relay = M.sublayer1(128, 8).to(dev), M.sublayer1(128, 8).to(dev)
w_g, wo, bo = sat._packed("qkv_grouped"), sat.dense.kernel.detach(), sat.dense.bias.detach()
wkv_r, wq_r = relay._packed("kv"), relay.wq.kernel.detach()
wo_r, bo_r = relay.dense.kernel.detach(), relay.dense.bias.detach()


def t(fn, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for prec in (1, 2):
    for tiles_per_cta in (1, 2, 4, 8, 16):
        S = 592 * tiles_per_cta
        xi = torch.randn(S * 4096, device=dev); atti = torch.empty(S * 4096, device=dev)
        kvei = torch.randn(S * 8192, device=dev); kv2i = torch.randn(S * 8192, device=dev)
        s_buf = torch.randn(S, 128, device=dev); q_r = torch.randn(S, 128, device=dev); att_r = torch.empty(S, 128, device=dev)
        a = t(lambda: L.star_sat_tc(xi, s_buf, kvei, w_g, atti, S, prec))
        b = t(lambda: L.star_mix_tc(atti, xi, None, s_buf, wo, bo, wkv_r, q_r, kv2i, 17, att_r, S, prec))
        c = t(lambda: L.star_relay_update(att_r, wo_r, bo_r, wq_r, s_buf, q_r))
        print(f"prec {prec} tiles/CTA {tiles_per_cta:2d} S {S:5d}: sat {a:7.1f} us  mix {b:7.1f} us  relay_update {c:6.1f} us", flush=True)
        del xi, atti, kvei, kv2i
