"""Timeline of one tile of the fused star kernel (CTA 0): where a cycle's ~12 us go.  Debug tool, not a benchmark.
Needs the debug-tools library (`python deepsc-gan_b200/build.py --debug`: the product sources with -DDSC_DEBUG_TOOLS=1).

    python tools/star_trace.py [prec] [n2] [cycles]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import deepsc_gan_b200  # noqa: F401
from deepsc_gan_b200 import _lib as L
import deepsc_gan_b200.models.modules as M

L.use_debug_library()          # the trace hooks are compiled out of libdeepsc_b200.so

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n2 = int(sys.argv[2]) if len(sys.argv) > 2 else 17
cycles = int(sys.argv[3]) if len(sys.argv) > 3 else 4
S = 592
M.set_precision(prec)
dev = torch.device("cuda:0")
torch.manual_seed(0)
sat, relay = M.sublayer1(128, 8).to(dev), M.sublayer1(128, 8).to(dev)
xi = torch.randn(S * 4096, device=dev); kvei = torch.randn(S * 8192, device=dev); kv2i = torch.randn(S * 8192, device=dev)
s_buf = torch.randn(S, 128, device=dev); q_r = torch.randn(S, 128, device=dev)
w_g, wo, bo = sat._packed("qkv_grouped"), sat.dense.kernel.detach(), sat.dense.bias.detach()
wkv_r, wq_r = relay._packed("kv"), relay.wq.kernel.detach()
wo_r, bo_r = relay.dense.kernel.detach(), relay.dense.bias.detach()
xrow = torch.empty(S * 4096, device=dev)
run = lambda: L.star_cycles_tc(xi, s_buf, q_r, kvei, kv2i, n2, w_g, wo, wkv_r, wo_r, wq_r, bo, bo_r, xrow, S, cycles, prec)
for _ in range(3):
    run()
buf = torch.zeros(768, dtype=torch.int64, device=dev)
assert L.load().dsc_debug_star_trace(buf.data_ptr()) == 0
run()
torch.cuda.synchronize()
L.load().dsc_debug_star_trace(None)
t = buf.cpu().tolist()
iss, w0, w8 = t[0:256], t[256:512], t[512:768]
t0 = min(x for x in t if x > 0)
ghz = 1.965
us = lambda c: (c - t0) / ghz / 1e3 if c else float("nan")
jobs_per_cycle = 9
print(f"prec={prec} n2={n2} cycles={cycles}; times in us from the first stamp (assuming {ghz} GHz)")
print("issuer: job, ready-to-issue, issued")
n = 0
for c in range(cycles):
    for j in ([0, 1, 2, 3] + ([8] if c else []) + [4, 5, 6, 7]):      # issue order: J8 (this cycle's relay query) behind J3
        print(f"  c{c} J{j}: ready {us(iss[2*n]):8.2f}  issued {us(iss[2*n+1]):8.2f}")
        n += 1
names0 = ["x_ready(tile)"]
per0 = ["J0 acc seen", "J0 acc freed", "ATT heads 0..3 staged (ta_ready)", "J2 acc seen", "J2 acc freed", "ATT staged (t_ready)", "J4 acc seen", "J4 acc freed",
        "J4 bias+relu done", "J4 hi/lo split done", "J4 tcgen05.st landed", "X' staged (x_ready)", "J5 acc seen", "J5 softmax weights", "J5 freed", "J6 acc seen", "J6 freed", "att_r staged (t_ready)", "J7 acc seen",
        "J7 freed", "s' staged (t_ready)", "s' barrier passed", "relay row patched", "X patched (x_ready)"]
print("compute warp 0:")
i = 0
print(f"  {names0[0]:28s} {us(w0[i]):8.2f}"); i += 1
for c in range(cycles):
    last = c + 1 == cycles
    ev = per0[:20] + per0[21:22] if last else per0     # last cycle: no s' staging for a next query, no patch
    for name in ev:
        print(f"  c{c} {name:25s} {us(w0[i]):8.2f}")
        i += 1
per8 = ["J1 acc seen", "J1 acc freed", "ATT heads 0..3 staged (ta_ready)", "J3 acc seen", "J3 acc freed", "ATT staged (t_ready)"]
print("compute warp 8 (first events of each cycle are its QKV jobs):")
i = 1
for c in range(min(cycles, 2)):
    for name in per8:
        print(f"  c{c} {name:25s} {us(w8[i]):8.2f}")
        i += 1
    i += (len(per0) if c + 1 < cycles else 21) - len(per8)
