"""Timeline of the two-tile star kernel (CTA 0, compute warps 0 and 8): per chunk, how long the warp waited for its
accumulator and how long the chunk's register work took.  Debug tool (needs `python deepsc-gan_b200/build.py --debug`).

    python tools/pp_trace.py [cycles] [flags: 0 | 0x300 ...] [n2]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import deepsc_gan_b200  # noqa: F401
from deepsc_gan_b200 import _lib as L
import deepsc_gan_b200.models.modules as M

L.use_debug_library()
cycles = int(sys.argv[1]) if len(sys.argv) > 1 else 8
flags = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0x300
n2 = int(sys.argv[3]) if len(sys.argv) > 3 else 17
S = 2368
M.set_precision(1)
dev = torch.device("cuda:0")
torch.manual_seed(0)
sat, relay = M.sublayer1(128, 8).to(dev), M.sublayer1(128, 8).to(dev)
xi = torch.randn(S * 4096, device=dev); kvei = torch.randn(S * 8192, device=dev); kv2i = torch.randn(S * 8192, device=dev)
s_buf = torch.randn(S, 128, device=dev); q_r = torch.randn(S, 128, device=dev)
xrow = torch.empty(S * 4096, device=dev)
L.STAR_FORM = L.STAR_FORM_TWO_TILE
run = lambda: L.star_cycles_tc(xi, s_buf, q_r, kvei, kv2i, n2, sat._packed("qkv_grouped"), sat.dense.kernel.detach(),
                               relay._packed("kv"), relay.dense.kernel.detach(), relay.wq.kernel.detach(),
                               sat.dense.bias.detach(), relay.dense.bias.detach(), xrow, S, cycles, 1 | flags)
for _ in range(3):
    run()
buf = torch.zeros(768, dtype=torch.int64, device=dev)
assert L.load().dsc_debug_star_trace(buf.data_ptr()) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record()
torch.cuda.synchronize()
L.load().dsc_debug_star_trace(None)
print(f"launch {e0.elapsed_time(e1) * 1e3:.1f} us (traced), cycles={cycles} flags={flags:#x} n2={n2}")
t = buf.cpu().tolist()

# the chunk sequence of CTA 0 (dsc_star_pp.cu Plan): 4 tiles -> 2 per slot
skip0, nfr = int(bool(flags & 0x100)), int(bool(flags & 0x200))
Lr = 2 * cycles - skip0 - nfr
Lp = (Lr + 1) & ~1
n_tiles = S // 4
my_tiles = (n_tiles - 1) // 148 + 1
nA, nB = (my_tiles + 1) // 2, my_tiles // 2


def decode(p, nt):
    if p < 0:
        return None
    ti, q = divmod(p, Lp)
    if ti >= nt or q >= Lr:
        return None
    idx = q + skip0
    return (ti, idx >> 1, idx & 1)


H = max(nA * Lp, nB * Lp + 1 if nB else 0)
seq = []
for h in range(H):
    phs = [decode(h, nA), decode(h - 1, nB)]
    for jc in range(3):
        for T in range(2):
            if phs[T]:
                ti, c, kind = phs[T]
                seq.append((h, "AB"[T], ti, c, ("S1a", "S1b", "S2")[jc] if kind == 0 else ("R1", "R2", "R3")[jc]))
ghz = 1.965
for wname, base in (("warp 0 (head pairs 0, 2)", 0), ("warp 8 (head pairs 1, 3)", 384)):
    st = t[base:base + 384]
    t0 = st[0]
    print(wname)
    agg = {}
    n = min(len(seq), 128)
    for i in range(n):
        a, b, c_ = st[3 * i], st[3 * i + 1], st[3 * i + 2]
        if not (a and b and c_):
            break
        h, slot, ti, c, name = seq[i]
        wait, work = (b - a) / ghz / 1e3, (c_ - b) / ghz / 1e3
        w, k, cnt = agg.get(name, (0.0, 0.0, 0))
        if ti == 0 and 1 <= c < cycles - 1 or True:
            agg[name] = (w + wait, k + work, cnt + 1)
        if i < 48:
            print(f"  h{h:2d} {slot}.{name:3s} tile{ti} c{c}: start {(a - t0) / ghz / 1e3:8.2f} us  wait {wait:5.2f}  work {work:5.2f}")
    print("  mean per chunk:   " + "   ".join(f"{k}: wait {w / c:.2f} work {x / c:.2f}" for k, (w, x, c) in agg.items()))
    tw = sum(w for w, _, _ in agg.values()); tx = sum(x for _, x, _ in agg.values())
    print(f"  total over {sum(c for _, _, c in agg.values())} chunks: wait {tw:.1f} us, work {tx:.1f} us, span {(st[3 * i - 1] - t0) / ghz / 1e3:.1f} us")
