"""Two-tile star kernel against the one-tile kernel: results must be bit-identical for every flag combination, tile count
and cycle count; then A/B timing at the bench size.   python tools/pp_check.py [time-only]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepsc_gan_b200  # noqa
from deepsc_gan_b200 import _lib as L
import deepsc_gan_b200.models.modules as M

L.use_debug_library()          # the two-tile kernel is compiled into libdeepsc_b200_debug.so only

dev = torch.device("cuda:0")
M.set_precision(1)
torch.manual_seed(0)
sat, relay = M.sublayer1(128, 8).to(dev), M.sublayer1(128, 8).to(dev)


def run(S, cycles, n2, flags, form, ws, kv2i, out):
    L.STAR_FORM = form
    try:
        skip = bool(flags & L.STAR_FIRST_SAT_DONE)
        return L.star_cycles_tc(ws.xi1 if skip else ws.xi0, ws.s0, ws.q0, ws.kvei, kv2i, n2, sat._packed("qkv_grouped"),
                                sat.dense.kernel.detach(), relay._packed("kv"), relay.dense.kernel.detach(),
                                relay.wq.kernel.detach(), sat.dense.bias.detach(), relay.dense.bias.detach(), out, S, cycles,
                                1 | flags)
    finally:
        L.STAR_FORM = 0


def setup(S, prec=1):
    M.set_precision(prec)
    e = torch.randn(S, 31, 128, device=dev)
    tile = L.star_pack(e)
    ws = M.StarWorkspace(S, dev)
    L.STAR_FORM = L.STAR_FORM_ONE_TILE
    M.prepare_kv_e(tile, sat, ws, relay, first_sat=True)
    L.STAR_FORM = 0
    pad = torch.zeros((S, 32, 256), device=dev)
    pad[:, :30] = torch.randn(S, 30, 256, device=dev)
    kv2i = L.star_interleave(pad, torch.empty_like(pad).view(-1), 32)
    return ws, kv2i


if "time-only" not in sys.argv:
    bad = 0
    for S in (8, 12, 16, 40, 596, 1188, 2368, 2372):
        ws, kv2i = setup(S)
        for cycles in (1, 2, 3, 8):
            for flags in (0, L.STAR_NO_FINAL_RELAY, L.STAR_FIRST_SAT_DONE, L.STAR_FIRST_SAT_DONE | L.STAR_NO_FINAL_RELAY):
                if (flags & L.STAR_FIRST_SAT_DONE) and cycles < 2:
                    continue
                for n2 in (0, 17):
                    a = run(S, cycles, n2, flags, L.STAR_FORM_ONE_TILE, ws, kv2i if n2 else None, torch.zeros(S, 32, 128, device=dev))
                    b = run(S, cycles, n2, flags, L.STAR_FORM_TWO_TILE, ws, kv2i if n2 else None, torch.zeros(S, 32, 128, device=dev))
                    torch.cuda.synchronize()
                    rows = slice(0, 31) if (flags & L.STAR_NO_FINAL_RELAY) else slice(0, 32)
                    same = torch.equal(a[:, rows], b[:, rows])
                    if not same:
                        bad += 1
                        d = (a[:, rows] - b[:, rows]).abs()
                        print(f"MISMATCH S={S} cycles={cycles} flags={flags:#x} n2={n2}: max|d|={float(d.max()):.3e} "
                              f"bad sentences={int((d.amax((1, 2)) > 0).sum())} nan={bool(torch.isnan(b).any())}", flush=True)
        print(f"S={S}: done", flush=True)
    print("bit-identical" if bad == 0 else f"{bad} mismatching cases")
    # bf16 single-pass form too
    ws, kv2i = setup(600, prec=2)
    for flags in (0, L.STAR_FIRST_SAT_DONE | L.STAR_NO_FINAL_RELAY):
        outs = []
        for form in (L.STAR_FORM_ONE_TILE, L.STAR_FORM_TWO_TILE):
            L.STAR_FORM = form
            skip = bool(flags & L.STAR_FIRST_SAT_DONE)
            outs.append(L.star_cycles_tc(ws.xi1 if skip else ws.xi0, ws.s0, ws.q0, ws.kvei, kv2i, 9, sat._packed("qkv_grouped"),
                                         sat.dense.kernel.detach(), relay._packed("kv"), relay.dense.kernel.detach(),
                                         relay.wq.kernel.detach(), sat.dense.bias.detach(), relay.dense.bias.detach(),
                                         torch.zeros(600, 32, 128, device=dev), 600, 4, 2 | flags).clone())
        L.STAR_FORM = 0
        rows = slice(0, 31) if flags else slice(0, 32)
        print("prec 2 flags", hex(flags), "identical:", torch.equal(outs[0][:, rows], outs[1][:, rows]))
    M.set_precision(1)

S = 2368
ws, kv2i = setup(S)
out = torch.zeros(S, 32, 128, device=dev)
for name, cycles, flags in (("full 8 cycles", 8, 0), ("greedy step (first-sat-done, no final relay)", 8,
                                                      L.STAR_FIRST_SAT_DONE | L.STAR_NO_FINAL_RELAY)):
    for form, fname in ((L.STAR_FORM_ONE_TILE, "one-tile"), (L.STAR_FORM_TWO_TILE, "two-tile")):
        for _ in range(3):
            run(S, cycles, 17, flags, form, ws, kv2i, out)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10):
            run(S, cycles, 17, flags, form, ws, kv2i, out)
        t1.record()
        torch.cuda.synchronize()
        print(f"{name}: {fname} {t0.elapsed_time(t1) / 10 * 1e3:.1f} us per launch on {S} sentences", flush=True)
