#!/usr/bin/env python
"""bench.py - sentences/sec of the SNR-sweep encode -> channel -> greedy decode -> BLEU counts path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--units U]

Workload (BASELINE.json configs[1]): Star-Transformer DeepSC-GAN (`Transeiver_Star`, SE/SD, cycle_num 8)
over AWGN, SNR sweep 0..18 dB, greedy argmax decode of 30 steps, BLEU n-gram counts.  One "step" = one pass of
the path over a batch of U 64-sentence synthetic Europarl-shape units (U = 37 by default = 592 four-sentence tiles = 4 x 148 SMs; SNR points of the
sweep are dealt round-robin), weights random-init (the reference's checkpoints are missing).  One "sentence" = one sentence
evaluated at one SNR point.  N > 1: each rank processes its own U units (weak scaling, no data-path collective);
the int32 BLEU count table is all-gathered at the end of every step.

value   : inputs resident in HBM, CUDA-event time, max over ranks.
e2e     : same metric through the public API with HOST buffers: per step the ids are copied from pinned host
          memory and the BLEU counts are read back, both inside the timed region.
roofline: the dominant kernel (star_fused_kernel: 8 star cycles per launch, tensor-pipe bound), algorithmic FLOPs x 3 bf16 passes
          per launch / mean event-timed launch, against the measured bf16 peak.
cpu_baseline / --impl reference: the CPU oracle (PyTorch restatement of the reference; the TensorFlow reference
          itself cannot run here) on a bounded sample of the same workload, all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "sentences/sec for SNR-sweep encode->channel->decode+BLEU"
UNIT = "sentences/s"
SNRS = list(range(19))
# dram__bytes_read.sum + dram__bytes_write.sum of one star_fused_kernel<3> launch (8 cycles) from the `ncu --set full`
# capture summarised in profiles/ (keyed by sentences per launch); null when no capture exists for the size that ran
ROOFLINE_TRAFFIC_BYTES = {2368: 177608704 + 24980992}   # profiles/r02_ncu_star_fused.txt (full 8-cycle launch, n2 = 17)
# sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active of the same capture (the kernel at its own clock)
NCU_TENSOR_ACTIVE_PCT = {2368: 43.2}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def cpu_oracle_rate(n_units: int, first_unit: int = 0, last_only: bool = True):
    """Oracle greedy decode + BLEU counts over ``n_units`` units (one SNR point each, cycling the sweep).
    ``last_only=False`` is greedy as the reference writes it (utlis/eval.py:106-113: the decoder's vocabulary projection
    over the whole prefix at every step); ``True`` projects the newest position only (same ids, the reference's CPU path
    given the one optimisation any port would make)."""
    from deepsc_gan_b200.dataset.synthetic import synthetic_unit
    from oracle import bleu_oracle, deepsc_oracle as O
    torch.set_num_threads(os.cpu_count())
    spec = O.Spec("Transeiver_Star")
    P = O.init_params(spec, seed=2024)
    t0 = time.perf_counter()
    with torch.no_grad():
        for u in range(n_units):
            inp = synthetic_unit(first_unit + u).long()
            g = torch.Generator().manual_seed(7 + u)
            z = torch.randn(64, 31, 16, generator=g)
            ids = O.greedy_decode_noattack(P, spec, inp, 0.0, "AWGN", O.snr_to_noise(SNRS[u % 19]), z, last_only=last_only)
            bleu_oracle.bleu_counts(inp.numpy(), ids.numpy())
    dt = time.perf_counter() - t0
    return 64 * n_units / dt, dt


def run_reference(args):
    """--impl reference: the CPU oracle timed on the host cores; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import deepsc_gan_b200  # noqa: F401
    units = args.ref_units
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_oracle_rate(1)
    times = []
    for s in range(args.steps):
        rate, dt = cpu_oracle_rate(units, first_unit=s * units)
        times.append(dt)
    T = sum(times)
    value = 64 * units * args.steps / T
    sample = f"{units} units x {args.steps} steps of the workload (64 sentences each, SNR points cycled), KV-free oracle greedy"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(units, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(units: int, world: int):
    return {"workload": "Transeiver_Star (SE/SD, cycle_num=8, d_model=128, 8 heads, vocab 22234) AWGN SNR sweep 0-18 dB, "
                        "greedy decode 30 steps + BLEU counts",
            "units_per_step_per_gpu": units, "sentences_per_unit": 64, "seq_len": 31, "snr_points_db": "0..18",
            "sharding": f"{world} rank(s) x {units} units, unit-granular, no data-path collective",
            "greedy_loop": "one CUDA graph replay per step (30 decode steps)",
            "l2_policy": "per-step working set (activations + logits workspace) exceeds the 126 MB L2"}


def train_leg(cfg, dev, rank, world, bs, steps, barrier):
    """BASELINE.json configs[4]: one ``gan_train_step`` (Transeiver_GAN forward, two backward sweeps, three Adam
    applications; utlis/gan_train.py:8-50) per step on ``bs`` sentences per rank, CUDA-graph replayed; under torchrun
    the flat gradient bucket is all-reduced over NCCL once per step.  Device-timed, max over ranks."""
    import torch.distributed as dist
    from deepsc_gan_b200 import models, sweep
    from deepsc_gan_b200.dataset.synthetic import synthetic_units
    from deepsc_gan_b200.models import modules
    from deepsc_gan_b200.utlis import gan_train as GT
    modules.set_precision(1)
    torch.manual_seed(2024)                                        # identical initial weights on every rank
    net = models.Transeiver_GAN(cfg).to(dev).train()
    opt = GT.make_optimizer(net, learning_rate=cfg.lr)
    torch.manual_seed(100 + rank)                                  # per-rank noise and dropout
    modules.set_dropout_seed(7000 + rank)
    n_u = (bs + 63) // 64
    batches = [synthetic_units((rank * 64 + s) * n_u, n_u)[:bs].to(dev) for s in range(4)]
    step = GT.GraphedGanTrainStep(net, opt, 0.5, n_std=float(sweep.snr_to_noise(3.0)), traingan=True)
    for s in range(3):
        out = step(batches[s % 4])
    first = float(out[0])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        out = step(batches[s % 4])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    digest = opt.fp.flat.double().sum().reshape(1)
    same = True
    if world > 1:
        parts = [torch.empty_like(digest) for _ in range(world)]
        dist.all_gather(parts, digest)
        same = all(bool(torch.equal(q, parts[0])) for q in parts)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    res = {"workload": "Transeiver_GAN gan_train_step (forward + 2 backward sweeps + 3 Adam applies), AWGN 3 dB, graph-replayed",
           "batch_per_rank": bs, "n_gpus": world, "steps": steps, "ms_per_step": ms,
           "sentences_per_s": bs * world / (ms * 1e-3), "all_reduce_bytes_per_step": int(opt.fp.grad_bucket.numel()) * 4 if world > 1 else 0,
           "replicas_identical": same, "loss_first": first, "loss_last": float(out[0])}
    del step, net, opt
    torch.cuda.synchronize()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--units", type=int, default=37, help="64-sentence units per step per GPU")
    ap.add_argument("--ref-units", type=int, default=4, help="units per step of the CPU reference arm")
    ap.add_argument("--cpu-units", type=int, default=32, help="units of the cpu_baseline sample (N=1 only): ~15 s of CPU work")
    ap.add_argument("--prec", type=int, default=1, help="0 fp32 FFMA, 1 tcgen05 bf16x3, 2 tcgen05 bf16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager greedy loop in the timed legs (one launch per kernel)")
    ap.add_argument("--train-steps", type=int, default=20, help="timed steps of the config-5 training leg (0 = skip)")
    ap.add_argument("--train-bs-large", type=int, default=512, help="second training figure at this batch per rank (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    import deepsc_gan_b200  # noqa: F401
    from deepsc_gan_b200 import _lib, sweep
    from deepsc_gan_b200.dataset.synthetic import synthetic_units
    from deepsc_gan_b200.models import Transeiver_Star, modules
    from deepsc_gan_b200.utlis.parameters import para_config

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    modules.set_precision(args.prec)
    peaks = load_peaks()

    cfg = para_config([])
    torch.manual_seed(2024)
    net = Transeiver_Star(cfg).to(dev).eval()
    U, S = args.units, args.units * 64
    # the timed legs replay the 30-step greedy loop as one CUDA graph per step; the star kernel's own duration (roofline) is
    # taken afterwards from a short eager pass with events round every launch, outside the timed region
    runner = sweep.SweepRunner(net, U, channel="AWGN", seed=1234 + rank, graph=not args.no_graph)
    n_std_host = torch.tensor([sweep.snr_to_noise(SNRS[(rank * U + u) % 19]) for u in range(U)], dtype=torch.float32)
    n_std = n_std_host.to(dev)
    total_steps = args.warmup + args.steps
    # distinct inputs per step (and per rank); pinned on the host for the e2e leg
    host_inputs = [synthetic_units((rank * total_steps + s) * U, U).pin_memory() for s in range(total_steps)]
    dev_inputs = [h.to(dev) for h in host_inputs]
    counts_host = torch.empty((S, 10), dtype=torch.int32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(s):
        ids, counts = runner.run(dev_inputs[s], n_std)
        return sweep.gather_counts(counts)

    def step_e2e(s):
        inp = host_inputs[s].to(dev, non_blocking=True)
        ids, counts = runner.run(inp, n_std)
        table = sweep.gather_counts(counts)
        counts_host.copy_(counts, non_blocking=True)
        return table

    # ---- device-resident leg -------------------------------------------------------------------------------
    for s in range(args.warmup):
        step_resident(s)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    _lib.STATS["launches"] = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s in range(args.warmup, total_steps):
        step_resident(s)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    t_ms = ev0.elapsed_time(ev1)
    launches = _lib.STATS["launches"]

    # ---- end-to-end leg (host buffers in, counts out) ------------------------------------------------------
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.warmup, total_steps):
        step_e2e(s)
    e1.record()
    barrier()
    t_e2e_ms = e0.elapsed_time(e1)

    # ---- per-launch timing of the dominant kernel: eager pass over two of the same steps ------------------------
    eager = runner if args.no_graph else sweep.SweepRunner(net, U, channel="AWGN", seed=1234 + rank, graph=False)
    eager.run(dev_inputs[0], n_std)
    torch.cuda.synchronize()
    _lib.PROFILE = []
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for s in range(2):
        eager.run(dev_inputs[args.warmup + s % args.steps], n_std)
    pe1.record()
    torch.cuda.synchronize()
    prof, _lib.PROFILE = _lib.PROFILE, None
    t_prof_ms = pe0.elapsed_time(pe1)

    if world > 1:
        tt = torch.tensor([t_ms, t_e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_ms, t_e2e_ms = float(tt[0]), float(tt[1])

    # ---- roofline of the dominant kernel -------------------------------------------------------------------
    # star_fused_kernel (dsc_star_cycles_tc): all 8 cycles of a star layer in one launch, tile state in TMEM/smem, so
    # it is bound by the tensor pipe, not HBM.  Algorithmic FLOPs per sentence and cycle:
    #   32 rows x 128 x (384 [Wq|Wk|Wv]_sat + 128 Wo_sat + 256 [Wk|Wv]_relay) x 2  +  2 relay GEMVs x 128 x 128 x 2
    # (only the FLOPs a launch really executes are counted: see launch_flops)
    # Every fp32-class product costs three bf16 UMMA passes (bf16x3), so `achieved` counts 3 tensor FLOPs per
    # algorithmic FLOP and is compared with the measured bf16 peak (sustained: the kernel is timed inside a long step).
    SAT_HALF = 32 * 128 * (384 + 128) * 2          # J0..J4 of a cycle, per sentence
    RELAY_KV = 32 * 128 * 256 * 2                  # J5, J6
    GEMV = 128 * 128 * 2                           # J7 every cycle, J8 every cycle but the first

    def launch_flops(meta):
        """Algorithmic FLOPs of one dsc_star_cycles_tc launch; a greedy-step launch skips the satellite half of its
        first cycle (DSC_STAR_FIRST_SAT_DONE: computed once per batch) and the relay half of its last (DSC_STAR_NO_FINAL_RELAY)."""
        n_sent, cycles, _n2, first_sat_done, no_final_relay = meta
        per_sent = cycles * (SAT_HALF + RELAY_KV + GEMV) + (cycles - 1) * GEMV - (SAT_HALF if first_sat_done else 0)
        if no_final_relay:                  # the last cycle stops after J4: no relay K|V, no J7, and no J8 for it
            per_sent -= RELAY_KV + GEMV + (GEMV if cycles > 1 else 0)
        return float(per_sent) * n_sent

    passes = {1: 3, 2: 1}.get(args.prec, 0)
    dom = [(a.elapsed_time(b), meta) for op, a, b, meta in prof if op == "dsc_star_cycles_tc" and meta[0] == S]
    roof = None
    if dom and passes:
        mean_ms = sum(t for t, _ in dom) / len(dom)
        flops = sum(launch_flops(m) for _, m in dom) / len(dom)          # mean over the timed launches
        achieved = passes * flops / (mean_ms * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        roof = {"bound": "tensor", "kernel": "star_fused_kernel<3> via dsc_star_cycles_tc (8 star cycles per launch: QKV / Wo / relay "
                                             "K|V / relay update UMMAs + satellite and relay attention, state in TMEM)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": ROOFLINE_TRAFFIC_BYTES.get(S),
                "peak_source": f"{peaks['source']} cuBLAS bf16 sustained (MEASURED_PEAKS.json bf16_tflops_sustained)",
                "launches_timed": len(dom), "mean_launch_ms": mean_ms, "algorithmic_flops_per_launch": flops,
                "tensor_passes_per_flop": passes, "algorithmic_tflops": flops / (mean_ms * 1e-3) / 1e12,
                "frac_algorithmic": flops / (mean_ms * 1e-3) / 1e12 / peak,
                "tensor_active_pct": NCU_TENSOR_ACTIVE_PCT.get(S),
                "share_of_step": (sum(t for t, _ in dom) / 2) / (t_ms / args.steps),
                "timed_in": f"eager pass of 2 steps after the timed legs ({t_prof_ms / 2:.2f} ms per eager step)",
                "arith": {1: "bf16x3 tcgen05 (3 bf16 UMMA passes per fp32-class product), fp32 accumulate/softmax",
                          2: "bf16 tcgen05"}[args.prec]}
    elif prof:
        # per-cycle kernels (DSC_STAR_FUSED=0): star_sat_kernel is HBM-bound, 64,000 algorithmic bytes per sentence
        SAT_BYTES_PER_SENTENCE = 32 * 512 + 31 * 1024 + 31 * 512
        dom = [(a.elapsed_time(b), meta) for op, a, b, meta in prof if op == "dsc_star_sat_tc" and meta == S]
        if dom:
            mean_ms = sum(t for t, _ in dom) / len(dom)
            nbytes = float(SAT_BYTES_PER_SENTENCE) * S
            achieved = nbytes / (mean_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": "star_sat_kernel<3> via dsc_star_sat_tc (one launch per star cycle)",
                    "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                    "traffic": None, "peak_source": f"{peaks['source']} HBM copy bandwidth", "launches_timed": len(dom),
                    "mean_launch_ms": mean_ms, "bytes_per_launch": nbytes, "share_of_step": sum(t for t, _ in dom) / t_ms}

    # ---- config 5: the data-parallel GAN training step on the same ranks -----------------------------------
    train = None
    if args.train_steps > 0:
        train = {}
        for key, bs in (("bs64", 64), ("bs_large", args.train_bs_large)):
            if bs <= 0:
                continue
            try:
                train[key] = train_leg(cfg, dev, rank, world, bs, args.train_steps, barrier)
            except Exception as e:                                   # the eval line must survive a training failure
                train[key] = {"error": f"{type(e).__name__}: {e}"[:300]}
        modules.set_precision(args.prec)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    sentences = S * world * args.steps
    value = sentences / (t_ms * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {0: "f32", 1: "bf16x3 (fp32-class)", 2: "bf16"}[args.prec], "data": "synthetic",
            "config": workload_config(U, world), "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": sentences / (t_e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": S * 31 * 4,
                    "d2h_bytes_per_step": S * 10 * 4},
            "roofline": roof}
    if train is not None:
        line["train"] = train
    if world == 1 and not args.no_cpu_baseline:
        rate, dt = cpu_oracle_rate(args.cpu_units)
        n_aw = max(1, args.cpu_units // 4)
        rate_aw, dt_aw = cpu_oracle_rate(n_aw, last_only=False)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{args.cpu_units} units ({64 * args.cpu_units} sentences) of the same workload, oracle greedy "
                                          f"(last-position logits), {dt:.1f} s wall",
                                "as_written": {"value": rate_aw, "unit": UNIT, "cores": os.cpu_count(),
                                               "sample": f"{n_aw} units ({64 * n_aw} sentences), greedy as utlis/eval.py:106-113 writes it "
                                                         f"(vocabulary logits of the whole prefix every step), {dt_aw:.1f} s wall"}}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
