"""Copy the judged summaries of a tools/profile_round.sh pass from gpurun_out/ (scratch) into profiles/ (tracked).

    python tools/profile_collect.py r01d r01      # scratch tag -> profile name prefix
"""
import csv
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "Block Size", "Grid Size", "gpu__time_duration.sum", "sm__cycles_elapsed.max.per_second",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "lts__t_bytes.sum.per_second", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_wait.ratio",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "smsp__average_warp_latency_issue_stalled_branch_resolving.ratio",
    "smsp__average_warp_latency_issue_stalled_no_instruction.ratio", "smsp__average_warp_latency_issue_stalled_sleeping.ratio",
    "smsp__average_warp_latency_issue_stalled_membar.ratio", "smsp__average_warp_latency_issue_stalled_tex_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio", "smsp__average_warp_latency_issue_stalled_imc_miss.ratio",
    "smsp__average_warp_latency_issue_stalled_not_selected.ratio", "smsp__average_warp_latency_issue_stalled_selected.ratio",
    "smsp__average_warp_latency_issue_stalled_drain.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "local_load_bytes", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def ncu_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def summarise(rep, title, dst):
    hdr, units, launches = ncu_raw(rep)
    with open(dst, "w") as f:
        f.write(f"# {title}\n# source: {os.path.relpath(rep, ROOT)} (scratch, not tracked); ncu --set full --clock-control none --import-source on\n")
        for vals in launches:
            d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
            f.write(f"Kernel Name  {d['Kernel Name'][1]}\n")
            for k in KEEP:
                if k in d:
                    f.write(f"  {k:92s} {d[k][0]:16s} {d[k][1]}\n")
            f.write("---\n")


def main(tag, name):
    out, prof = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
    shutil.copy(os.path.join(out, f"{tag}_bench.json"), os.path.join(prof, f"{name}_bench_plain_run.json"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), os.path.join(out, f"{tag}_launches.csv")],
                       capture_output=True, text=True)
    with open(os.path.join(prof, f"{name}_launches_bench_step.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --train-steps 0\n"
                "# (per-launch times are cold-cache and serialised: the SHARE per kernel is what is compared with bench.py's roofline.share_of_step)\n")
        f.write(r.stdout.replace(out + "/", "gpurun_out/"))
    summarise(os.path.join(out, f"{tag}_star_fused.ncu-rep"),
              "star_fused_kernel<3>: tools/prof_star.py 1 2368 8 (2,368 sentences = 592 tiles, 8 cycles, n2 = 17), B200", os.path.join(prof, f"{name}_ncu_star_fused.txt"))
    # the other kernels of a bench step (tools/profile_round.sh `kernels`: one capture each out of `bench.py --steps 1 --warmup 1`)
    for f, title in (("vocab_argmax_tc_", "vocab_argmax_tc_kernel<3>: one greedy step of bench.py (2,368 rows x 22,234 vocabulary entries)"),
                     ("tar_tail_", "tar_tail_kernel<3>: Dense + residual + LayerNorm + relay k|v projection + key-cache write of a greedy step (2,368 rows)"),
                     ("mha_decode_attention_", "mha_decode_attention_kernel: single-query attention of a greedy step (2,368 sentences x 8 heads)"),
                     ("add_layernorm_", "add_layernorm_kernel: the two final LayerNorms of a greedy step (2,368 rows)"),
                     ("gemm_k128_", "gemm_k128_persistent_kernel<3>: tools/time_linear.py, 73,408 x 128 -> 256 Dense + ReLU"),
                     ("gemm_tc_kernel_128__3__2", "gemm_tc_kernel<128,3,2>: the q|k|v projection of a greedy step (2,368 x 128 -> 384)")):
        rep = os.path.join(out, f"{tag}_k_{f}.ncu-rep")
        if os.path.exists(rep):
            summarise(rep, title + ", B200", os.path.join(prof, f"{name}_ncu_{f.strip('_')}.txt"))
    for extra in ("smoke.log", "bench.err"):
        if os.path.exists(os.path.join(out, f"{tag}_{extra}")):
            shutil.copy(os.path.join(out, f"{tag}_{extra}"), os.path.join(prof, f"{name}_{extra.replace('.', '_run.') if extra == 'bench.err' else extra}"))
    log = os.path.join(out, f"{tag}_pytest_gpu.log")
    if os.path.exists(log):
        shutil.copy(log, os.path.join(prof, f"{name}_pytest_gpu.log"))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "r01")
