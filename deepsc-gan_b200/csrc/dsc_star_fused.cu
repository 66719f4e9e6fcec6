// All cycles of a star layer in ONE persistent launch (K2+K3+K4 of SURVEY.md 2.3, models/modules.py:283-306, 359-378).
//
// The per-cycle kernels of dsc_star_tc.cu stream the node states X, the attention output ATT and the relay vectors
// through HBM twice per cycle.  Here a CTA keeps one tile (4 sentences = 128 rows = 128 TMEM lanes) on chip for all
// cycles: X and ATT live in tensor memory as the bf16 hi/lo A operands of the next UMMA, the relay vectors s / q in
// shared memory, and only the constant keys (e-keys KVEI, cached target keys KV2I) are re-read - from L2, since a
// tile's 256 KB of keys are touched every cycle.  The five weight matrices (512 KB as bf16 hi/lo planes) do not fit
// in shared memory, so a producer thread streams them per job through a shared-memory ring with cp.async.bulk.
//
// Jobs of one cycle (one elected thread issues the UMMAs, fp32 accumulators in TMEM):
//   J0..J3  QKV of head pair g = X @ [Wq|Wk|Wv]_sat[g]          M = 128, N = 96  -> satellite attention over the 5 keys
//                                                                          {h[i+1], h[i], h[i-1], e[i], s} -> ATT (TMEM)
//   J4      O   = ATT @ Wo_sat                                   N = 128 -> X' = relu(O + b) (relay row keeps s) -> X (TMEM)
//   J5, J6  K|V = X' @ [Wk|Wv]_relay                              N = 128 each -> relay attention over 32 (+n2) keys -> att_r
//   J7      R1^T = Wo_relay^T @ att_r^T                           M = 128 features, N = 16 (4 sentences) -> s' = relu(R1 + b)
//   J8      R2^T = Wq_relay^T @ s'^T                              same shape -> q' (the next cycle's relay query)
// J7/J8 are GEMVs (one row per sentence).  As X @ W they would spend a full M = 128, N = 128 UMMA on 4 useful rows, so
// they run transposed: the streamed weight chunk (K-major, the same image either way) is the A operand and the four
// vectors, written as bf16 hi/lo into a 16-row K-major shared-memory operand, are B.  The result lands feature-major
// (TMEM lane = feature, column = sentence): each compute thread reads ONE accumulator word, and s' / q' return to the
// sentence-major consumers through shared memory and a named barrier of the compute warps.
// Accumulators alternate between two 128-column TMEM buffers (J7 -> buffer 0, J8 -> buffer 1), so the UMMA of job j+1
// overlaps the epilogue of job j wherever the data dependence allows (all of J0..J3, J5/J6, and J0 of the next cycle
// under J8).
//
// Where the time goes (tools/star_trace.py: clock64 stamps of CTA 0 at every hand-off; 2,368 sentences, 8 cycles,
// ~450 us, 12.1 us per tile and cycle): J0..J3 until ATT is staged ~5 us (2.5 us of tensor time; the satellite attention
// of the 8 heads is ~3.9 us of instruction-issue-bound CUDA-core work: 460 instructions per warp and head, of which
// 166 FFMA and 96 SHFL), J4 + its epilogue ~1.9 us, J5/J6 + the relay attention ~4.0 us, J7 + the relay-row patch ~1.8 us.
// Tensor time per tile and cycle is ~5.9 us at 1.965 GHz (24 UMMAs per job at 58 / 74 cycles for N = 96 / 128, ~45
// cycles for the N = 16 GEMVs): the pipe is busy ~37 % of the time (ncu sm__pipe_tensor_cycles_active), the compute
// warps are issue-bound inside each epilogue burst (smsp__issue_active 36 % on average) and wait while the UMMAs run.
// What moved the needle in round 1 (537 -> 470 us per launch, and two skipped half-cycles per greedy step): transposed
// relay GEMVs; QKV accumulators released right after the q|k|v read; the relay query issued behind J3 into spare columns;
// J4 started on the first K-block of ATT; biases and the relay-lane select from shared memory; suspend-time hints on the
// mbarrier polls (the polling loops were a quarter of all issued instructions); a 2-stage instead of a 3-stage weight
// ring (470 -> 450 us: the smaller shared-memory carve-out leaves L1 for the key streams and the spill slots).
// Tried and dropped, all parity-correct: (a) 8 instead of 16 compute warps: slower; (b) two tiles per CTA on the same
// warps with ATT held in registers: 1.9x slower (spills); (c) two tile pipelines per CTA (8 warps each, ATT in shared
// memory, single accumulator per tile, shared weight ring, in-order UMMA issue): 31 us per PAIR and cycle = exactly two
// sequential tiles - the epilogues of both tiles contend for the same four schedulers, so overlapping them with the
// other tile's UMMAs buys nothing while the per-tile double buffering is lost; (d) folding the relay-query GEMV into the
// K|V phase: 6 % slower; (e) 112 registers per thread: not launchable (warps are allocated in fours: 18 warps count as
// 20, 20 x 32 x 112 > 64 K registers); (f) both relay-attention heads of a warp in one 32-column pass: more spills,
// 5 % slower; (g) L2 prefetch of the next tile's node states and keys (cp.async.bulk.prefetch.L2): no change or worse;
// (h) J4's K-block 1 as two column halves with their own commits + J5 issued per K-block: no change (the J4 epilogue
// is issue-bound, staggering its halves does not shorten it).
// Next: the bound is the CUDA-core instruction stream of the two attentions between dependent UMMAs.  A second tile in
// flight only helps with a second set of warps (register file: 96 x 1,152 threads does not fit) or with fewer
// instructions per epilogue; candidates are the neighbour exchange through shared memory instead of 96 shuffles per
// warp and head, and a 2-CTA cluster (cta_group::2) that halves the weight stream per SM and frees shared memory.
//
// Warps: 0-15 compute (quarter = warp & 3 is the sentence / TMEM lane quarter, sub = warp >> 2 the column quarter),
// 16 = UMMA issuer, 17 = weight producer.  Every warp owns 32 of the 128 columns of a row.  Arithmetic as in
// dsc_star_tc.cu: prec 1 = bf16x3, 2 = bf16; fp32 softmax.
#include "dsc_star_common.cuh"

namespace dsc {

using namespace tc;

namespace sf {
// Register reallocation: 20 warps = 5 warpgroups (16 compute warps; issuer, producer and two idle warps).  At launch every
// thread has 96 registers (65,536 / 640); the last warpgroup gives back all but 32 (setmaxnreg.dec) and the compute
// warpgroups take them (setmaxnreg.inc 112: (112 - 96) x 512 = (96 - 32) x 128, the pool is exact).  At 112 registers the
// compute code has no local-memory traffic at all (96: 140 B of spill stores / 212 B of loads per thread, whose misses go to
// L2 because the key streams thrash L1): 447 -> 422 us per 8-cycle launch, a greedy-step launch 394 -> 369 us.
// SF_PRED_KV2: lanes whose target-key row is masked (lane >= n2) do not fetch it.
// SF_RELAY_PAIR: the relay attention reduces both heads of a warp together - the two softmax reductions and the two 31-shuffle
// reduce-scatters are independent dependency chains, interleaved instruction by instruction (the loop form serialises them:
// shuffles are not reordered across the tcgen05.ld of the second head).  With 112 registers this no longer spills (round 1: it
// did, and was 5 % slower): 419 -> 409 us per 8-cycle launch, greedy-step launch 368 -> 359 us, results bit-identical.
#ifndef SF_SETMAXNREG
#define SF_SETMAXNREG 1
#endif
#ifndef SF_COMPUTE_REGS
#define SF_COMPUTE_REGS 112
#define SF_SIDE_REGS 32
#endif
#ifndef SF_PRED_KV2
#define SF_PRED_KV2 1
#endif
#ifndef SF_ISSUER_SPIN
#define SF_ISSUER_SPIN 1
#endif
#if SF_ISSUER_SPIN
#define SF_ISSUER_WAIT mbar_wait_spin
#else
#define SF_ISSUER_WAIT mbar_wait
#endif
#ifndef SF_RELAY_PAIR
#define SF_RELAY_PAIR 1
#endif
constexpr int kCompute = 16, kMmaWarp = 16, kProdWarp = 17, kThreads = SF_SETMAXNREG ? 640 : 576;
constexpr uint32_t ACC0 = 0, ACC1 = 128, AX_HI = 256, AX_LO = 320, AT_HI = 384, AT_LO = 448;
constexpr uint32_t ACC_Q = ACC1 + 96;                         // J8's 16 columns: behind the 96 columns of a QKV job
// Issue order of a cycle: J0..J3, then J8 (the query q' of THIS cycle's relay attention, from the previous cycle's s';
// absent in cycle 0, whose query is an input), then J4..J7.  J8 sits behind the QKV jobs so that it runs while the
// compute warps are still busy with the satellite attention, instead of delaying J0.
// skip0: the satellite phase of cycle 0 (J0..J4) was computed by an earlier launch (it depends on the e tile only, not on
// the decoded prefix): x_tile0 holds X' and cycle 0 is J5, J6, J7.
// nfr ("no final relay"): the caller does not read the relay row of the result (a greedy decoder reads satellite rows only),
// so the last cycle stops after J4 and needs no relay query either.
struct Seq {
  int n_cycles; bool skip0, nfr;
  __device__ __forceinline__ bool has_relay(int c) const { return !(nfr && c + 1 == n_cycles); }
  __device__ __forceinline__ bool has_j8(int c) const { return c > 0 && has_relay(c); }
  __device__ __forceinline__ int jobs(int c) const {
    const int sat = (c == 0 && skip0) ? 0 : 5;                 // J0..J4
    return sat + (has_j8(c) ? 1 : 0) + (has_relay(c) ? 3 : 0);  // J8, J5..J7
  }
  __device__ __forceinline__ int job_at(int c, int i) const {
    if (c == 0 && skip0) return 5 + i;
    if (i < 4) return i;
    if (has_j8(c)) return i == 4 ? 8 : i - 1;
    return i;
  }
};
struct Bars {
  uint64_t w_full[STAGES], w_free[STAGES];
  uint64_t acc_full[2], acc_free[2];
  uint64_t x_ready, t_ready;
  uint64_t ta_ready;                        // the first half of ATT (heads 0..3 = K-block 0 of J4's operand) is staged
  uint64_t q_full;                          // the relay query q' (J8) is in its accumulator columns
};
}  // namespace sf

// Debug timeline (tools/star_trace.py): when a buffer is registered with dsc_debug_star_trace, CTA 0 stamps clock64()
// at the hand-offs of its first tile: issuer [0, 256) = (operands ready, UMMAs + commits issued) per job; compute warp 0
// [256, 512) and warp 8 [512, 768) = their accumulator-seen / operand-staged events in program order.
#ifdef DSC_DEBUG_TOOLS
__device__ unsigned long long* g_star_trace = nullptr;
#else
#define g_star_trace ((unsigned long long*)nullptr)
#endif
#define DSC_TR(base) do { if (TRACE) { if (tr_buf && tr_n < 256) tr_buf[(base) + tr_n] = (unsigned long long)clock64(); ++tr_n; } } while (0)

template <int NPASS, bool TRACE>
__global__ void __launch_bounds__(sf::kThreads, 1)
star_fused_kernel(const float* __restrict__ XI0, const float* __restrict__ S0, const float* __restrict__ Q0,
                  const float* __restrict__ KVEI, const float* __restrict__ KV2I, int n2, sf::Weights W,
                  const float* __restrict__ bias_o, const float* __restrict__ bias_r,
                  float* __restrict__ Xrow, int n_tiles, int n_cycles, int flags) {
  using namespace sf;
  const bool skip0 = (flags & 1) != 0, nfr = (flags & 2) != 0;
  const Seq seq{n_cycles, skip0, nfr};
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* rb = ring + STAGES * STAGE_BYTES;         // att_r, then s': B operand of the transposed relay GEMVs (rows 4..15 zero)
  __shared__ __align__(8) sf::Bars bars;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_cur[4][128];      // relay node of the 4 sentences of the tile
  __shared__ __align__(16) float q_cur[4][128];      // its query under the relay weights
  __shared__ __align__(16) float bias_s[2][128];     // [satellite dense | relay dense] biases: L1 is thrashed by the key streams
  __shared__ __align__(16) uint32_t patch_w[2][4][4][16];   // [hi|lo][sentence][column quarter]: bf16 pairs of s' for the X patch
  constexpr int parts = (NPASS == 3) ? 2 : 1;
  constexpr uint32_t kArrivals = kCompute;            // one elected arrival per compute warp (after __syncwarp)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned long long* tr_buf = (TRACE && blockIdx.x == 0 && (tid == sf::kMmaWarp * 32 || tid == 0 || tid == 256)) ? g_star_trace : nullptr;
  int tr_n = 0;
  for (uint32_t i = tid; i < RB_BYTES / 16; i += sf::kThreads) reinterpret_cast<uint4*>(rb)[i] = make_uint4(0, 0, 0, 0);
  fence_async_smem();
  if (tid < 128) bias_s[0][tid] = __ldg(bias_o + tid);
  else if (tid < 256) bias_s[1][tid - 128] = __ldg(bias_r + tid - 128);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bars.w_full[s], 1); mbar_init(&bars.w_free[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&bars.acc_full[b], 1); mbar_init(&bars.acc_free[b], kArrivals); }
    mbar_init(&bars.x_ready, kArrivals);
    mbar_init(&bars.t_ready, kArrivals);
    mbar_init(&bars.ta_ready, kArrivals);
    mbar_init(&bars.q_full, 1);
    fence_barrier_init();
  }
  if (warp == sf::kMmaWarp) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

#if SF_SETMAXNREG
  if (warp >= kCompute) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(SF_SIDE_REGS));
#endif
  if (warp > kProdWarp) {
  } else if (warp == kProdWarp) {
    // ------------------------------------------------------------------ weight producer
    if (lane == 0) {
      uint32_t n = 0;                                   // chunks issued so far
      for (int t = 0; t < my_tiles; ++t)
        for (int c = 0; c < n_cycles; ++c) {
          const int jobs = seq.jobs(c);
          for (int ji = 0; ji < jobs; ++ji, ++n) {
            const int j = seq.job_at(c, ji);
            const uint32_t st = n % STAGES;
            mbar_wait(&bars.w_free[st], ((n / STAGES) - 1) & 1);
            const uint8_t* blob; uint32_t rows, row0, n_pad;
            chunk_of(W, j, blob, rows, row0, n_pad);
            const uint32_t plane = rows * 128u;
            mbar_expect_tx(&bars.w_full[st], parts * 2 * plane);
            for (int p = 0; p < parts * 2; ++p)
              bulk_g2s(ring + st * STAGE_BYTES + p * plane, blob + ((size_t)p * n_pad + row0) * 128, plane, &bars.w_full[st]);
          }
        }
    }
    __syncwarp();
  } else if (warp == sf::kMmaWarp) {
    // ------------------------------------------------------------------ UMMA issuer
    // The whole warp runs the job loop in lock-step and ONE elected lane issues: tcgen05.mma / commit issued from a
    // divergent `if (lane == 0)` region cost ~94 cycles each (compiler-generated elect loop), from here 58 (N = 96) and
    // 74 (N = 128) cycles (tools/umma_probe.py).
    {
      const bool leader = elect_one();
      const uint32_t ring_base = smem_u32(ring), rb_base = smem_u32(rb);
      uint32_t n = 0, use0 = 0, use1 = 0, xr = 0, tr = 0, ta = 0;  // chunks consumed, accumulator uses, operand phases consumed
      for (int t = 0; t < my_tiles; ++t)
        for (int c = 0; c < n_cycles; ++c) {
          const int jobs = seq.jobs(c);
          for (int ji = 0; ji < jobs; ++ji, ++n) {
            const int j = seq.job_at(c, ji);
            const uint32_t st = n % STAGES, b = (j < 7) ? ((uint32_t)j & 1u) : 0u;                    // J7 -> ACC0
            if (j == 0 || j == 5) { SF_ISSUER_WAIT(&bars.x_ready, xr & 1); ++xr; }          // X staged / X' restaged
            if (j == 7 || j == 8) { SF_ISSUER_WAIT(&bars.t_ready, tr & 1); ++tr; }          // att_r / s' staged
            if (j == 4) { SF_ISSUER_WAIT(&bars.ta_ready, ta & 1); ++ta; }                   // first half of ATT staged
            SF_ISSUER_WAIT(&bars.w_full[st], (n / STAGES) & 1);
            if (j != 8) {                                                              // J8 has its own columns (ACC_Q); every
              if (b) { SF_ISSUER_WAIT(&bars.acc_free[1], (use1 - 1) & 1); ++use1; }         // warp read the previous q' before it
              else   { SF_ISSUER_WAIT(&bars.acc_free[0], (use0 - 1) & 1); ++use0; }         // freed this cycle's first accumulators
            }
            tc_fence_after();
            if (t == 0) DSC_TR(0);
            const uint32_t bb = ring_base + st * STAGE_BYTES;
            const uint32_t acc = b ? ACC1 : ACC0;
            if (leader) {
              if (j < 4) issue_group<NPASS, 96>(tmem_base, acc, AX_HI, AX_LO, bb, 96u * 128u, 0u);
              else if (j == 4) issue_group_kb<NPASS, 128>(tmem_base, acc, AT_HI, AT_LO, bb, 128u * 128u, 0, true);
              else if (j == 7) issue_relay_gemv<NPASS>(tmem_base + acc, bb, rb_base);
              else if (j == 8) issue_relay_gemv<NPASS>(tmem_base + ACC_Q, bb, rb_base);
              else issue_group<NPASS, 128>(tmem_base, acc, AX_HI, AX_LO, bb, 128u * 128u, 0u);
            }
            if (j == 4) {                                     // second half of ATT (heads 4..7), then the K-block 1 UMMAs
              SF_ISSUER_WAIT(&bars.t_ready, tr & 1); ++tr;
              tc_fence_after();
              if (leader) issue_group_kb<NPASS, 128>(tmem_base, acc, AT_HI, AT_LO, bb, 128u * 128u, 1, false);
            }
            if (leader) {
              umma_commit(&bars.w_free[st]);
              umma_commit(j == 8 ? &bars.q_full : &bars.acc_full[b]);
            }
            if (t == 0) DSC_TR(0);
            __syncwarp();
          }
        }
    }
  } else {
    // ------------------------------------------------------------------ compute warps (16)
#if SF_SETMAXNREG
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(SF_COMPUTE_REGS));
#endif
    // quarter = TMEM lane quarter = sentence of the tile; sub = column quarter (32 of the 128 columns, = heads 2*sub,
    // 2*sub+1 of the relay attention).  In the QKV phase sub also picks the accumulator: warps with sub < 2 take the
    // even head pairs (ACC0), the others the odd ones (ACC1), head 2g + (sub & 1) each.
    const int quarter = warp & 3, sub = warp >> 2;
    const int gp = sub >> 1, hh = sub & 1;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int row_in_tile = quarter * 32 + lane;
    const int up = (lane >= 30) ? 0 : lane + 1;                    // roll(h,-1)[i] = h[(i+1) mod 31]
    const int dn = (lane == 0) ? 30 : lane - 1;                    // roll(h,+1)[i] = h[(i-1) mod 31]
    float* my_s = &s_cur[quarter][sub * 32];
    float* my_q = &q_cur[quarter][sub * 32];
    uint32_t use0 = 0, use1 = 0;                                   // accumulator phases seen (scalars: no local memory)
    auto wait_acc = [&](int b) {
      if (b) { mbar_wait(&bars.acc_full[1], use1 & 1); ++use1; } else { mbar_wait(&bars.acc_full[0], use0 & 1); ++use0; }
      tc_fence_after();
      DSC_TR(warp ? 512 : 256);
    };
    // every lane fences its own TMEM accesses, the warp converges, one lane arrives for the warp
    auto warp_arrive = [&](uint64_t* bar, uint32_t n) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_n(bar, n);
      DSC_TR(warp ? 512 : 256);
    };
    auto free_acc = [&](int b) { warp_arrive(&bars.acc_free[b], 1); };

    for (int ti = 0; ti < my_tiles; ++ti) {
      if (ti) tr_buf = nullptr;
      const int t = blockIdx.x + ti * gridDim.x;
      const int64_t sent = (int64_t)t * 4 + quarter;
      const uint4* kve_base = reinterpret_cast<const uint4*>(KVEI + (int64_t)t * 32768) + row_in_tile;
      uint32_t kv[32];                                             // e-keys k[16] | v[16] of (row, head) as raw fp32 bits
      auto load_kve = [&](int g) {
        const int head = 2 * g + hh;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 a = __ldg(kve_base + (head * 4 + q) * 128), c = __ldg(kve_base + (32 + head * 4 + q) * 128);
          kv[4*q] = a.x; kv[4*q+1] = a.y; kv[4*q+2] = a.z; kv[4*q+3] = a.w;
          kv[16+4*q] = c.x; kv[16+4*q+1] = c.y; kv[16+4*q+2] = c.z; kv[16+4*q+3] = c.w;
        }
      };
      // ---- tile start: relay vectors to shared memory, X rows to tensor memory (all UMMAs of the previous tile are
      //      complete: their last accumulator was drained below)
      {
        const float4* s0 = reinterpret_cast<const float4*>(S0 + sent * 128 + sub * 32);
        const float4* q0 = reinterpret_cast<const float4*>(Q0 + sent * 128 + sub * 32);
        if (lane < 8) reinterpret_cast<float4*>(my_s)[lane] = __ldg(s0 + lane);
        else if (lane < 16) reinterpret_cast<float4*>(my_q)[lane - 8] = __ldg(q0 + (lane - 8));
        __syncwarp();
        uint32_t hi[16], lo[16];
        if (lane == 31) split_quarter_row(my_s, hi, lo);
        else load_quarter_row(reinterpret_cast<const float4*>(XI0 + (int64_t)t * 16384) + (sub * 8) * 128 + row_in_tile, 128, hi, lo);
        store_quarter_row<NPASS>(lane_addr, AX_HI, AX_LO, sub, hi, lo);
        tmem_st_wait();
        warp_arrive(&bars.x_ready, 1);
      }
      if (!skip0) load_kve(gp);
      const int j8_per_tile = (n_cycles > 1) ? n_cycles - 1 - (nfr ? 1 : 0) : 0;

      for (int c = 0; c < n_cycles; ++c) {
        const bool last = (c + 1 == n_cycles);
        if (!(skip0 && c == 0)) {
        // ================= J0..J3: this warp takes head pairs g = gp and gp + 2 (accumulator gp), head 2g + hh
#pragma unroll
        for (int gi = 0; gi < 2; ++gi) {
          const int g = gp + 2 * gi, head = 2 * g + hh;
          wait_acc(gp);
          const uint32_t col = lane_addr + (gp ? ACC1 : ACC0) + hh * 16;
          float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f, l4 = 0.f;
          float v[16];
          {
            // q, k, v of (row, head) to registers in one go, then the accumulator goes straight back to the issuer: holding
            // it through the logits and the softmax (0.75 us, longer than a QKV UMMA) serialised J2 / J3 behind J0 / J1
            float q[16], k[16];
            tmem_ld16(col, q);
            tmem_ld16(col + 32, k);
            tmem_ld16(col + 64, v);
            tmem_ld_wait();
            warp_arrive(&bars.acc_free[gp], 2);                     // 8 of the 16 warps drain a QKV accumulator
#pragma unroll
            for (int d = 0; d < 16; ++d) {
              const float ku = __shfl_sync(0xffffffffu, k[d], up);
              const float kd = __shfl_sync(0xffffffffu, k[d], dn);
              const float ks = __shfl_sync(0xffffffffu, k[d], 31);
              l0 = fmaf(q[d], ku, l0);
              l1 = fmaf(q[d], k[d], l1);
              l2 = fmaf(q[d], kd, l2);
              l3 = fmaf(q[d], __uint_as_float(kv[d]), l3);
              l4 = fmaf(q[d], ks, l4);
            }
          }
          l0 *= 0.25f; l1 *= 0.25f; l2 *= 0.25f; l3 *= 0.25f; l4 *= 0.25f;
          const float mx = fmaxf(fmaxf(fmaxf(l0, l1), fmaxf(l2, l3)), l4);
          l0 = expf(l0 - mx); l1 = expf(l1 - mx); l2 = expf(l2 - mx); l3 = expf(l3 - mx); l4 = expf(l4 - mx);
          const float inv = 1.0f / (l0 + l1 + l2 + l3 + l4);
          l0 *= inv; l1 *= inv; l2 *= inv; l3 *= inv; l4 *= inv;
          uint32_t ohi[8], olo[8];
#pragma unroll
          for (int d2 = 0; d2 < 8; ++d2) {
            float o2[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int d = 2 * d2 + e;
              const float vu = __shfl_sync(0xffffffffu, v[d], up);
              const float vd = __shfl_sync(0xffffffffu, v[d], dn);
              const float vs = __shfl_sync(0xffffffffu, v[d], 31);
              float acc = l0 * vu;
              acc = fmaf(l1, v[d], acc);
              acc = fmaf(l2, vd, acc);
              acc = fmaf(l3, __uint_as_float(kv[16 + d]), acc);
              acc = fmaf(l4, vs, acc);
              o2[e] = (lane == 31) ? 0.f : acc;                    // the relay row carries no satellite output
            }
            split2(o2[0], o2[1], ohi[d2], olo[d2]);
          }
          // head `head` covers k = 16*head .. 16*head+15 = operand columns 8*head .. 8*head+7
          tmem_st8(lane_addr + AT_HI + head * 8, ohi);
          if (NPASS == 3) tmem_st8(lane_addr + AT_LO + head * 8, olo);
          if (gi == 0) {
            tmem_st_wait();
            warp_arrive(&bars.ta_ready, 1);                         // heads 0..3 = K-block 0 of J4's operand
            load_kve(g + 2);                                        // e-keys of this warp's second head pair
          }
        }
        if (gp) use0 += 2; else use1 += 2;                          // the two QKV jobs drained by the other warps
        tmem_st_wait();
        warp_arrive(&bars.t_ready, 1);
        if (seq.has_j8(c)) {
          // J8 (transposed, issued behind J3): q'[sentence sub][feature 32*quarter + lane] = s' @ Wq_relay
          mbar_wait(&bars.q_full, (uint32_t)(c - 1 + ti * j8_per_tile) & 1u);
          tc_fence_after();
          const float qv = tmem_ld1(lane_addr + ACC_Q + sub);
          tmem_ld_wait();
          q_cur[sub][quarter * 32 + lane] = qv;                      // read after the barrier at the head of the J5 block
        }

        // ================= J4: X' = relu(ATT @ Wo + b), columns 32*sub..; the relay row keeps s; re-staged as X
        {
          wait_acc(0);
          float v[32];
          tmem_ld32(lane_addr + ACC0 + sub * 32, v);
          tmem_ld_wait();
          free_acc(0);
          {
            // warp-uniform code (no divergent relay-lane branch): every lane reads the bias and its sentence's s from
            // shared memory (broadcast loads), the relay lane keeps s
            const bool relay_lane = (lane == 31);
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) {
              const float4 b4 = reinterpret_cast<const float4*>(&bias_s[0][sub * 32])[q4];
              const float4 s4 = reinterpret_cast<const float4*>(my_s)[q4];
              v[4*q4]   = relay_lane ? s4.x : fmaxf(v[4*q4]   + b4.x, 0.f);
              v[4*q4+1] = relay_lane ? s4.y : fmaxf(v[4*q4+1] + b4.y, 0.f);
              v[4*q4+2] = relay_lane ? s4.z : fmaxf(v[4*q4+2] + b4.z, 0.f);
              v[4*q4+3] = relay_lane ? s4.w : fmaxf(v[4*q4+3] + b4.w, 0.f);
            }
          }
          if (last) {
            float4* xr = reinterpret_cast<float4*>(Xrow + ((int64_t)t * 128 + row_in_tile) * 128 + sub * 32);
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) xr[q4] = make_float4(v[4*q4], v[4*q4+1], v[4*q4+2], v[4*q4+3]);
          }
          DSC_TR(warp ? 512 : 256);
          if (seq.has_relay(c)) {
            uint32_t hi[16], lo[16];
            split_quarter_row(v, hi, lo);
            DSC_TR(warp ? 512 : 256);
            store_quarter_row<NPASS>(lane_addr, AX_HI, AX_LO, sub, hi, lo);   // J0..J3 have completed (their commit precedes J4's)
            tmem_st_wait();
            DSC_TR(warp ? 512 : 256);
            warp_arrive(&bars.x_ready, 1);
          }
        }
        }   // !(skip0 && c == 0)
        if (!seq.has_relay(c)) break;                                // nfr: the result's relay row is not wanted

        // ================= J5 (K -> ACC1), J6 (V -> ACC0): relay attention, heads 2*sub and 2*sub+1, lane = key row
        {
          // h2 keys / values of this lane's row: KV2I [sentence][64 k4][32 rows][4]; k columns are k4 0..31, v 32..63.
          // Row `lane` < 32 is always inside KV2I; rows >= n2 are masked in the arithmetic (unconditional loads keep the
          // arrays in registers).
          const float4* kv2 = reinterpret_cast<const float4*>(KV2I + sent * 8192) + (sub * 8) * 32 + lane;
          const bool has2 = lane < n2;
          float4 k2[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) k2[i] = (!SF_PRED_KV2 || has2) ? __ldg(kv2 + i * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
          float w1[2], w2[2];
          compute_warps_sync();                                      // q' (written feature-major by the J8 epilogue) is complete
          wait_acc(1);
#if SF_RELAY_PAIR
          // both heads of the warp in one pass: the two softmax reductions (and below the two 31-shuffle reduce-scatters) are
          // independent dependency chains, interleaved instruction by instruction; same arithmetic per head as the loop form
          {
            float d1[2] = {0.f, 0.f}, d2[2] = {0.f, 0.f};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float k[16];
              tmem_ld16(lane_addr + ACC1 + sub * 32 + h * 16, k);
              tmem_ld_wait();
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                const float4 qq = reinterpret_cast<const float4*>(my_q + h * 16)[q4];
                d1[h] = fmaf(qq.x, k[4*q4], d1[h]); d1[h] = fmaf(qq.y, k[4*q4+1], d1[h]);
                d1[h] = fmaf(qq.z, k[4*q4+2], d1[h]); d1[h] = fmaf(qq.w, k[4*q4+3], d1[h]);
                const float4 kk = k2[h * 4 + q4];
                d2[h] = fmaf(qq.x, kk.x, d2[h]); d2[h] = fmaf(qq.y, kk.y, d2[h]); d2[h] = fmaf(qq.z, kk.z, d2[h]); d2[h] = fmaf(qq.w, kk.w, d2[h]);
              }
            }
            d1[0] *= 0.25f; d1[1] *= 0.25f;
            d2[0] = has2 ? d2[0] * 0.25f : -3.4e38f;
            d2[1] = has2 ? d2[1] * 0.25f : -3.4e38f;
            float m0 = fmaxf(d1[0], d2[0]), m1 = fmaxf(d1[1], d2[1]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float t0 = __shfl_xor_sync(0xffffffffu, m0, o), t1 = __shfl_xor_sync(0xffffffffu, m1, o);
              m0 = fmaxf(m0, t0); m1 = fmaxf(m1, t1);
            }
            const float e10 = expf(d1[0] - m0), e20 = has2 ? expf(d2[0] - m0) : 0.f;
            const float e11 = expf(d1[1] - m1), e21 = has2 ? expf(d2[1] - m1) : 0.f;
            float s0 = e10 + e20, s1 = e11 + e21;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float t0 = __shfl_xor_sync(0xffffffffu, s0, o), t1 = __shfl_xor_sync(0xffffffffu, s1, o);
              s0 += t0; s1 += t1;
            }
            const float i0 = 1.0f / s0, i1 = 1.0f / s1;
            w1[0] = e10 * i0; w2[0] = e20 * i0;
            w1[1] = e11 * i1; w2[1] = e21 * i1;
          }
          DSC_TR(warp ? 512 : 256);
#pragma unroll
          for (int i = 0; i < 8; ++i) k2[i] = (!SF_PRED_KV2 || has2) ? __ldg(kv2 + (32 + i) * 32) : make_float4(0.f, 0.f, 0.f, 0.f);      // the values take the keys' registers
          free_acc(1);
          wait_acc(0);
          {
            float p[32];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float v[16];
              tmem_ld16(lane_addr + ACC0 + sub * 32 + h * 16, v);
              tmem_ld_wait();
              if (h == 1) free_acc(0);
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                float4 vv = k2[h * 4 + q4];
                if (!has2) vv = make_float4(0.f, 0.f, 0.f, 0.f);                // masked rows may hold anything
                p[h*16 + 4*q4]     = fmaf(w1[h], v[4*q4],     w2[h] * vv.x);
                p[h*16 + 4*q4 + 1] = fmaf(w1[h], v[4*q4 + 1], w2[h] * vv.y);
                p[h*16 + 4*q4 + 2] = fmaf(w1[h], v[4*q4 + 2], w2[h] * vv.z);
                p[h*16 + 4*q4 + 3] = fmaf(w1[h], v[4*q4 + 3], w2[h] * vv.w);
              }
            }
            // sum over the 32 key lanes, both heads at once: fold the half-warps, then reduce-scatter 16 values over 16 lanes
#pragma unroll
            for (int i = 0; i < 32; ++i) p[i] += __shfl_xor_sync(0xffffffffu, p[i], 16);
#pragma unroll
            for (int off = 8, n = 8; off >= 1; off >>= 1, n >>= 1) {
              const bool upper = (lane & off) != 0;
#pragma unroll
              for (int i = 0; i < n; ++i) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const float send = upper ? p[h*16 + i] : p[h*16 + i + n];
                  const float keep = upper ? p[h*16 + i + n] : p[h*16 + i];
                  p[h*16 + i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
              }
            }
            if (lane < 16) {
              put_relay_operand<NPASS>(rb, quarter, sub * 32 + lane, p[0]);
              put_relay_operand<NPASS>(rb, quarter, sub * 32 + 16 + lane, p[16]);
            }
          }
#else
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            // (both heads in one 32-column pass with interleaved reductions was tried: more spills, 5 % slower)
            float k[16];
            tmem_ld16(lane_addr + ACC1 + sub * 32 + h * 16, k);
            tmem_ld_wait();
            float d1 = 0.f, d2 = 0.f;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const float4 qq = reinterpret_cast<const float4*>(my_q + h * 16)[q4];
              d1 = fmaf(qq.x, k[4*q4], d1); d1 = fmaf(qq.y, k[4*q4+1], d1);
              d1 = fmaf(qq.z, k[4*q4+2], d1); d1 = fmaf(qq.w, k[4*q4+3], d1);
              const float4 kk = k2[h * 4 + q4];
              d2 = fmaf(qq.x, kk.x, d2); d2 = fmaf(qq.y, kk.y, d2); d2 = fmaf(qq.z, kk.z, d2); d2 = fmaf(qq.w, kk.w, d2);
            }
            d1 *= 0.25f;
            d2 = has2 ? d2 * 0.25f : -3.4e38f;
            const float mx = warp_max(fmaxf(d1, d2));
            const float e1 = expf(d1 - mx), e2 = has2 ? expf(d2 - mx) : 0.f;
            const float inv = 1.0f / warp_sum(e1 + e2);
            w1[h] = e1 * inv;
            w2[h] = e2 * inv;
          }
          DSC_TR(warp ? 512 : 256);
#pragma unroll
          for (int i = 0; i < 8; ++i) k2[i] = (!SF_PRED_KV2 || has2) ? __ldg(kv2 + (32 + i) * 32) : make_float4(0.f, 0.f, 0.f, 0.f);      // the values take the keys' registers
          free_acc(1);
          wait_acc(0);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float v[16], p[16];
            tmem_ld16(lane_addr + ACC0 + sub * 32 + h * 16, v);
            tmem_ld_wait();
            if (h == 1) free_acc(0);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              float4 vv = k2[h * 4 + q4];
              if (!has2) vv = make_float4(0.f, 0.f, 0.f, 0.f);                // masked rows may hold anything
              p[4*q4]     = fmaf(w1[h], v[4*q4],     w2[h] * vv.x);
              p[4*q4 + 1] = fmaf(w1[h], v[4*q4 + 1], w2[h] * vv.y);
              p[4*q4 + 2] = fmaf(w1[h], v[4*q4 + 2], w2[h] * vv.z);
              p[4*q4 + 3] = fmaf(w1[h], v[4*q4 + 3], w2[h] * vv.w);
            }
            // sum over the 32 key lanes: fold the half-warps, then reduce-scatter 16 values over 16 lanes;
            // lane l (and l ^ 16) ends with dimension l & 15 of this head
#pragma unroll
            for (int i = 0; i < 16; ++i) p[i] += __shfl_xor_sync(0xffffffffu, p[i], 16);
#pragma unroll
            for (int off = 8, n = 8; off >= 1; off >>= 1, n >>= 1) {
              const bool upper = (lane & off) != 0;
#pragma unroll
              for (int i = 0; i < n; ++i) {
                const float send = upper ? p[i] : p[i + n];
                const float keep = upper ? p[i + n] : p[i];
                p[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
              }
            }
            // att_r[sentence quarter][32*sub + 16*h + lane] straight into the operand of J7 (its last reader, J8 of the
            // previous cycle, has completed: every warp drained that accumulator)
            if (lane < 16) put_relay_operand<NPASS>(rb, quarter, sub * 32 + h * 16 + lane, p[0]);
          }
#endif
          fence_async_smem();
          warp_arrive(&bars.t_ready, 1);
        }

        // ================= J7 (ACC0, transposed): s'[sentence sub][feature 32*quarter + lane] = relu(att_r @ Wo_relay + b)
        {
          const int f = quarter * 32 + lane;
          wait_acc(0);
          float v = tmem_ld1(lane_addr + ACC0 + sub);
          tmem_ld_wait();
          free_acc(0);
          v = fmaxf(v + bias_s[1][f], 0.f);
          s_cur[sub][f] = v;
          const bool next_j8 = !last && seq.has_j8(c + 1);
          if (!last) {
            // bf16 hi / lo of s'[sub][f], packed in pairs by the even lanes: operand of J8 (J7 has completed) and the words
            // that warp (quarter = sub, column quarter = quarter) patches into the relay row of X
            const __nv_bfloat16 hb = __float2bfloat16_rn(v);
            const __nv_bfloat16 lb = __float2bfloat16_rn(v - __bfloat162float(hb));
            uint32_t hw = (uint32_t)__bfloat16_as_ushort(hb), lw = (uint32_t)__bfloat16_as_ushort(lb);
            hw |= __shfl_down_sync(0xffffffffu, hw, 1) << 16;
            lw |= __shfl_down_sync(0xffffffffu, lw, 1) << 16;
            if ((lane & 1) == 0) {
              const uint32_t off = ((uint32_t)f >> 6) * RB_PLANE + sw128_offset((uint32_t)sub, (uint32_t)f & 63u);
              if (next_j8) *reinterpret_cast<uint32_t*>(rb + off) = hw;
              patch_w[0][sub][quarter][lane >> 1] = hw;
              if (NPASS == 3) {
                if (next_j8) *reinterpret_cast<uint32_t*>(rb + 2 * RB_PLANE + off) = lw;
                patch_w[1][sub][quarter][lane >> 1] = lw;
              }
            }
            if (next_j8) {
              fence_async_smem();
              warp_arrive(&bars.t_ready, 1);
            }
          }
          compute_warps_sync();                                              // s' of the four sentences is complete
          DSC_TR(warp ? 512 : 256);
          if (last) {
            if (lane == 31) {
              float4* xr = reinterpret_cast<float4*>(Xrow + ((int64_t)t * 128 + row_in_tile) * 128 + sub * 32);
#pragma unroll
              for (int q4 = 0; q4 < 8; ++q4) xr[q4] = reinterpret_cast<const float4*>(my_s)[q4];
            }
          } else {
            // patch the relay row of the X operand with s' (J5/J6, the last readers of X', have completed): every lane
            // rewrites its own row unchanged, the relay lane substitutes the new words
#pragma unroll
            for (int part = 0; part < ((NPASS == 3) ? 2 : 1); ++part) {
              uint32_t w[16];
              const uint32_t col = lane_addr + (part ? AX_LO : AX_HI) + sub * 16;
              tmem_ld16(col, reinterpret_cast<float*>(w));
              tmem_ld_wait();
              if (lane == 31) {
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                  const uint4 pw = reinterpret_cast<const uint4*>(patch_w[part][quarter][sub])[i4];
                  w[4*i4] = pw.x; w[4*i4+1] = pw.y; w[4*i4+2] = pw.z; w[4*i4+3] = pw.w;
                }
              }
              tmem_st16(col, w);
            }
            DSC_TR(warp ? 512 : 256);
            tmem_st_wait();
            warp_arrive(&bars.x_ready, 1);
            load_kve(gp);                                                // e-keys of the next cycle's first head pair
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == sf::kMmaWarp) tmem_dealloc<512>(tmem_base);
}

}  // namespace dsc

using namespace dsc;

#ifdef DSC_DEBUG_TOOLS
#include "debug/deepsc_b200_debug.h"
static bool g_star_trace_on = false;
namespace dsc { extern unsigned long long* g_pp_trace_host; }
#endif


template <int NPASS, bool TRACE>
static int launch_star_fused(const float* xi0, const float* s0, const float* q0, const float* kvei, const float* kv2i, int n2,
                             const sf::Weights& w, const float* bias_o, const float* bias_r, float* xrow, int n_tiles,
                             int n_cycles, int skip0, cudaStream_t s) {
  constexpr size_t smem = (size_t)sf::STAGES * sf::STAGE_BYTES + sf::RB_BYTES + 1024;
  cudaError_t e = cudaFuncSetAttribute(star_fused_kernel<NPASS, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("dsc_star_cycles_tc: %s", cudaGetErrorString(e)); return DSC_ERR_CUDA; }
  const int grid = n_tiles < kSMs ? n_tiles : kSMs;
  star_fused_kernel<NPASS, TRACE><<<grid, sf::kThreads, smem, s>>>(xi0, s0, q0, kvei, kv2i, n2, w, bias_o, bias_r, xrow, n_tiles, n_cycles, skip0);
  return check_launch("dsc_star_cycles_tc");
}

#ifdef DSC_DEBUG_TOOLS
extern "C" int dsc_debug_star_trace(void* device_buffer_768_u64) {
  unsigned long long* p = reinterpret_cast<unsigned long long*>(device_buffer_768_u64);
  cudaError_t e = cudaMemcpyToSymbol(dsc::g_star_trace, &p, sizeof(p));
  if (e != cudaSuccess) { set_error("dsc_debug_star_trace: %s", cudaGetErrorString(e)); return DSC_ERR_CUDA; }
  g_star_trace_on = p != nullptr;
  dsc::g_pp_trace_host = p;
  return DSC_OK;
}
#endif

extern "C" int dsc_star_cycles_tc(const float* x_tile0, const float* s0, const float* q0, const float* kv_e,
                                  const float* kv2, int n2,
                                  const void* packed_wqkv_grouped, const void* packed_wo, const void* packed_wkv_relay,
                                  const void* packed_wo_relay, const void* packed_wq_relay,
                                  const float* bias_o, const float* bias_o_relay,
                                  float* x_rowmajor, int n_sent, int n_cycles, int prec, void* stream) {
  DSC_REQUIRE(x_tile0 && s0 && q0 && kv_e && packed_wqkv_grouped && packed_wo && packed_wkv_relay && packed_wo_relay &&
              packed_wq_relay && bias_o && bias_o_relay && x_rowmajor, "dsc_star_cycles_tc: null pointer");
  DSC_REQUIRE(n_sent >= 0 && (n_sent % 4) == 0 && n_cycles >= 1, "dsc_star_cycles_tc: n_sent must be a multiple of 4, n_cycles >= 1");
  DSC_REQUIRE(n2 >= 0 && n2 <= 32 && (n2 == 0 || kv2), "dsc_star_cycles_tc: bad h2 key count");
  DSC_REQUIRE(aligned16(x_tile0) && aligned16(s0) && aligned16(q0) && aligned16(kv_e) && (!kv2 || aligned16(kv2)) &&
              aligned16(bias_o) && aligned16(bias_o_relay) && aligned16(x_rowmajor), "dsc_star_cycles_tc: misaligned pointer");
  DSC_REQUIRE((((uintptr_t)packed_wqkv_grouped | (uintptr_t)packed_wo | (uintptr_t)packed_wkv_relay | (uintptr_t)packed_wo_relay |
                (uintptr_t)packed_wq_relay) & 127u) == 0, "dsc_star_cycles_tc: packed weights must be 128-byte aligned");
  const int skip0 = ((prec & DSC_STAR_FIRST_SAT_DONE) ? 1 : 0) | ((prec & DSC_STAR_NO_FINAL_RELAY) ? 2 : 0);   // kernel flags
  const int form = prec & (DSC_STAR_FORM_ONE_TILE | DSC_STAR_FORM_TWO_TILE);
  prec &= 255;
  DSC_REQUIRE(prec == 1 || prec == 2, "dsc_star_cycles_tc: prec must be 1 (bf16x3) or 2 (bf16)");
  DSC_REQUIRE(form != (DSC_STAR_FORM_ONE_TILE | DSC_STAR_FORM_TWO_TILE), "dsc_star_cycles_tc: pick at most one kernel form");
  DSC_REQUIRE(!(skip0 & 1) || n_cycles >= 2, "dsc_star_cycles_tc: DSC_STAR_FIRST_SAT_DONE needs n_cycles >= 2");
  if (n_sent == 0) return DSC_OK;
  if (n2 == 0) kv2 = kv_e;        // rows are read but masked (lane < n2 is false): any readable [n_sent][64][32][4] floats do
  sf::Weights w{reinterpret_cast<const uint8_t*>(packed_wqkv_grouped), reinterpret_cast<const uint8_t*>(packed_wo),
                reinterpret_cast<const uint8_t*>(packed_wkv_relay), reinterpret_cast<const uint8_t*>(packed_wo_relay),
                reinterpret_cast<const uint8_t*>(packed_wq_relay)};
  cudaStream_t st = as_stream(stream);
  const int n_tiles = n_sent / 4;
  // The product library has ONE form: the one-tile kernel of this file.  The two-tile (half-cycle-apart) kernel of
  // debug/dsc_star_pp.cu - bit-identical results, measured 0-2 % faster at 4 tiles per SM before this kernel's register
  // reallocation and slower since - is an experiment kept in the debug-tools library only.
#ifdef DSC_DEBUG_TOOLS
  if (form == DSC_STAR_FORM_TWO_TILE)
    return launch_star_pp(x_tile0, s0, q0, kv_e, kv2, n2, w, bias_o, bias_o_relay, x_rowmajor, n_tiles, n_cycles, skip0, prec == 1 ? 3 : 1, st);
#else
  DSC_REQUIRE(form != DSC_STAR_FORM_TWO_TILE, "dsc_star_cycles_tc: the two-tile form lives in libdeepsc_b200_debug.so only");
#endif
#ifdef DSC_DEBUG_TOOLS
  if (g_star_trace_on)
    return prec == 1 ? launch_star_fused<3, true>(x_tile0, s0, q0, kv_e, kv2, n2, w, bias_o, bias_o_relay, x_rowmajor, n_tiles, n_cycles, skip0, st)
                     : launch_star_fused<1, true>(x_tile0, s0, q0, kv_e, kv2, n2, w, bias_o, bias_o_relay, x_rowmajor, n_tiles, n_cycles, skip0, st);
#endif
  return prec == 1 ? launch_star_fused<3, false>(x_tile0, s0, q0, kv_e, kv2, n2, w, bias_o, bias_o_relay, x_rowmajor, n_tiles, n_cycles, skip0, st)
                   : launch_star_fused<1, false>(x_tile0, s0, q0, kv_e, kv2, n2, w, bias_o, bias_o_relay, x_rowmajor, n_tiles, n_cycles, skip0, st);
}
