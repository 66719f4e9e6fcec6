// Two-tile form of the fused star layer (the throughput form behind dsc_star_cycles_tc; dsc_star_fused.cu is the
// one-tile latency form).  Same jobs, same arithmetic and the same UMMA issue order INSIDE every job as star_fused_kernel
// (models/modules.py:283-306, 359-378), so both forms give bit-identical results.
//
// Why: a tile's cycle is a serial chain  UMMA -> register epilogue -> UMMA -> ...; with one tile per CTA the compute
// warps wait on the tensor pipe a third of the time (ncu: 33 % of all warp samples sit in mbarrier polls) and the pipe
// waits on the warps the rest (38 % tensor-active).  Here a CTA works on TWO tiles (slots A, B) that are half a cycle
// apart: while slot A is in its satellite half S (J0..J4) slot B is in its relay half R (J5..J7 [+J8]), then they swap.
// The same 16 compute warps run the halves as six chunks each and alternate between the slots chunk by chunk
//     A.c0  B.c0  A.c1  B.c1  A.c2  B.c2          S: c0 = heads round 0, c1 = heads round 1, c2 = Wo epilogue
//                                                  R: c0 = relay logits,  c1 = weighted values, c2 = relay update
// so the UMMA that feeds a slot's next chunk runs while the warps are busy with the other slot's chunk.
//
// TMEM (512 columns): X operand of slot A [0,128) and of slot B [128,256) (bf16 hi | lo), BIG accumulator [256,384)
// (J4, J5, J6 of either slot, strictly alternating), QKV accumulator [384,480) (one 96-column head pair at a time: it is
// released right after the q|k|v load), J7 result [480,496), J8 result [496,512).  There is no room for a second X-sized
// operand, so the attention output ATT (J4's A operand) lives in shared memory (64 KB, K-major 128-byte swizzle, the
// layout of a weight chunk) and J4 is an SS-mode UMMA; one ATT buffer serves both slots because only one of them is in
// its satellite half at any time.  Shared memory: weight ring 3 x 32 KB + ATT 64 KB + relay-vector operand 8 KB.
//
// All three roles (weight producer, UMMA issuer, compute warps) walk the same static schedule (Plan / for_each_job):
// a tile is a list of phases [S(0)] R(0) S(1) ... S(n-1) [R(n-1)] (S(0) absent with DSC_STAR_FIRST_SAT_DONE, R(n-1) with
// DSC_STAR_NO_FINAL_RELAY), padded to an even length so that phase p of a slot's tile stream is of type (first + p) & 1;
// slot B runs one phase behind slot A, hence always in the opposite half.
#include "dsc_star_common.cuh"

namespace dsc {

using namespace tc;

namespace pp {
// 20 warps = 5 warpgroups: 16 compute warps, then issuer, producer and two idle warps.  The last warpgroup gives most of its
// registers back (setmaxnreg.dec) and the compute warpgroups take them (setmaxnreg.inc): 96 registers per thread at launch
// (65,536 / 640), PP_COMPUTE_REGS in the chunk code, which keeps most of its state out of local memory - whose round trips to
// L2 (the L1 left beside 193 KB of shared memory is thrashed by the key streams) were ~0.5 us per chunk hand-over.  setmaxnreg.inc
// can only take what setmaxnreg.dec released into the CTA's pool: (C - 96) x 512 <= (96 - S) x 128, hence 104 / 64.
#ifndef PP_COMPUTE_REGS
#define PP_COMPUTE_REGS 104
#endif
#ifndef PP_SIDE_REGS
#define PP_SIDE_REGS 64
#endif
constexpr int kCompute = 16, kMmaWarp = 16, kProdWarp = 17, kThreads = 640;
constexpr uint32_t BIG = 256, QKV = 384, G7 = 480, G8 = 496;
// Weight ring: a job's chunk is streamed as two half-chunks - the bf16 hi planes (K-blocks 0, 1), used by the passes
// hi*hi and lo*hi, then the lo planes for the pass hi*lo - so a stage is 32 KB instead of 64 KB and the issue order inside a
// job is unchanged.  Shared memory decides the L1 that is left for the spill slots (576 threads x ~70 B of stack) and the key
// streams: 3 x 32 KB ring + 64 KB ATT + 16 KB + 11 KB static = 193 KB -> 196 KB carve-out, 60 KB of L1; with the 2 x 64 KB
// ring (217 KB -> 228 KB carve-out, 28 KB of L1) every chunk hand-over cost ~0.6 us of local-memory round trips to L2.
#ifndef PP_RING
#define PP_RING 3
#endif
// key loads go through L1 on purpose: a tile re-reads its keys every cycle, and __ldcs / L1::no_allocate loads were 5 % slower
#define PP_LD(p) __ldg(p)
constexpr int RING = PP_RING;
using sf::RSTAGE;
constexpr uint32_t ATT_PLANE = 128 * 128;                     // one (part, K-block) plane: 128 rows x 128 B
constexpr uint32_t ATT_BYTES = 4 * ATT_PLANE;
__device__ __forceinline__ uint32_t ax_hi(int T) { return (uint32_t)T * 128u; }
__device__ __forceinline__ uint32_t ax_lo(int T) { return (uint32_t)T * 128u + 64u; }

struct Ph { bool ok, first; int ti, c, kind; };                // phase of a slot: tile of its stream, cycle, 0 = S / 1 = R

struct Plan {
  int n_cycles, skip0, nfr, L, Lp;
  __device__ Plan(int n, bool s, bool f) : n_cycles(n), skip0(s ? 1 : 0), nfr(f ? 1 : 0) {
    L = 2 * n - skip0 - nfr;
    Lp = (L + 1) & ~1;
  }
  __device__ __forceinline__ int first_type() const { return skip0; }
  __device__ __forceinline__ Ph decode(int p, int n_tiles) const {
    Ph r{false, false, 0, 0, 0};
    if (p < 0) return r;
    r.ti = p / Lp;
    const int q = p - r.ti * Lp;
    if (r.ti >= n_tiles || q >= L) return r;
    const int idx = q + skip0;
    r.ok = true; r.first = (q == 0); r.c = idx >> 1; r.kind = idx & 1;
    return r;
  }
  __device__ __forceinline__ bool has_relay(int c) const { return !(nfr && c + 1 == n_cycles); }
  __device__ __forceinline__ bool j8_after(int c) const { return c + 1 < n_cycles && has_relay(c + 1); }   // R(c) feeds R(c+1)
  __device__ __forceinline__ bool has_j8(int c) const { return c > 0 && has_relay(c); }                    // S(c) reads q'
};

// The global UMMA job order: f(slot, job, cycle).  Jobs: 0..3 QKV head pairs, 4 Wo_sat, 5 / 6 K / V relay, 7 Wo_relay GEMV,
// 8 Wq_relay GEMV (the query of the NEXT cycle's relay attention).  The first job of a phase (J0 / J5) is issued at the
// tail of the half-step before, so that it runs under the other slot's last chunk; J8 goes last (its result is read two
// chunks into the slot's next half).
// Nine list entries per half-step, 7 bits each: job | slot << 4 | condition << 5 (0: the slot has a phase in this half-step,
// 1: it has one in the NEXT half-step - the head job J0 / J5 of that phase, 2: this R phase feeds another one -> J8).
__host__ __device__ constexpr unsigned long long pp_entry(int k, int slot, int job, int cond) {
  return (unsigned long long)(job | slot << 4 | cond << 5) << (7 * k);
}
// A in S, B in R (chunks A.S1a B.R1 A.S1b B.R2 A.S2 B.R3):  A1 A2 B6 A3 A4 B7 | A5' B8 B0'
constexpr unsigned long long kListAS = pp_entry(0, 0, 1, 0) | pp_entry(1, 0, 2, 0) | pp_entry(2, 1, 6, 0) | pp_entry(3, 0, 3, 0) |
                                       pp_entry(4, 0, 4, 0) | pp_entry(5, 1, 7, 0) | pp_entry(6, 0, 5, 1) | pp_entry(7, 1, 8, 2) |
                                       pp_entry(8, 1, 0, 1);
// A in R, B in S (chunks A.R1 B.S1a A.R2 B.S1b A.R3 B.S2):  A6 B1 B2 B3 A7 B4 | A8 A0' B5'
constexpr unsigned long long kListAR = pp_entry(0, 0, 6, 0) | pp_entry(1, 1, 1, 0) | pp_entry(2, 1, 2, 0) | pp_entry(3, 1, 3, 0) |
                                       pp_entry(4, 0, 7, 0) | pp_entry(5, 1, 4, 0) | pp_entry(6, 0, 8, 2) | pp_entry(7, 0, 0, 1) |
                                       pp_entry(8, 1, 5, 1);

template <class F>
__device__ __forceinline__ void for_each_job(const Plan& P, int nA, int nB, int H, F&& f) {
  if (nA > 0) f(0, P.skip0 ? 5 : 0);                            // head job of slot A's first phase
  int qA = 0, tiA = 0, qB = -1, tiB = 0;                        // as in the compute warps: phase inside the tile, tile of the stream
#pragma unroll 1
  for (int h = 0; h < H; ++h) {
    int qA1 = qA + 1, tiA1 = tiA, qB1 = qB + 1, tiB1 = tiB;
    if (qA1 == P.Lp) { qA1 = 0; ++tiA1; }
    if (qB1 == P.Lp) { qB1 = 0; ++tiB1; }
    // bit T: slot T has a phase now / next half-step / this phase is an R phase followed by another R phase of the tile
    const int cA = (qA + P.skip0) >> 1, cB = (qB + P.skip0) >> 1;
    const bool eA = qA < P.L && tiA < nA, eB = qB >= 0 && qB < P.L && tiB < nB;
    const uint32_t now = (eA ? 1u : 0u) | (eB ? 2u : 0u);
    const uint32_t nxt = ((qA1 < P.L && tiA1 < nA) ? 1u : 0u) | ((qB1 < P.L && tiB1 < nB) ? 2u : 0u);
    const uint32_t j8 = ((eA && P.j8_after(cA)) ? 1u : 0u) | ((eB && P.j8_after(cB)) ? 2u : 0u);
    unsigned long long list = (((P.skip0 + h) & 1) == 0) ? kListAS : kListAR;
#pragma unroll 1
    for (int k = 0; k < 9; ++k, list >>= 7) {
      const uint32_t e = (uint32_t)list & 127u, T = (e >> 4) & 1u, cond = e >> 5;
      const uint32_t mask = cond == 0 ? now : cond == 1 ? nxt : j8;
      if ((mask >> T) & 1u) f((int)T, (int)(e & 15u));
    }
    qA = qA1; tiA = tiA1; qB = qB1; tiB = tiB1;
  }
}

struct Bars {
  uint64_t w_full[RING], w_free[RING];
  uint64_t qkv_full[2], qkv_free;        // even / odd head pairs complete; accumulator drained (8 warps)
  uint64_t big_full, big_free;
  uint64_t g7_full, g8_full;
  uint64_t x_ready[2];                   // X operand of a slot staged (tile start, X' after J4, relay-row patch)
  uint64_t ta_ready, tb_ready;           // ATT heads 0..3 / 4..7 staged (K-block 0 / 1 of J4's operand)
  uint64_t rb_ready;                     // relay-vector operand staged (att_r for J7, s' for J8)
};

// one K-block of J4 in SS mode (A = ATT planes in shared memory), all passes: hi*hi, lo*hi, hi*lo as in issue_group_kb
template <int NPASS>
__device__ __forceinline__ void issue_j4_kb(uint32_t d_tmem, uint32_t att_base, uint32_t b_hi, uint32_t b_lo, int kb, bool first) {
  constexpr uint32_t IDESC = idesc_bf16_f32(128, 128);
#pragma unroll
  for (int pass = 0; pass < NPASS; ++pass) {
    const uint32_t pa = (pass == 1) ? 1u : 0u;
    const uint32_t bb = (pass == 2) ? b_lo : b_hi;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      umma_ss(d_tmem, smem_desc_sw128(att_base + (pa * 2 + kb) * ATT_PLANE + ks * 32u),
              smem_desc_sw128(bb + kb * (128u * 128u) + ks * 32u), IDESC, (first && pass == 0 && ks == 0) ? 0u : 1u);
  }
}
}  // namespace pp

// Debug timeline (tools/pp_trace.py, debug-tools library only): compute warps 0 (head pairs 0, 2) and 8 (head pairs 1, 3) of
// CTA 0 stamp clock64() at the start of every chunk, after its accumulator wait and at its end: [warp][chunk][3].
#ifdef DSC_DEBUG_TOOLS
unsigned long long* g_pp_trace_host = nullptr;
#define PP_TR() do { if (tr_buf && tr_n < 384) tr_buf[tr_n] = (unsigned long long)clock64(); ++tr_n; } while (0)
#else
#define PP_TR() do { } while (0)
#endif

template <int NPASS>
__global__ void __launch_bounds__(pp::kThreads, 1)
star_pp_kernel(const float* __restrict__ XI0, const float* __restrict__ S0, const float* __restrict__ Q0,
               const float* __restrict__ KVEI, const float* __restrict__ KV2I, int n2, sf::Weights W,
               const float* __restrict__ bias_o, const float* __restrict__ bias_r,
               float* __restrict__ Xrow, int n_tiles, int n_cycles, int flags, unsigned long long* trace) {
  using namespace pp;
  using sf::RB_PLANE; using sf::RB_BYTES;
  const Plan plan(n_cycles, (flags & 1) != 0, (flags & 2) != 0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* att = ring + RING * RSTAGE;         // ATT: [hi|lo][K-block] planes of 128 rows x 128 B, 128-byte swizzle
  uint8_t* rb0 = att + ATT_BYTES;                     // [slot] att_r, then s': B operand of the transposed relay GEMVs (rows 4..15 zero)
  __shared__ __align__(8) pp::Bars bars;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_cur[2][4][128];    // [slot]: relay node of the 4 sentences of the tile
  __shared__ __align__(16) float q_cur[2][4][128];    // its query under the relay weights
  __shared__ __align__(16) float bias_s[2][128];
  __shared__ __align__(16) uint32_t patch_w[2][4][4][16];
  constexpr int parts = (NPASS == 3) ? 2 : 1;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (uint32_t i = tid; i < 2 * RB_BYTES / 16; i += kThreads) reinterpret_cast<uint4*>(rb0)[i] = make_uint4(0, 0, 0, 0);
  fence_async_smem();
  if (tid < 128) bias_s[0][tid] = __ldg(bias_o + tid);
  else if (tid < 256) bias_s[1][tid - 128] = __ldg(bias_r + tid - 128);
  if (tid == 0) {
    for (int s = 0; s < RING; ++s) { mbar_init(&bars.w_full[s], 1); mbar_init(&bars.w_free[s], 1); }
    mbar_init(&bars.qkv_full[0], 1); mbar_init(&bars.qkv_full[1], 1); mbar_init(&bars.qkv_free, kCompute / 2);
    mbar_init(&bars.big_full, 1); mbar_init(&bars.big_free, kCompute);
    mbar_init(&bars.g7_full, 1); mbar_init(&bars.g8_full, 1);
    mbar_init(&bars.x_ready[0], kCompute); mbar_init(&bars.x_ready[1], kCompute);
    mbar_init(&bars.ta_ready, kCompute); mbar_init(&bars.tb_ready, kCompute);
    mbar_init(&bars.rb_ready, kCompute);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int nA = (my_tiles + 1) >> 1, nB = my_tiles >> 1;
  const int hA = nA * plan.Lp, hB = nB ? nB * plan.Lp + 1 : 0;
  const int H = hA > hB ? hA : hB;

  if (warp >= kCompute) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(PP_SIDE_REGS));      // the whole last warpgroup, one instruction
  if (warp == kProdWarp) {
    // ------------------------------------------------------------------ weight producer
    if (lane == 0) {
      uint32_t n = 0;
      for_each_job(plan, nA, nB, H, [&](int, int j) {
        const uint8_t* blob; uint32_t rows, row0, n_pad;
        sf::chunk_of(W, j, blob, rows, row0, n_pad);
        const uint32_t plane = rows * 128u;
        for (int half = 0; half < parts; ++half, ++n) {            // hi planes, then lo planes
          const uint32_t st = n % RING;
          mbar_wait(&bars.w_free[st], ((n / RING) - 1) & 1);
          mbar_expect_tx(&bars.w_full[st], 2 * plane);
          for (int kb = 0; kb < 2; ++kb)
            bulk_g2s(ring + st * RSTAGE + kb * plane, blob + ((size_t)(half * 2 + kb) * n_pad + row0) * 128, plane, &bars.w_full[st]);
        }
      });
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ UMMA issuer (whole warp in lock-step, one elected lane issues)
    const bool leader = elect_one();
    const uint32_t ring_base = smem_u32(ring), rb_base0 = smem_u32(rb0), att_base = smem_u32(att);
    uint32_t n = 0, qj = 0, bj = 0, ta = 0, tb = 0, rbn = 0, xr0 = 0, xr1 = 0;
    for_each_job(plan, nA, nB, H, [&](int T, int j) {
      const uint32_t st_hi = n % RING, par_hi = (n / RING) & 1;
      ++n;
      uint32_t st_lo = 0, par_lo = 0;
      if (NPASS == 3) { st_lo = n % RING; par_lo = (n / RING) & 1; ++n; }
      if (j == 0 || j == 5) {                                       // X of the slot staged / X' restaged / relay row patched
        if (T) { mbar_wait(&bars.x_ready[1], xr1 & 1); ++xr1; } else { mbar_wait(&bars.x_ready[0], xr0 & 1); ++xr0; }
      }
      if (j == 7 || j == 8) { mbar_wait(&bars.rb_ready, rbn & 1); ++rbn; }
      if (j == 4) { mbar_wait(&bars.ta_ready, ta & 1); ++ta; }
      mbar_wait(&bars.w_full[st_hi], par_hi);
      if (j < 4) mbar_wait(&bars.qkv_free, (qj - 1) & 1);          // the previous head pair was loaded by its 8 warps
      else if (j < 7) mbar_wait(&bars.big_free, (bj - 1) & 1);     // the previous BIG result was loaded by all 16
      tc_fence_after();
      const uint32_t b_hi = ring_base + st_hi * RSTAGE, b_lo = ring_base + st_lo * RSTAGE;
      const uint32_t rb_base = rb_base0 + (uint32_t)T * RB_BYTES;
      uint64_t* done = j < 4 ? &bars.qkv_full[j & 1] : j < 7 ? &bars.big_full : j == 7 ? &bars.g7_full : &bars.g8_full;
      if (j == 4) {
        // K-block-major: heads 0..3 of ATT first, heads 4..7 when they are staged; both half-chunks stay until the end
        if (NPASS == 3) { mbar_wait(&bars.w_full[st_lo], par_lo); tc_fence_after(); }
        if (leader) issue_j4_kb<NPASS>(tmem_base + BIG, att_base, b_hi, b_lo, 0, true);
        mbar_wait(&bars.tb_ready, tb & 1); ++tb;
        tc_fence_after();
        if (leader) {
          issue_j4_kb<NPASS>(tmem_base + BIG, att_base, b_hi, b_lo, 1, false);
          umma_commit(&bars.w_free[st_hi]);
          if (NPASS == 3) umma_commit(&bars.w_free[st_lo]);
          umma_commit(done);
        }
      } else {
        // passes hi*hi and lo*hi on the hi planes, which are released at once; then hi*lo on the lo planes
        if (leader) {
          if (j < 4) {
            ts_pass<96>(tmem_base + QKV, tmem_base + ax_hi(T), b_hi, 96u * 128u, true);
            if (NPASS == 3) ts_pass<96>(tmem_base + QKV, tmem_base + ax_lo(T), b_hi, 96u * 128u, false);
          } else if (j < 7) {
            ts_pass<128>(tmem_base + BIG, tmem_base + ax_hi(T), b_hi, 128u * 128u, true);
            if (NPASS == 3) ts_pass<128>(tmem_base + BIG, tmem_base + ax_lo(T), b_hi, 128u * 128u, false);
          } else {
            const uint32_t d = tmem_base + (j == 7 ? G7 : G8);
            gemv_pass(d, b_hi, rb_base, true);                                   // w_hi * v_hi
            if (NPASS == 3) gemv_pass(d, b_hi, rb_base + 2 * RB_PLANE, false);    // w_hi * v_lo
          }
          umma_commit(&bars.w_free[st_hi]);
        }
        if (NPASS == 3) {
          mbar_wait(&bars.w_full[st_lo], par_lo);
          tc_fence_after();
          if (leader) {
            if (j < 4) ts_pass<96>(tmem_base + QKV, tmem_base + ax_hi(T), b_lo, 96u * 128u, false);
            else if (j < 7) ts_pass<128>(tmem_base + BIG, tmem_base + ax_hi(T), b_lo, 128u * 128u, false);
            else gemv_pass(tmem_base + (j == 7 ? G7 : G8), b_lo, rb_base, false);  // w_lo * v_hi
            umma_commit(&bars.w_free[st_lo]);
          }
        }
        if (leader) umma_commit(done);
      }
      if (j < 4) ++qj; else if (j < 7) ++bj;
      __syncwarp();
    });
  }
  } else {
    // ------------------------------------------------------------------ compute warps (16)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(PP_COMPUTE_REGS));
    const int quarter = warp & 3, sub = warp >> 2;
    const int gp = sub >> 1, hh = sub & 1;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int row_in_tile = quarter * 32 + lane;
    const int up = (lane >= 30) ? 0 : lane + 1;
    const int dn = (lane == 0) ? 30 : lane - 1;
    uint32_t qf_use = 0, big_use = 0, g7_use = 0, g8_use = 0;       // barrier phases consumed so far
#ifdef DSC_DEBUG_TOOLS
    unsigned long long* tr_buf = (trace && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 8)) ? trace + (warp ? 384 : 0) : nullptr;
    int tr_n = 0;
#endif
    float w1[2], w2[2];                                            // relay softmax weights, from chunk R1 to chunk R2 of the same slot
    w1[0] = w1[1] = w2[0] = w2[1] = 0.f;

    auto warp_arrive = [&](uint64_t* bar) {                        // every lane fences its TMEM accesses, one lane arrives
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    auto wait_big = [&]() { mbar_wait(&bars.big_full, big_use & 1); ++big_use; tc_fence_after(); PP_TR(); };

    // per-slot position in its tile stream: phase index q inside the tile (slot B starts one half-step late) and tile ti;
    // everything else (cycle, half, first phase of the tile) is derived - few live registers between chunks
    int qA = 0, tiA = 0, qB = -1, tiB = 0;
    const int sk = plan.skip0, Lr = plan.L, Lpad = plan.Lp;
    for (int h = 0; h < H; ++h) {
      // ---- tile start of a slot (before any chunk of the half-step: the slot's first UMMA job sits in the issue order ahead
      //      of this half-step's other jobs): relay vectors to shared memory, X rows (X' with DSC_STAR_FIRST_SAT_DONE) to TMEM
#pragma unroll 1
      for (int T = 0; T < 2; ++T) {
        const int q = T ? qB : qA, ti = T ? tiB : tiA;
        if (q != 0 || ti >= (T ? nB : nA)) continue;
        const int t = (int)blockIdx.x + (2 * ti + T) * (int)gridDim.x;
        const int64_t sent = (int64_t)t * 4 + quarter;
        float* my_s = &s_cur[T][quarter][sub * 32];
        float* my_q = &q_cur[T][quarter][sub * 32];
        const float4* s0 = reinterpret_cast<const float4*>(S0 + sent * 128 + sub * 32);
        const float4* q0 = reinterpret_cast<const float4*>(Q0 + sent * 128 + sub * 32);
        if (lane < 8) reinterpret_cast<float4*>(my_s)[lane] = __ldg(s0 + lane);
        else if (lane < 16) reinterpret_cast<float4*>(my_q)[lane - 8] = __ldg(q0 + (lane - 8));
        __syncwarp();
        uint32_t hi[16], lo[16];
        if (lane == 31) split_quarter_row(my_s, hi, lo);
        else load_quarter_row(reinterpret_cast<const float4*>(XI0 + (int64_t)t * 16384) + (sub * 8) * 128 + row_in_tile, 128, hi, lo);
        store_quarter_row<NPASS>(lane_addr, ax_hi(T), ax_lo(T), sub, hi, lo);
        tmem_st_wait();
        warp_arrive(&bars.x_ready[T]);
      }
      for (int jc = 0; jc < 3; ++jc) {
#pragma unroll 1
        for (int T = 0; T < 2; ++T) {
#ifdef DSC_DEBUG_TOOLS
          const unsigned long long t_top = clock64();
#endif
          const int q = T ? qB : qA, ti = T ? tiB : tiA;
          if (q < 0 || q >= Lr || ti >= (T ? nB : nA)) continue;
          const int c = (q + sk) >> 1, kind = (q + sk) & 1;
          const bool last = (c + 1 == n_cycles);
          const int t = (int)blockIdx.x + (2 * ti + T) * (int)gridDim.x;
          const int64_t sent = (int64_t)t * 4 + quarter;
          float* my_s = &s_cur[T][quarter][sub * 32];
          float* my_q = &q_cur[T][quarter][sub * 32];
          const uint32_t axh = ax_hi(T), axl = ax_lo(T);
          uint8_t* rb = rb0 + (uint32_t)T * RB_BYTES;
#ifdef DSC_DEBUG_TOOLS
          if (tr_buf && tr_n < 384) tr_buf[tr_n] = t_top;          // the chunk's first stamp is taken at the loop top
          ++tr_n;
#endif

          if (kind == 0 && jc < 2) {
            // ================= S1a / S1b: head pair g = gp + 2*jc (QKV accumulator), head 2g + hh of this warp
            const int g = gp + 2 * jc, head = 2 * g + hh;
            // e-keys k[16] | v[16] of (row, head): issued first, consumed last (the L2 latency hides under the q|k|v load
            // and the neighbour logits)
            const uint4* kve_base = reinterpret_cast<const uint4*>(KVEI + (int64_t)t * 32768) + row_in_tile;
            uint32_t kv[32];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 a = PP_LD(kve_base + (head * 4 + q) * 128), cc = PP_LD(kve_base + (32 + head * 4 + q) * 128);
              kv[4*q] = a.x; kv[4*q+1] = a.y; kv[4*q+2] = a.z; kv[4*q+3] = a.w;
              kv[16+4*q] = cc.x; kv[16+4*q+1] = cc.y; kv[16+4*q+2] = cc.z; kv[16+4*q+3] = cc.w;
            }
            if (jc == 1 && plan.has_j8(c)) {
              // J8 (transposed): q'[sentence sub][feature 32*quarter + lane] = s' @ Wq_relay, issued behind the slot's J0
              mbar_wait(&bars.g8_full, g8_use & 1); ++g8_use;
              tc_fence_after();
              const float qv = tmem_ld1(lane_addr + G8 + sub);
              tmem_ld_wait();
              q_cur[T][sub][quarter * 32 + lane] = qv;                // read after the barrier at the head of chunk R1
            }
            mbar_wait(&bars.qkv_full[gp], qf_use & 1); ++qf_use;
            tc_fence_after();
            PP_TR();
            const uint32_t col = lane_addr + QKV + hh * 16;
            float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f, l4 = 0.f;
            float v[16];
            {
              float q[16], k[16];
              tmem_ld16(col, q);
              tmem_ld16(col + 32, k);
              tmem_ld16(col + 64, v);
              tmem_ld_wait();
              warp_arrive(&bars.qkv_free);                          // 8 warps drain a head pair
#pragma unroll
              for (int d = 0; d < 16; ++d) {
                const float ku = __shfl_sync(0xffffffffu, k[d], up);
                const float kd = __shfl_sync(0xffffffffu, k[d], dn);
                const float ks = __shfl_sync(0xffffffffu, k[d], 31);
                l0 = fmaf(q[d], ku, l0);
                l1 = fmaf(q[d], k[d], l1);
                l2 = fmaf(q[d], kd, l2);
                l4 = fmaf(q[d], ks, l4);
              }
#pragma unroll
              for (int d = 0; d < 16; ++d) l3 = fmaf(q[d], __uint_as_float(kv[d]), l3);
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
            // head `head` covers k = 16*head .. 16*head+15: K-block head >> 2, 16-byte units 2*(head & 3), +1 of the row
            {
              const uint32_t r = (uint32_t)row_in_tile, u0 = (uint32_t)(head & 3) * 2u;
              uint8_t* base = att + (uint32_t)(head >> 2) * ATT_PLANE + r * 128u;
              const uint32_t o0 = ((u0 ^ (r & 7u)) << 4), o1 = (((u0 + 1u) ^ (r & 7u)) << 4);
              *reinterpret_cast<uint4*>(base + o0) = make_uint4(ohi[0], ohi[1], ohi[2], ohi[3]);
              *reinterpret_cast<uint4*>(base + o1) = make_uint4(ohi[4], ohi[5], ohi[6], ohi[7]);
              if (NPASS == 3) {
                *reinterpret_cast<uint4*>(base + 2 * ATT_PLANE + o0) = make_uint4(olo[0], olo[1], olo[2], olo[3]);
                *reinterpret_cast<uint4*>(base + 2 * ATT_PLANE + o1) = make_uint4(olo[4], olo[5], olo[6], olo[7]);
              }
            }
            fence_async_smem();
            warp_arrive(jc == 0 ? &bars.ta_ready : &bars.tb_ready);
          } else if (kind == 0) {
            // ================= S2 (J4 in BIG): X' = relu(ATT @ Wo + b), columns 32*sub..; the relay row keeps s; re-staged as X
            wait_big();
            float v[32];
            tmem_ld32(lane_addr + BIG + sub * 32, v);
            tmem_ld_wait();
            warp_arrive(&bars.big_free);
            {
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
            if (plan.has_relay(c)) {
              uint32_t hi[16], lo[16];
              split_quarter_row(v, hi, lo);
              store_quarter_row<NPASS>(lane_addr, axh, axl, sub, hi, lo);     // J0..J3 of this slot have completed
              tmem_st_wait();
              warp_arrive(&bars.x_ready[T]);
            }
          } else if (jc == 0) {
            // ================= R1 (J5 = K in BIG): relay logits and softmax weights, heads 2*sub and 2*sub+1, lane = key row
            const float4* kv2 = reinterpret_cast<const float4*>(KV2I + sent * 8192) + (sub * 8) * 32 + lane;
            const bool has2 = lane < n2;
            float4 k2[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) k2[i] = has2 ? PP_LD(kv2 + i * 32) : make_float4(0.f, 0.f, 0.f, 0.f);   // rows >= n2 are masked: not fetched
            compute_warps_sync();                                      // q' (written feature-major by the J8 epilogue) is complete
            wait_big();
#pragma unroll
            for (int hd = 0; hd < 2; ++hd) {
              float k[16];
              tmem_ld16(lane_addr + BIG + sub * 32 + hd * 16, k);
              tmem_ld_wait();
              if (hd == 1) warp_arrive(&bars.big_free);
              float d1 = 0.f, d2 = 0.f;
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                const float4 qq = reinterpret_cast<const float4*>(my_q + hd * 16)[q4];
                d1 = fmaf(qq.x, k[4*q4], d1); d1 = fmaf(qq.y, k[4*q4+1], d1);
                d1 = fmaf(qq.z, k[4*q4+2], d1); d1 = fmaf(qq.w, k[4*q4+3], d1);
                const float4 k4 = k2[hd * 4 + q4];
                d2 = fmaf(qq.x, k4.x, d2); d2 = fmaf(qq.y, k4.y, d2); d2 = fmaf(qq.z, k4.z, d2); d2 = fmaf(qq.w, k4.w, d2);
              }
              d1 *= 0.25f;
              d2 = has2 ? d2 * 0.25f : -3.4e38f;
              const float mx = warp_max(fmaxf(d1, d2));
              const float e1 = expf(d1 - mx), e2 = has2 ? expf(d2 - mx) : 0.f;
              const float inv = 1.0f / warp_sum(e1 + e2);
              w1[hd] = e1 * inv;
              w2[hd] = e2 * inv;
            }
          } else if (jc == 1) {
            // ================= R2 (J6 = V in BIG): att_r = sum over the key lanes of w * v -> operand of J7
            const float4* kv2 = reinterpret_cast<const float4*>(KV2I + sent * 8192) + (sub * 8) * 32 + lane;
            const bool has2 = lane < n2;
            float4 v2[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v2[i] = has2 ? PP_LD(kv2 + (32 + i) * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
            wait_big();
#pragma unroll
            for (int hd = 0; hd < 2; ++hd) {
              float v[16], p[16];
              tmem_ld16(lane_addr + BIG + sub * 32 + hd * 16, v);
              tmem_ld_wait();
              if (hd == 1) warp_arrive(&bars.big_free);
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                float4 vv = v2[hd * 4 + q4];
                if (!has2) vv = make_float4(0.f, 0.f, 0.f, 0.f);                // masked rows may hold anything
                p[4*q4]     = fmaf(w1[hd], v[4*q4],     w2[hd] * vv.x);
                p[4*q4 + 1] = fmaf(w1[hd], v[4*q4 + 1], w2[hd] * vv.y);
                p[4*q4 + 2] = fmaf(w1[hd], v[4*q4 + 2], w2[hd] * vv.z);
                p[4*q4 + 3] = fmaf(w1[hd], v[4*q4 + 3], w2[hd] * vv.w);
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) p[i] += __shfl_xor_sync(0xffffffffu, p[i], 16);
#pragma unroll
              for (int off = 8, nn = 8; off >= 1; off >>= 1, nn >>= 1) {
                const bool upper = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < nn; ++i) {
                  const float send = upper ? p[i] : p[i + nn];
                  const float keep = upper ? p[i + nn] : p[i];
                  p[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
              }
              if (lane < 16) put_relay_operand<NPASS>(rb, quarter, sub * 32 + hd * 16 + lane, p[0]);
            }
            fence_async_smem();
            warp_arrive(&bars.rb_ready);
          } else {
            // ================= R3 (J7 in G7, transposed): s'[sentence sub][feature 32*quarter + lane] = relu(att_r @ Wo_relay + b)
            const int f = quarter * 32 + lane;
            mbar_wait(&bars.g7_full, g7_use & 1); ++g7_use;
            tc_fence_after();
            PP_TR();
            float v = tmem_ld1(lane_addr + G7 + sub);
            tmem_ld_wait();
            v = fmaxf(v + bias_s[1][f], 0.f);
            s_cur[T][sub][f] = v;
            const bool next_j8 = plan.j8_after(c);
            if (!last) {
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
                warp_arrive(&bars.rb_ready);
              }
            }
            compute_warps_sync();                                              // s' of the four sentences is complete
            if (last) {
              if (lane == 31) {
                float4* xr = reinterpret_cast<float4*>(Xrow + ((int64_t)t * 128 + row_in_tile) * 128 + sub * 32);
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) xr[q4] = reinterpret_cast<const float4*>(my_s)[q4];
              }
            } else {
              // patch the relay row of the slot's X operand with s' (J5/J6, its last readers, have completed)
#pragma unroll
              for (int part = 0; part < ((NPASS == 3) ? 2 : 1); ++part) {
                uint32_t w[16];
                const uint32_t col = lane_addr + (part ? axl : axh) + sub * 16;
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
              tmem_st_wait();
              warp_arrive(&bars.x_ready[T]);
            }
          }
          PP_TR();
        }
      }
      if (++qA == Lpad) { qA = 0; ++tiA; }
      if (++qB == Lpad) { qB = 0; ++tiB; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == pp::kMmaWarp) tmem_dealloc<512>(tmem_base);
}

template <int NPASS>
static int launch_pp(const float* xi0, const float* s0, const float* q0, const float* kvei, const float* kv2i, int n2,
                     const sf::Weights& w, const float* bias_o, const float* bias_r, float* xrow, int n_tiles,
                     int n_cycles, int flags, cudaStream_t s) {
  constexpr size_t smem = (size_t)pp::RING * pp::RSTAGE + pp::ATT_BYTES + 2 * sf::RB_BYTES + 1024;
  cudaError_t e = cudaFuncSetAttribute(star_pp_kernel<NPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("dsc_star_cycles_tc: %s", cudaGetErrorString(e)); return DSC_ERR_CUDA; }
  const int pairs = (n_tiles + 1) / 2;
  const int grid = pairs < kSMs ? pairs : kSMs;                  // every CTA owns at least two tiles (one per slot)
#ifdef DSC_DEBUG_TOOLS
  unsigned long long* trace = g_pp_trace_host;
#else
  unsigned long long* trace = nullptr;
#endif
  star_pp_kernel<NPASS><<<grid, pp::kThreads, smem, s>>>(xi0, s0, q0, kvei, kv2i, n2, w, bias_o, bias_r, xrow, n_tiles, n_cycles, flags, trace);
  return check_launch("dsc_star_cycles_tc");
}

int launch_star_pp(const float* xi0, const float* s0, const float* q0, const float* kvei, const float* kv2i, int n2,
                   const sf::Weights& w, const float* bias_o, const float* bias_r, float* xrow, int n_tiles, int n_cycles,
                   int flags, int npass, cudaStream_t s) {
  return npass == 3 ? launch_pp<3>(xi0, s0, q0, kvei, kv2i, n2, w, bias_o, bias_r, xrow, n_tiles, n_cycles, flags, s)
                    : launch_pp<1>(xi0, s0, q0, kvei, kv2i, n2, w, bias_o, bias_r, xrow, n_tiles, n_cycles, flags, s);
}

}  // namespace dsc
