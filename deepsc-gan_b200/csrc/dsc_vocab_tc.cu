// K12+K13 on tcgen05: ids = argmax_j (x . W[:, j] + bias[j]) with the logits never leaving the SM.
//
// Grid (n_splits, m_tiles).  A CTA owns 256 rows of x (two M=128 halves, staged once as bf16 hi/lo pairs in its own
// tensor memory: TS-mode UMMA, thread = row) and a contiguous range of 64-column vocabulary tiles.  A producer warp
// streams the pre-packed weight planes (dsc_pack_weight) of one tile per pipeline stage with cp.async.bulk into a
// 6-stage shared-memory ring; one elected thread issues the UMMAs (M=128, N=64, K=16; 3 passes for bf16x3) into
// double-buffered TMEM accumulators; eight epilogue warps (thread = row) read the accumulators back, add the bias and
// keep a running (max, first index).  Per-row partial results of the n_splits CTAs of an m-tile meet in the caller's
// workspace; the last CTA to arrive (atomic counter per m-tile) folds them in ascending column order, so ties
// resolve to the smallest index exactly like tf.argmax (utlis/eval.py:112-113), and writes the int32 ids.
#include "dsc_common.cuh"
#include "dsc_tc.cuh"

namespace dsc {

using namespace tc;

namespace vtc {
constexpr int BM = 256, BN = 64, STAGES = 6;
constexpr int kEpiWarps = 8, kProducerWarp = 8, kMmaWarp = 9, kThreads = 320;
constexpr uint32_t PLANE = BN * 128;                        // one (part, kb) plane of a stage: 64 rows x 128 B
constexpr uint32_t STAGE_BYTES = 4 * PLANE;                 // [part(2)][kb(2)] (prec 2 uses the first two planes)
constexpr uint32_t COL_ACC = 0;                             // (buf*2 + mh) * 64
constexpr uint32_t COL_A = 256;                             // mh * 128 (+64 for the lo plane)

struct Bars {
  uint64_t b_full[STAGES], b_empty[STAGES], acc_full[2], acc_empty[2];
};
}  // namespace vtc

template <int NPASS>
__global__ void __launch_bounds__(vtc::kThreads, 1)
vocab_argmax_tc_kernel(const float* __restrict__ x, int64_t ldx, const uint8_t* __restrict__ blob, int n_pad,
                       const float* __restrict__ bias, int M, int N, int tiles_per_split,
                       float* __restrict__ part_val, int32_t* __restrict__ part_idx, int32_t* __restrict__ counters,
                       int32_t* __restrict__ ids, int64_t ids_stride) {
  using namespace vtc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sB = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) Bars bars;
  __shared__ uint32_t tmem_base_s;
  __shared__ int last_flag;
  // per epilogue warp, double-buffered: the 64 bias values of a vocabulary tile (16 lanes fetch one float4 each a tile
  // ahead; every lane then reads them as shared-memory broadcasts instead of issuing 16 global loads per tile)
  __shared__ __align__(16) float bias_s[kEpiWarps][2][BN];
  constexpr int parts = (NPASS == 3) ? 2 : 1;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x, n_splits = gridDim.x, m_tile = blockIdx.y;
  const int m0 = m_tile * BM;
  const int n_tiles = n_pad / BN;
  const int j0 = split * tiles_per_split;
  const int my_tiles = max(0, min(tiles_per_split, n_tiles - j0));

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bars.b_full[s], 1); mbar_init(&bars.b_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&bars.acc_full[b], 1); mbar_init(&bars.acc_empty[b], kEpiWarps * 32); }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < kEpiWarps) {
    const int quarter = warp & 3, mh = warp >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int row = m0 + mh * 128 + quarter * 32 + lane;
    // ---- stage this thread's row as the A operand (bf16 hi/lo pairs, 64 + 64 columns)
    {
      const float4* src = reinterpret_cast<const float4*>(x + (int64_t)row * ldx);
#pragma unroll 1
      for (int c8 = 0; c8 < 8; ++c8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = (row < M) ? __ldg(src + c8 * 4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
          split2(v.x, v.y, hi[2 * q], lo[2 * q]);
          split2(v.z, v.w, hi[2 * q + 1], lo[2 * q + 1]);
        }
        tmem_st8(lane_addr + COL_A + mh * 128 + c8 * 8, hi);
        if (NPASS == 3) tmem_st8(lane_addr + COL_A + mh * 128 + 64 + c8 * 8, lo);
      }
      tmem_st_wait();
    }
    tc_fence_before();
    asm volatile("bar.sync 1, %0;" :: "n"((kEpiWarps + 1) * 32) : "memory");      // epilogue warps + MMA warp
    // ---- running argmax over this CTA's vocabulary tiles
    float best = -INFINITY;
    int best_idx = 0x7fffffff;
    auto fetch_bias = [&](int i) {                 // tile i of this CTA -> bias_s[warp][i & 1]; columns >= N read as 0
      if (lane < 16) {
        const int c0 = (j0 + i) * BN + 4 * lane;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + 3 < N) b4 = __ldg(reinterpret_cast<const float4*>(bias + c0));
        else {
          if (c0 < N) b4.x = __ldg(bias + c0);
          if (c0 + 1 < N) b4.y = __ldg(bias + c0 + 1);
          if (c0 + 2 < N) b4.z = __ldg(bias + c0 + 2);
        }
        reinterpret_cast<float4*>(&bias_s[warp][i & 1][0])[lane] = b4;
      }
    };
    if (my_tiles > 0) fetch_bias(0);
    for (int i = 0; i < my_tiles; ++i) {
      const int b = i & 1;
      __syncwarp();                                // bias of tile i is in place; the buffer of tile i + 1 is free
      if (i + 1 < my_tiles) fetch_bias(i + 1);
      const float4* bs = reinterpret_cast<const float4*>(&bias_s[warp][b][0]);
      mbar_wait(&bars.acc_full[b], (i >> 1) & 1);
      tc_fence_after();
      float v[64];
      tmem_ld32(lane_addr + COL_ACC + (b * 2 + mh) * 64, v);
      tmem_ld32(lane_addr + COL_ACC + (b * 2 + mh) * 64 + 32, v + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&bars.acc_empty[b]);
      const int n0 = (j0 + i) * BN;
      if (n0 + BN <= N) {
#pragma unroll
        for (int c4 = 0; c4 < 16; ++c4) {
          const float4 b4 = bs[c4];
          const float t0 = v[4 * c4] + b4.x, t1 = v[4 * c4 + 1] + b4.y, t2 = v[4 * c4 + 2] + b4.z, t3 = v[4 * c4 + 3] + b4.w;
          if (t0 > best) { best = t0; best_idx = n0 + 4 * c4; }
          if (t1 > best) { best = t1; best_idx = n0 + 4 * c4 + 1; }
          if (t2 > best) { best = t2; best_idx = n0 + 4 * c4 + 2; }
          if (t3 > best) { best = t3; best_idx = n0 + 4 * c4 + 3; }
        }
      } else {
#pragma unroll
        for (int c = 0; c < BN; ++c) {                 // static indices: v[] must stay in registers
          const float t = v[c] + bias_s[warp][b][c];
          if (n0 + c < N && t > best) { best = t; best_idx = n0 + c; }
        }
      }
    }
    if (row < M) {
      part_val[(int64_t)row * n_splits + split] = best;
      part_idx[(int64_t)row * n_splits + split] = best_idx;
    }
  } else if (warp == kProducerWarp) {
    if (lane == 0) {
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i % STAGES;
        mbar_wait(&bars.b_empty[s], ((i / STAGES) - 1) & 1);
        mbar_expect_tx(&bars.b_full[s], parts * 2 * PLANE);
        const size_t n0 = (size_t)(j0 + i) * BN;
        for (int p = 0; p < parts * 2; ++p)
          bulk_g2s(sB + s * STAGE_BYTES + p * PLANE, blob + ((size_t)p * n_pad + n0) * 128, PLANE, &bars.b_full[s]);
      }
    }
    __syncwarp();
  } else {
    asm volatile("bar.sync 1, %0;" :: "n"((kEpiWarps + 1) * 32) : "memory");      // A operands are in TMEM
    tc_fence_after();
    {
      // whole warp in lock-step, one elected lane issues (see dsc_tc.cuh elect_one: 42 instead of 94 cycles per N = 64 UMMA)
      const bool leader = elect_one();
      constexpr uint32_t IDESC = idesc_bf16_f32(128, BN);
      const uint32_t b_base = smem_u32(sB);
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i % STAGES, b = i & 1;
        mbar_wait(&bars.b_full[s], (i / STAGES) & 1);
        mbar_wait(&bars.acc_empty[b], ((i >> 1) - 1) & 1);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int mh = 0; mh < 2; ++mh) {
            const uint32_t d = tmem_base + COL_ACC + (b * 2 + mh) * 64;
#pragma unroll
            for (int pass = 0; pass < NPASS; ++pass) {
              const uint32_t a_col = COL_A + mh * 128 + ((pass == 1) ? 64u : 0u);       // hi*hi, lo*hi, hi*lo
              const uint32_t pb = (pass == 2) ? 1u : 0u;
#pragma unroll
              for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  const uint64_t db = smem_desc_sw128(b_base + s * STAGE_BYTES + (pb * 2 + kb) * PLANE + ks * 32u);
                  umma_ts(d, tmem_base + a_col + (uint32_t)(kb * 4 + ks) * 8u, db, IDESC, (pass > 0 || kb > 0 || ks > 0) ? 1u : 0u);
                }
            }
          }
          umma_commit(&bars.b_empty[s]);
          umma_commit(&bars.acc_full[b]);
        }
        __syncwarp();
      }
    }
  }

  // ---- cross-CTA fold: the last CTA of this m-tile to arrive writes the ids
  __threadfence();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    const int prev = atomicAdd(counters + m_tile, 1);
    last_flag = (prev == n_splits - 1);
    if (last_flag) counters[m_tile] = 0;                     // self-resetting for the next launch
  }
  __syncthreads();
  if (last_flag) {
    __threadfence();
    for (int r = tid; r < BM; r += kThreads) {
      const int row = m0 + r;
      if (row >= M) break;
      float best = -INFINITY;
      int best_idx = 0x7fffffff;
      for (int sp = 0; sp < n_splits; ++sp) {                // ascending column ranges: strict > keeps the first maximum
        const float v = __ldcg(part_val + (int64_t)row * n_splits + sp);
        const int ix = __ldcg(part_idx + (int64_t)row * n_splits + sp);
        if (v > best) { best = v; best_idx = ix; }
      }
      ids[(int64_t)row * ids_stride] = best_idx == 0x7fffffff ? 0 : best_idx;
    }
  }
  if (warp == kMmaWarp) tmem_dealloc<512>(tmem_base);
}

static void vocab_tc_shape(int M, int N, int* m_tiles, int* n_splits, int* tiles_per_split) {
  const int n_pad = (N + 127) / 128 * 128;
  const int n_tiles = n_pad / vtc::BN;
  *m_tiles = (M + vtc::BM - 1) / vtc::BM;
  int want = kSMs / (*m_tiles > 0 ? *m_tiles : 1);
  if (want < 1) want = 1;
  if (want > n_tiles) want = n_tiles;
  *tiles_per_split = (n_tiles + want - 1) / want;
  *n_splits = (n_tiles + *tiles_per_split - 1) / *tiles_per_split;
}

}  // namespace dsc

using namespace dsc;

extern "C" int64_t dsc_vocab_argmax_tc_workspace(int M, int N) {
  if (M <= 0 || N <= 0) return 0;
  int m_tiles, n_splits, tps;
  vocab_tc_shape(M, N, &m_tiles, &n_splits, &tps);
  // [M][n_splits] float + [M][n_splits] int32 + [m_tiles] int32 counters, each section 16-byte aligned
  const int64_t sec = ((int64_t)M * n_splits * 4 + 15) / 16 * 16;
  return 2 * sec + ((int64_t)m_tiles * 4 + 15) / 16 * 16;
}

extern "C" int dsc_vocab_argmax_tc(const float* x, int64_t ldx, const void* packed_w, const float* bias,
                                   int32_t* ids, int64_t ids_stride, void* workspace, int64_t workspace_bytes,
                                   int M, int N, int prec, void* stream) {
  DSC_REQUIRE(x && packed_w && bias && ids && workspace, "dsc_vocab_argmax_tc: null pointer");
  DSC_REQUIRE(M >= 0 && N > 0, "dsc_vocab_argmax_tc: bad shape");
  DSC_REQUIRE((ldx & 3) == 0 && ldx >= 128 && aligned16(x) && aligned16(bias), "dsc_vocab_argmax_tc: x rows / bias must be 16-byte aligned");
  DSC_REQUIRE((reinterpret_cast<uintptr_t>(packed_w) & 127u) == 0 && aligned16(workspace), "dsc_vocab_argmax_tc: misaligned packed weights / workspace");
  DSC_REQUIRE(prec == 1 || prec == 2, "dsc_vocab_argmax_tc: prec must be 1 (bf16x3) or 2 (bf16)");
  DSC_REQUIRE(workspace_bytes >= dsc_vocab_argmax_tc_workspace(M, N), "dsc_vocab_argmax_tc: workspace too small");
  if (M == 0) return DSC_OK;
  int m_tiles, n_splits, tps;
  vocab_tc_shape(M, N, &m_tiles, &n_splits, &tps);
  const int n_pad = (N + 127) / 128 * 128;
  const int64_t sec = ((int64_t)M * n_splits * 4 + 15) / 16 * 16;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  float* part_val = reinterpret_cast<float*>(ws);
  int32_t* part_idx = reinterpret_cast<int32_t*>(ws + sec);
  int32_t* counters = reinterpret_cast<int32_t*>(ws + 2 * sec);
  constexpr size_t smem = (size_t)vtc::STAGES * vtc::STAGE_BYTES + 1024;
  cudaStream_t s = as_stream(stream);
  dim3 grid(n_splits, m_tiles);
  cudaError_t e;
  if (prec == 1) {
    e = cudaFuncSetAttribute(vocab_argmax_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("dsc_vocab_argmax_tc: %s", cudaGetErrorString(e)); return DSC_ERR_CUDA; }
    vocab_argmax_tc_kernel<3><<<grid, vtc::kThreads, smem, s>>>(x, ldx, reinterpret_cast<const uint8_t*>(packed_w), n_pad, bias,
                                                                M, N, tps, part_val, part_idx, counters, ids, ids_stride);
  } else {
    e = cudaFuncSetAttribute(vocab_argmax_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("dsc_vocab_argmax_tc: %s", cudaGetErrorString(e)); return DSC_ERR_CUDA; }
    vocab_argmax_tc_kernel<1><<<grid, vtc::kThreads, smem, s>>>(x, ldx, reinterpret_cast<const uint8_t*>(packed_w), n_pad, bias,
                                                                M, N, tps, part_val, part_idx, counters, ids, ids_stride);
  }
  return check_launch("dsc_vocab_argmax_tc");
}
