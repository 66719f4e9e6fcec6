// Shared pieces of the fused star-cycle kernels (dsc_star_tc.cu: one phase per launch; dsc_star_pair.cu: a
// satellite CTA and a mix CTA per SM pair running every cycle of a greedy step in one launch).
#pragma once
#include "dsc_common.cuh"
#include "dsc_tc.cuh"

namespace dsc {

using namespace tc;

// TMEM column map (512 columns allocated)
constexpr uint32_t COL_ACC0 = 0, COL_ACC1 = 128, COL_A_HI = 256, COL_A_LO = 320, COL_KVE0 = 384, COL_KVE1 = 448;   // satellite kernel
constexpr uint32_t MIX_ACC_O = 0, MIX_ACC_KV = 128, MIX_A_HI = 384, MIX_A_LO = 448;     // mix kernel
constexpr int kLoaderWarps = 8, kEpiWarps = 8;
constexpr int kThreads = (kLoaderWarps + kEpiWarps + 1) * 32;      // 544
constexpr int kMmaWarp = kLoaderWarps + kEpiWarps;                 // 16

struct Bars {
  uint64_t w_full;          // weights landed (once)
  uint64_t a_full;          // operand of the current tile staged in TMEM (256 loader threads)
  uint64_t a_free;          // every UMMA reading the operand has completed (tcgen05.commit)
  uint64_t a2_full;         // mix kernel: X' re-staged by the epilogue warps (256 threads)
  uint64_t acc_full[2];     // accumulator buffer ready (tcgen05.commit)
  uint64_t acc_free[2];     // accumulator buffer drained by the epilogue warps (256 threads)
  uint64_t kv_full;         // mix kernel: K|V accumulators ready
  uint64_t kv_free;         // mix kernel: K|V accumulators drained
  uint64_t o_full;          // mix kernel: dense (Wo) accumulators ready
  uint64_t o_free;          // mix kernel: dense accumulators drained
  uint64_t kve_full[2];     // satellite kernel: e-key slot staged in TMEM by the loader warps
  uint64_t kve_free[2];     // satellite kernel: e-key slot consumed by the epilogue warps
  uint64_t h2_full;         // mix kernel: h2-key partial softmax published by the loader warps
  uint64_t h2_free;         // mix kernel: partials consumed by the epilogue warps
};

// stage one half row (64 fp32 -> 32 hi + 32 lo packed words) into the thread's TMEM lane
// src points at element (k4 = 0, this row); consecutive k4 are `stride4` float4 apart (128 for the interleaved
// tile layout, 1 for a compact row such as the relay buffer).
__device__ __forceinline__ void load_half_row(const float4* __restrict__ src, int stride4, uint32_t* hi, uint32_t* lo) {
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    float4 v = __ldg(src + (int64_t)q * stride4);
    split2(v.x, v.y, hi[2 * q], lo[2 * q]);
    split2(v.z, v.w, hi[2 * q + 1], lo[2 * q + 1]);
  }
}
template <int NPASS>
__device__ __forceinline__ void store_half_row(uint32_t lane_addr, uint32_t a_hi, uint32_t a_lo, int half,
                                               const uint32_t* hi, const uint32_t* lo) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    tmem_st8(lane_addr + a_hi + half * 32 + c * 8, hi + c * 8);
    if (NPASS == 3) tmem_st8(lane_addr + a_lo + half * 32 + c * 8, lo + c * 8);
  }
  tmem_st_wait();
}

// issue the NPASS x 8 UMMAs of one N-group: D[acc_col] = A(tmem) * B(smem rows [n_row0, n_row0 + N))
template <int NPASS, int N>
__device__ __forceinline__ void issue_group(uint32_t tmem_base, uint32_t acc_col, uint32_t a_hi, uint32_t a_lo,
                                            uint32_t b_base, uint32_t b_plane_bytes, uint32_t n_row0) {
  constexpr uint32_t IDESC = idesc_bf16_f32(128, N);
#pragma unroll
  for (int pass = 0; pass < NPASS; ++pass) {
    const uint32_t a_col = (pass == 1) ? a_lo : a_hi;               // hi*hi, lo*hi, hi*lo
    const uint32_t pb = (pass == 2) ? 1u : 0u;
#pragma unroll
    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t db = smem_desc_sw128(b_base + (pb * 2 + kb) * b_plane_bytes + n_row0 * 128u + ks * 32u);
        umma_ts(tmem_base + acc_col, tmem_base + a_col + (uint32_t)(kb * 4 + ks) * 8u, db, IDESC,
                (pass > 0 || kb > 0 || ks > 0) ? 1u : 0u);
      }
  }
}


// coherent (L2) loads for tensors that another CTA rewrites while this kernel runs
__device__ __forceinline__ float4 ld_cg(const float4* p) { return __ldcg(p); }
__device__ __forceinline__ void load_half_row_cg(const float4* src, int stride4, uint32_t* hi, uint32_t* lo) {
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    float4 v = __ldcg(src + (int64_t)q * stride4);
    split2(v.x, v.y, hi[2 * q], lo[2 * q]);
    split2(v.z, v.w, hi[2 * q + 1], lo[2 * q + 1]);
  }
}

}  // namespace dsc
