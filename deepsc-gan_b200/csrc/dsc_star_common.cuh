// Shared pieces of the fused star-cycle kernel (dsc_star_fused.cu).
#pragma once
#include "dsc_common.cuh"
#include "dsc_tc.cuh"

namespace dsc {

using namespace tc;

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

namespace sf {
// Weight rings.  One-tile kernel: 2 stages of a whole chunk (64 KB).  Streaming half-chunks (hi planes, then lo planes,
// 32 KB stages, as the two-tile kernel does to fit its ATT buffer) was measured on the one-tile kernel at ring depths 3..6:
// 5 % SLOWER at every depth (the mid-job wait and second commit cost more than the L1 that 3 x 32 KB frees).
constexpr uint32_t RSTAGE = 2 * 128 * 128;                    // the two-tile kernel's ring stage (half a chunk)
#ifndef SF_STAGES
#define SF_STAGES 2
#endif
constexpr int STAGES = SF_STAGES;                             // one-tile kernel: whole chunks; 3 stages are 5 % slower
constexpr uint32_t STAGE_BYTES = 4 * 128 * 128;               // largest chunk: 4 planes x 128 rows x 128 B = 64 KB
constexpr uint32_t RB_PLANE = 16 * 128;                       // relay-vector operand: one (part, kb) plane = 16 rows x 128 B
constexpr uint32_t RB_BYTES = 4 * RB_PLANE;                   // 8 KB behind the ring

struct Weights {
  const uint8_t* qkv;    // grouped [Wq|Wk|Wv]_sat, 384 rows
  const uint8_t* wo;     // Wo_sat, 128 rows
  const uint8_t* wkv;    // [Wk|Wv]_relay, 256 rows
  const uint8_t* wo_r;   // Wo_relay
  const uint8_t* wq_r;   // Wq_relay
};
// chunk j of a cycle: source blob, rows per plane, first row, rows of the blob (n_pad)
__device__ __forceinline__ void chunk_of(const Weights& w, int j, const uint8_t*& blob, uint32_t& rows, uint32_t& row0, uint32_t& n_pad) {
  if (j < 4)       { blob = w.qkv;  rows = 96;  row0 = 96u * j; n_pad = 384; }
  else if (j == 4) { blob = w.wo;   rows = 128; row0 = 0;       n_pad = 128; }
  else if (j == 5) { blob = w.wkv;  rows = 128; row0 = 0;       n_pad = 256; }
  else if (j == 6) { blob = w.wkv;  rows = 128; row0 = 128;     n_pad = 256; }
  else if (j == 7) { blob = w.wo_r; rows = 128; row0 = 0;       n_pad = 128; }
  else             { blob = w.wq_r; rows = 128; row0 = 0;       n_pad = 128; }
}
}  // namespace sf

// 64 fp32 values in shared memory -> 32 hi + 32 lo packed bf16 words (plain loads: load_half_row uses ld.global.nc)
__device__ __forceinline__ void split_half_row_smem(const float* p, uint32_t* hi, uint32_t* lo) {
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const float4 v = reinterpret_cast<const float4*>(p)[q];
    split2(v.x, v.y, hi[2 * q], lo[2 * q]);
    split2(v.z, v.w, hi[2 * q + 1], lo[2 * q + 1]);
  }
}

// 32 fp32 values -> 16 hi + 16 lo packed bf16 words; src = element (k4 = 0, this row), consecutive k4 `stride4` float4 apart
__device__ __forceinline__ void load_quarter_row(const float4* __restrict__ src, int stride4, uint32_t* hi, uint32_t* lo) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 v = __ldg(src + (int64_t)q * stride4);
    split2(v.x, v.y, hi[2 * q], lo[2 * q]);
    split2(v.z, v.w, hi[2 * q + 1], lo[2 * q + 1]);
  }
}
__device__ __forceinline__ void split_quarter_row(const float* v, uint32_t* hi, uint32_t* lo) {
#pragma unroll
  for (int q2 = 0; q2 < 16; ++q2) split2(v[2 * q2], v[2 * q2 + 1], hi[q2], lo[q2]);
}
// operand columns of k = 32*sub .. 32*sub+31: 16 hi columns and 16 lo columns
template <int NPASS>
__device__ __forceinline__ void store_quarter_row(uint32_t lane_addr, uint32_t a_hi, uint32_t a_lo, int sub,
                                                  const uint32_t* hi, const uint32_t* lo) {
  tmem_st16(lane_addr + a_hi + sub * 16, hi);
  if (NPASS == 3) tmem_st16(lane_addr + a_lo + sub * 16, lo);
}
__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(n) : "memory");
}

// relay GEMVs (J7, J8) in transposed form: D[feature][sentence] = W^T (A, the streamed chunk, M = 128) x vectors (B, N = 16
// rows of which 4 are sentences).  passes: w_hi*v_hi, w_hi*v_lo, w_lo*v_hi - the same products in the same order as
// issue_group's hi*hi, lo*hi, hi*lo with the activations as A.
template <int NPASS>
__device__ __forceinline__ void issue_relay_gemv(uint32_t d_tmem, uint32_t w_base, uint32_t v_base) {
  constexpr uint32_t IDESC = idesc_bf16_f32(128, 16);
#pragma unroll
  for (int pass = 0; pass < NPASS; ++pass) {
    const uint32_t pw = (pass == 2) ? 1u : 0u, pv = (pass == 1) ? 1u : 0u;
#pragma unroll
    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_ss(d_tmem, smem_desc_sw128(w_base + (pw * 2 + kb) * (128u * 128u) + ks * 32u),
                smem_desc_sw128(v_base + (pv * 2 + kb) * sf::RB_PLANE + ks * 32u), IDESC, (pass > 0 || kb > 0 || ks > 0) ? 1u : 0u);
  }
}
// one K-block (64 of the 128 k) of a TS-mode job, all passes: J4 starts on the first half of ATT while the satellite
// attention of heads 4..7 is still running (K-block-major accumulation order for this job)
template <int NPASS, int N>
__device__ __forceinline__ void issue_group_kb(uint32_t tmem_base, uint32_t acc_col, uint32_t a_hi, uint32_t a_lo,
                                               uint32_t b_base, uint32_t b_plane_bytes, int kb, bool first) {
  constexpr uint32_t IDESC = idesc_bf16_f32(128, N);
#pragma unroll
  for (int pass = 0; pass < NPASS; ++pass) {
    const uint32_t a_col = (pass == 1) ? a_lo : a_hi;               // hi*hi, lo*hi, hi*lo
    const uint32_t pb = (pass == 2) ? 1u : 0u;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint64_t db = smem_desc_sw128(b_base + (pb * 2 + kb) * b_plane_bytes + ks * 32u);
      umma_ts(tmem_base + acc_col, tmem_base + a_col + (uint32_t)(kb * 4 + ks) * 8u, db, IDESC,
              (first && pass == 0 && ks == 0) ? 0u : 1u);
    }
  }
}
__device__ __forceinline__ void compute_warps_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }   // the 16 compute warps
// one fp32 value -> bf16 hi / lo of element (row, k) of the relay-vector operand
template <int NPASS>
__device__ __forceinline__ void put_relay_operand(uint8_t* rb, int row, int k, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const uint32_t off = ((uint32_t)k >> 6) * sf::RB_PLANE + sw128_offset((uint32_t)row, (uint32_t)k & 63u);
  *reinterpret_cast<__nv_bfloat16*>(rb + off) = h;
  if (NPASS == 3) *reinterpret_cast<__nv_bfloat16*>(rb + 2 * sf::RB_PLANE + off) = __float2bfloat16_rn(v - __bfloat162float(h));
}


// one pass of a TS-mode job (A = bf16 words of the slot's X operand in tensor memory, B = two K-block planes of a half-chunk)
template <int N>
__device__ __forceinline__ void ts_pass(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_base, uint32_t plane, bool first) {
  constexpr uint32_t IDESC = idesc_bf16_f32(128, N);
#pragma unroll
  for (int kb = 0; kb < 2; ++kb)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      umma_ts(d_tmem, a_tmem + (uint32_t)(kb * 4 + ks) * 8u, smem_desc_sw128(b_base + kb * plane + ks * 32u), IDESC,
              (first && kb == 0 && ks == 0) ? 0u : 1u);
}
// one pass of a transposed relay GEMV (A = weight planes of a half-chunk, M = 128; B = relay-vector planes, N = 16)
__device__ __forceinline__ void gemv_pass(uint32_t d_tmem, uint32_t w_base, uint32_t v_base, bool first) {
  constexpr uint32_t IDESC = idesc_bf16_f32(128, 16);
#pragma unroll
  for (int kb = 0; kb < 2; ++kb)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      umma_ss(d_tmem, smem_desc_sw128(w_base + kb * (128u * 128u) + ks * 32u),
              smem_desc_sw128(v_base + kb * sf::RB_PLANE + ks * 32u), IDESC, (first && kb == 0 && ks == 0) ? 0u : 1u);
}
// one K-block of a TS-mode job with the hi / lo weight planes in separate ring stages, all passes (J4 of the one-tile kernel)
template <int NPASS, int N>
__device__ __forceinline__ void issue_group_kb2(uint32_t tmem_base, uint32_t acc_col, uint32_t a_hi, uint32_t a_lo,
                                                uint32_t b_hi, uint32_t b_lo, uint32_t plane, int kb, bool first) {
  constexpr uint32_t IDESC = idesc_bf16_f32(128, N);
#pragma unroll
  for (int pass = 0; pass < NPASS; ++pass) {
    const uint32_t a_col = (pass == 1) ? a_lo : a_hi;               // hi*hi, lo*hi, hi*lo
    const uint32_t bb = (pass == 2) ? b_lo : b_hi;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      umma_ts(tmem_base + acc_col, tmem_base + a_col + (uint32_t)(kb * 4 + ks) * 8u,
              smem_desc_sw128(bb + kb * plane + ks * 32u), IDESC, (first && pass == 0 && ks == 0) ? 0u : 1u);
  }
}

// debug-tools library only: the experimental two-tile form of the fused star layer (debug/dsc_star_pp.cu: two tiles per CTA
// half a cycle apart; dsc_star_fused.cu is the product's one-tile kernel)
int launch_star_pp(const float* xi0, const float* s0, const float* q0, const float* kvei, const float* kv2i, int n2,
                   const sf::Weights& w, const float* bias_o, const float* bias_r, float* xrow, int n_tiles, int n_cycles,
                   int flags, int npass, cudaStream_t s);

}  // namespace dsc
