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

}  // namespace dsc
