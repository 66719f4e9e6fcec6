// Tail of the causal target branch of a star decoder step, fused (models/modules.py:352-354 + the relay key/value
// projection of :375-377):  h2 = LayerNorm1(x_t + attn @ Wo + b)  ->  k|v = h2 @ [Wk|Wv]_relay  -> row `row_index` of the
// interleaved key cache KV2I [sentence][64][32][4] that the star-cycle kernels read.
// In the greedy loop this replaces four launches (Dense, residual+LayerNorm, Dense, cache put) whose 128-column rows are
// far too small to fill the machine one after the other; here a CTA owns 128 sentences (thread = row = TMEM lane), both
// weight matrices are resident in shared memory (bulk-copied once), the Dense output never leaves the SM: accumulator ->
// registers (+ bias + residual, LayerNorm computed per thread over its own 128 values) -> bf16 hi/lo A operand in TMEM ->
// second UMMA -> cache rows.
#include "dsc_common.cuh"
#include "dsc_tc.cuh"

namespace dsc {

using namespace tc;

template <int NPASS>
__global__ void __launch_bounds__(128, 1)
tar_tail_kernel(const float* __restrict__ attn, int64_t ld_attn, const float* __restrict__ resid, int64_t ld_resid,
                const uint8_t* __restrict__ wo_blob, const float* __restrict__ bias_o,
                const float* __restrict__ gamma, const float* __restrict__ beta,
                const uint8_t* __restrict__ wkv_blob, float* __restrict__ kv2i, int row_index,
                float* __restrict__ kv_rows, int64_t ld_kv, float* __restrict__ h2_out, int64_t ld_h2, int M) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sWo = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int parts = (NPASS == 3) ? 2 : 1;
  constexpr uint32_t WO_PLANE = 128 * 128, WKV_PLANE = 256 * 128;
  uint8_t* sWkv = sWo + parts * 2 * WO_PLANE;
  __shared__ __align__(8) uint64_t bar_wo, bar_wkv, bar_mma;
  __shared__ uint32_t tmem_base_s;
  constexpr uint32_t ACC = 0, A_HI = 256, A_LO = 320;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = blockIdx.x * 128 + tid;
  if (tid == 0) {
    mbar_init(&bar_wo, 1); mbar_init(&bar_wkv, 1); mbar_init(&bar_mma, 1);
    fence_barrier_init();
    mbar_expect_tx(&bar_wo, parts * 2 * WO_PLANE);
    for (int p = 0; p < parts * 2; ++p) bulk_g2s(sWo + p * WO_PLANE, wo_blob + (size_t)p * WO_PLANE, WO_PLANE, &bar_wo);
    mbar_expect_tx(&bar_wkv, parts * 2 * WKV_PLANE);
    for (int p = 0; p < parts * 2; ++p) bulk_g2s(sWkv + p * WKV_PLANE, wkv_blob + (size_t)p * WKV_PLANE, WKV_PLANE, &bar_wkv);
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);

  // ---- A operand 1: this thread's attention row as bf16 hi/lo pairs
  {
    const float4* src = reinterpret_cast<const float4*>(attn + (int64_t)row * ld_attn);
#pragma unroll 1
    for (int c8 = 0; c8 < 8; ++c8) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = (row < M) ? __ldg(src + c8 * 4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        split2(v.x, v.y, hi[2 * q], lo[2 * q]);
        split2(v.z, v.w, hi[2 * q + 1], lo[2 * q + 1]);
      }
      tmem_st8(lane_addr + A_HI + c8 * 8, hi);
      if (NPASS == 3) tmem_st8(lane_addr + A_LO + c8 * 8, lo);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  auto issue = [&](uint32_t b_base, uint32_t plane, int n) {
    tc_fence_after();
#pragma unroll
    for (int pass = 0; pass < NPASS; ++pass) {
      const uint32_t a_col = (pass == 1) ? A_LO : A_HI;
      const uint32_t pb = (pass == 2) ? 1u : 0u;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t db = smem_desc_sw128(b_base + (pb * 2 + kb) * plane + ks * 32u);
          const uint32_t idesc = (n == 128) ? idesc_bf16_f32(128, 128) : idesc_bf16_f32(128, 256);
          umma_ts(tmem_base + ACC, tmem_base + a_col + (uint32_t)(kb * 4 + ks) * 8u, db, idesc, (pass > 0 || kb > 0 || ks > 0) ? 1u : 0u);
        }
    }
    umma_commit(&bar_mma);
  };
  if (warp == 0) {                       // warp 0 converged, one elected lane issues (dsc_tc.cuh elect_one)
    const bool leader = elect_one();
    mbar_wait(&bar_wo, 0);
    if (leader) issue(smem_u32(sWo), WO_PLANE, 128);
    __syncwarp();
  }
  // ---- epilogue 1: + bias + residual, LayerNorm over the thread's own 128 values, re-staged as A operand 2
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  {
    float h[128];
    const float4* res = reinterpret_cast<const float4*>(resid + (int64_t)row * ld_resid);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v[32];
      tmem_ld32(lane_addr + ACC + j * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias_o + j * 32) + q4);
        const float4 r4 = (row < M) ? __ldg(res + j * 8 + q4) : make_float4(0.f, 0.f, 0.f, 0.f);
        h[j*32 + 4*q4]     = v[4*q4]     + b4.x + r4.x;
        h[j*32 + 4*q4 + 1] = v[4*q4 + 1] + b4.y + r4.y;
        h[j*32 + 4*q4 + 2] = v[4*q4 + 2] + b4.z + r4.z;
        h[j*32 + 4*q4 + 3] = v[4*q4 + 3] + b4.w + r4.w;
        sum += (h[j*32 + 4*q4] + h[j*32 + 4*q4 + 1]) + (h[j*32 + 4*q4 + 2] + h[j*32 + 4*q4 + 3]);
      }
    }
    const float mean = sum * (1.f / 128.f);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 128; ++i) { const float d = h[i] - mean; var = fmaf(d, d, var); }
    const float inv = 1.0f / sqrtf(var * (1.f / 128.f) + 1e-6f);
    tc_fence_before();
    __syncthreads();                 // every thread has drained the accumulator; the first UMMA is complete: A is free
    tc_fence_after();
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      uint32_t hi[8], lo[8];
      float o[16];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma) + c8 * 4 + q4);
        const float4 e4 = __ldg(reinterpret_cast<const float4*>(beta) + c8 * 4 + q4);
        o[4*q4]     = (h[c8*16 + 4*q4]     - mean) * inv * g4.x + e4.x;
        o[4*q4 + 1] = (h[c8*16 + 4*q4 + 1] - mean) * inv * g4.y + e4.y;
        o[4*q4 + 2] = (h[c8*16 + 4*q4 + 2] - mean) * inv * g4.z + e4.z;
        o[4*q4 + 3] = (h[c8*16 + 4*q4 + 3] - mean) * inv * g4.w + e4.w;
      }
      if (h2_out != nullptr && row < M) {
        float4* dst = reinterpret_cast<float4*>(h2_out + (int64_t)row * ld_h2) + c8 * 4;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) dst[q4] = make_float4(o[4*q4], o[4*q4+1], o[4*q4+2], o[4*q4+3]);
      }
#pragma unroll
      for (int q2 = 0; q2 < 8; ++q2) split2(o[2*q2], o[2*q2+1], hi[q2], lo[q2]);
      tmem_st8(lane_addr + A_HI + c8 * 8, hi);
      if (NPASS == 3) tmem_st8(lane_addr + A_LO + c8 * 8, lo);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    const bool leader = elect_one();
    mbar_wait(&bar_wkv, 0);
    if (leader) issue(smem_u32(sWkv), WKV_PLANE, 256);
    __syncwarp();
  }
  // ---- epilogue 2: k|v of the new h2 row -> cache
  mbar_wait(&bar_mma, 1);
  tc_fence_after();
#pragma unroll 1
  for (int j = 0; j < 8; ++j) {
    float v[32];
    tmem_ld32(lane_addr + ACC + j * 32, v);
    tmem_ld_wait();
    if (row < M) {
      if (kv2i != nullptr) {
        float4* dst = reinterpret_cast<float4*>(kv2i) + ((int64_t)row * 64 + j * 8) * 32 + row_index;
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) dst[q4 * 32] = make_float4(v[4*q4], v[4*q4+1], v[4*q4+2], v[4*q4+3]);
      }
      if (kv_rows != nullptr) {
        float4* dst = reinterpret_cast<float4*>(kv_rows + (int64_t)row * ld_kv) + j * 8;
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) dst[q4] = make_float4(v[4*q4], v[4*q4+1], v[4*q4+2], v[4*q4+3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

}  // namespace dsc

using namespace dsc;

extern "C" int dsc_target_tail_tc(const float* attn, int64_t ld_attn, const float* resid, int64_t ld_resid,
                                  const void* packed_wo, const float* bias_o, const float* gamma, const float* beta,
                                  const void* packed_wkv_relay, float* kv2, int row_index,
                                  float* kv_rows, int64_t ld_kv, float* h2_out, int64_t ld_h2,
                                  int M, int prec, void* stream) {
  DSC_REQUIRE(attn && resid && packed_wo && bias_o && gamma && beta && packed_wkv_relay && (kv2 || kv_rows),
              "dsc_target_tail_tc: null pointer");
  DSC_REQUIRE(M >= 0 && row_index >= 0 && row_index < 32, "dsc_target_tail_tc: bad sizes");
  DSC_REQUIRE(((ld_attn | ld_resid | ld_kv | ld_h2) & 3) == 0 && aligned16(attn) && aligned16(resid) && aligned16(bias_o) &&
              aligned16(gamma) && aligned16(beta) && (!kv2 || aligned16(kv2)) && (!kv_rows || aligned16(kv_rows)) &&
              (!h2_out || aligned16(h2_out)), "dsc_target_tail_tc: rows must be 16-byte aligned");
  DSC_REQUIRE((((uintptr_t)packed_wo | (uintptr_t)packed_wkv_relay) & 127u) == 0, "dsc_target_tail_tc: packed weights must be 128-byte aligned");
  DSC_REQUIRE(prec == 1 || prec == 2, "dsc_target_tail_tc: prec must be 1 (bf16x3) or 2 (bf16)");
  if (M == 0) return DSC_OK;
  cudaStream_t s = as_stream(stream);
  const int grid = (M + 127) / 128;
  cudaError_t e;
  if (prec == 1) {
    constexpr size_t smem = 4 * (size_t)(128 + 256) * 128 + 1024;
    e = cudaFuncSetAttribute(tar_tail_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("dsc_target_tail_tc: %s", cudaGetErrorString(e)); return DSC_ERR_CUDA; }
    tar_tail_kernel<3><<<grid, 128, smem, s>>>(attn, ld_attn, resid, ld_resid, reinterpret_cast<const uint8_t*>(packed_wo), bias_o,
                                               gamma, beta, reinterpret_cast<const uint8_t*>(packed_wkv_relay), kv2, row_index,
                                               kv_rows, ld_kv, h2_out, ld_h2, M);
  } else {
    constexpr size_t smem = 2 * (size_t)(128 + 256) * 128 + 1024;
    e = cudaFuncSetAttribute(tar_tail_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("dsc_target_tail_tc: %s", cudaGetErrorString(e)); return DSC_ERR_CUDA; }
    tar_tail_kernel<1><<<grid, 128, smem, s>>>(attn, ld_attn, resid, ld_resid, reinterpret_cast<const uint8_t*>(packed_wo), bias_o,
                                               gamma, beta, reinterpret_cast<const uint8_t*>(packed_wkv_relay), kv2, row_index,
                                               kv_rows, ld_kv, h2_out, ld_h2, M);
  }
  return check_launch("dsc_target_tail_tc");
}
