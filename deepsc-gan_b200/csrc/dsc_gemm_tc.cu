// tcgen05 Dense path: y = act(x @ W + bias) with fp32 x/y in HBM and the product on the 5th-gen tensor
// cores.  prec 1 ("bf16x3"): x = x_hi + x_lo, W = W_hi + W_lo (bf16 pairs), three UMMA passes
// x_hi*W_hi + x_lo*W_hi + x_hi*W_lo accumulated in fp32 in TMEM -> ~2^-17 relative error per product
// (fp32-class).  prec 2: single bf16 pass.
//
// One CTA per 128 x BN output tile: x rows are split to bf16 hi/lo and written K-major with the 128-byte
// swizzle into shared memory by all four warps; the pre-packed weight planes (dsc_pack_weight) arrive by
// cp.async.bulk; one elected thread issues the UMMAs (M=128, N=BN, K=16 each); tcgen05.commit signals an
// mbarrier; the four warps read their 32 TMEM lanes back (tcgen05.ld 32x32b) and apply bias/activation.
#include "dsc_common.cuh"
#include "dsc_tc.cuh"

namespace dsc {

using namespace tc;

constexpr int TC_BM = 128;

// ---------------------------------------------------------------- weight packing
// blob layout: [part (hi, lo)][K/64 blocks][N_pad rows][128 B swizzled], N_pad = round_up(N, 128).
__global__ void __launch_bounds__(256)
pack_weight_kernel(const float* __restrict__ w, int64_t ldw, int K, int N, int n_pad, uint8_t* __restrict__ blob) {
  const int64_t total = (int64_t)n_pad * K;
  const int kblocks = K / 64;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx % n_pad);            // consecutive threads -> consecutive n (coalesced reads of W[k, :])
    const int k = (int)(idx / n_pad);
    const float v = (n < N) ? w[(int64_t)k * ldw + n] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    const int kb = k >> 6;
    const size_t plane = (size_t)n_pad * 128;
    const size_t off = (size_t)kb * plane + sw128_offset((uint32_t)n, (uint32_t)(k & 63));
    *reinterpret_cast<__nv_bfloat16*>(blob + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(blob + (size_t)kblocks * plane + off) = lo;
  }
}

// ---------------------------------------------------------------- GEMM
// KB = K-blocks of 64 staged per chunk.  KB = 2 (one chunk for K = 128) is the low-latency form for few tiles; KB = 1
// halves the shared memory (64 KB at BN = 128), so three CTAs share an SM and the load -> convert -> UMMA -> store
// phases of different tiles overlap: the form for the 73,408-row Dense layers of the channel codec.
template <int BN, int NPASS, int KB>
__global__ void __launch_bounds__(128)
gemm_tc_kernel(const float* __restrict__ x, int64_t ldx, const uint8_t* __restrict__ blob, int n_pad,
               const float* __restrict__ bias, float* __restrict__ y, int64_t ldy,
               int M, int K, int N, int act, int row_mod, int row_skip) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // A planes: [part][kb(KB)][128 rows][128 B]; B planes: [part][kb(KB)][BN rows][128 B]
  constexpr uint32_t A_PLANE = TC_BM * 128, B_PLANE = BN * 128;
  constexpr int KC = KB * 64;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 2 * KB * A_PLANE;
  __shared__ __align__(8) uint64_t bar_b, bar_mma;
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_s[BN];                    // the tile's bias columns: per-element global loads stalled the epilogue

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
  for (int i = tid; i < BN; i += 128) bias_s[i] = (bias != nullptr && n0 + i < N) ? __ldg(bias + n0 + i) : 0.f;
  const int kblocks_total = K / 64;
  const size_t g_plane = (size_t)n_pad * 128;

  if (tid == 0) {
    mbar_init(&bar_b, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<BN>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  constexpr uint32_t IDESC = idesc_bf16_f32(TC_BM, BN);

  const int n_chunks = K / KC;
  for (int c = 0; c < n_chunks; ++c) {
    if (c > 0) {                       // operands of the previous chunk must be consumed before overwriting
      mbar_wait(&bar_mma, (c - 1) & 1);
      tc_fence_after();
    }
    // ---- B: bulk copies of the packed planes for this chunk
    if (tid == 0) {
      constexpr int parts = (NPASS == 3) ? 2 : 1;
      mbar_expect_tx(&bar_b, parts * KB * B_PLANE);
      for (int p = 0; p < parts; ++p)
        for (int kb = 0; kb < KB; ++kb)
          bulk_g2s(sB + (p * KB + kb) * B_PLANE,
                   blob + ((size_t)p * kblocks_total + (size_t)(c * KB + kb)) * g_plane + (size_t)n0 * 128, B_PLANE, &bar_b);
    }
    // ---- A: fp32 rows -> bf16 hi/lo, swizzled K-major.  KB = 2: warp w handles rows w, w+4, ...; lane l holds
    //      k = 4l..4l+3 of the 128-wide chunk.  KB = 1: half-warps take two rows at a time (16 lanes x 4 k = 64).
    if (KB == 2) {
      const int kb_l = lane >> 4;                               // which K-block of the chunk this lane writes
      const uint32_t k_in = (uint32_t)((lane & 15) << 2);       // k offset inside the K-block
#pragma unroll 8
      for (int it = 0; it < 32; ++it) {
        const int r = warp + 4 * it;
        const int gr = m0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < M) v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)gr * ldx + c * KC) + lane);
        uint32_t h0, l0, h1, l1;
        split2(v.x, v.y, h0, l0);
        split2(v.z, v.w, h1, l1);
        const uint32_t off = kb_l * A_PLANE + sw128_offset((uint32_t)r, k_in);
        *reinterpret_cast<uint2*>(sA + off) = make_uint2(h0, h1);
        if (NPASS == 3) *reinterpret_cast<uint2*>(sA + 2 * A_PLANE + off) = make_uint2(l0, l1);
      }
    } else {
      const uint32_t k_in = (uint32_t)((lane & 15) << 2);
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int r = 2 * (warp + 4 * it) + (lane >> 4);
        const int gr = m0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < M) v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)gr * ldx + c * KC) + (lane & 15));
        uint32_t h0, l0, h1, l1;
        split2(v.x, v.y, h0, l0);
        split2(v.z, v.w, h1, l1);
        const uint32_t off = sw128_offset((uint32_t)r, k_in);
        *reinterpret_cast<uint2*>(sA + off) = make_uint2(h0, h1);
        if (NPASS == 3) *reinterpret_cast<uint2*>(sA + A_PLANE + off) = make_uint2(l0, l1);
      }
    }
    fence_async_smem();
    __syncthreads();
    // ---- MMA issue: warp 0 converged, one elected lane issues (dsc_tc.cuh elect_one)
    if (warp == 0) {
      const bool leader = elect_one();
      mbar_wait(&bar_b, c & 1);
      tc_fence_after();
      const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
      if (leader) {
#pragma unroll
      for (int pass = 0; pass < NPASS; ++pass) {
        const int pa = (pass == 1) ? 1 : 0;                   // pass 0: hi*hi, pass 1: lo*hi, pass 2: hi*lo
        const int pb = (pass == 2) ? 1 : 0;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            uint64_t da = smem_desc_sw128(a_base + (pa * KB + kb) * A_PLANE + ks * 32);
            uint64_t db = smem_desc_sw128(b_base + (pb * KB + kb) * B_PLANE + ks * 32);
            umma_ss(tmem_d, da, db, IDESC, (c > 0 || pass > 0 || kb > 0 || ks > 0) ? 1u : 0u);
          }
      }
      umma_commit(&bar_mma);
      }
      __syncwarp();
    }
  }
  // ---- epilogue: TMEM -> registers -> bias/act -> global
  mbar_wait(&bar_mma, (n_chunks - 1) & 1);
  tc_fence_after();
  const int row = m0 + warp * 32 + lane;
  const bool row_ok = row < M && !(row_mod > 0 && (row % row_mod) == row_skip);
  const bool vec_ok = ((ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15u) == 0);
#pragma unroll 1
  for (int j = 0; j < BN / 32; ++j) {
    float v[32];
    tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(j * 32), v);
    tmem_ld_wait();
    const int c0 = n0 + j * 32;
    if (row_ok && c0 < N) {
      float* dst = y + (int64_t)row * ldy + c0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float t = v[q * 4 + e] + bias_s[j * 32 + q * 4 + e];
          if (act == 1) t = fmaxf(t, 0.f);
          o[e] = t;
        }
        if (vec_ok && c0 + q * 4 + 3 < N) {
          *reinterpret_cast<float4*>(dst + q * 4) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (c0 + q * 4 + e < N) dst[q * 4 + e] = o[e];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<BN>(tmem_d);
}


// ---------------------------------------------------------------- bring-up: A operand from TMEM (TS mode)
// Same tile as gemm_tc_kernel but thread t owns row t: it splits its 128 fp32 values into bf16 hi/lo pairs and
// stores them into its own TMEM lane (tcgen05.st 32x32b), columns [A_HI, A_HI+64) and [A_LO, A_LO+64); column c
// holds k = 2c (low half) and 2c+1 (high half) unless `swap_halves`.  K = 128 only.
template <int BN, int NPASS>
__global__ void __launch_bounds__(128, 1)
gemm_tc_ts_kernel(const float* __restrict__ x, int64_t ldx, const uint8_t* __restrict__ blob, int n_pad,
                  const float* __restrict__ bias, float* __restrict__ y, int64_t ldy,
                  int M, int N, int act, int swap_halves) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sB = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr uint32_t B_PLANE = BN * 128;
  __shared__ __align__(8) uint64_t bar_b, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
  const size_t g_plane = (size_t)n_pad * 128;
  if (tid == 0) { mbar_init(&bar_b, 1); mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  const uint32_t A_HI = 256, A_LO = 320;                  // column offsets of the A operand
  constexpr uint32_t IDESC = idesc_bf16_f32(TC_BM, BN);
  if (tid == 0) {
    constexpr int parts = (NPASS == 3) ? 2 : 1;
    mbar_expect_tx(&bar_b, parts * 2 * B_PLANE);
    for (int p = 0; p < parts; ++p)
      for (int kb = 0; kb < 2; ++kb)
        bulk_g2s(sB + (p * 2 + kb) * B_PLANE, blob + ((size_t)p * 2 + kb) * g_plane + (size_t)n0 * 128, B_PLANE, &bar_b);
  }
  {
    const int gr = m0 + tid;
    const float4* src = reinterpret_cast<const float4*>(x + (int64_t)gr * ldx);
    const uint32_t lane_addr = tmem_d + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c8 = 0; c8 < 8; ++c8) {                       // 8 columns = 16 k values per store
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 v = (gr < M) ? __ldg(src + c8 * 4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (swap_halves) { split2(v.y, v.x, hi[2*q], lo[2*q]); split2(v.w, v.z, hi[2*q+1], lo[2*q+1]); }
        else             { split2(v.x, v.y, hi[2*q], lo[2*q]); split2(v.z, v.w, hi[2*q+1], lo[2*q+1]); }
      }
      tmem_st8(lane_addr + A_HI + c8 * 8, hi);
      if (NPASS == 3) tmem_st8(lane_addr + A_LO + c8 * 8, lo);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    mbar_wait(&bar_b, 0);
    tc_fence_after();
    const uint32_t b_base = smem_u32(sB);
#pragma unroll
    for (int pass = 0; pass < NPASS; ++pass) {
      const uint32_t a_col = (pass == 1) ? A_LO : A_HI;
      const int pb = (pass == 2) ? 1 : 0;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint64_t db = smem_desc_sw128(b_base + (pb * 2 + kb) * B_PLANE + ks * 32);
          umma_ts(tmem_d, tmem_d + a_col + (kb * 4 + ks) * 8, db, IDESC, (pass > 0 || kb > 0 || ks > 0) ? 1u : 0u);
        }
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  const int row = m0 + warp * 32 + lane;
#pragma unroll 1
  for (int j = 0; j < BN / 32; ++j) {
    float v[32];
    tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(j * 32), v);
    tmem_ld_wait();
    const int c0 = n0 + j * 32;
    if (row < M)
      for (int e = 0; e < 32; ++e)
        if (c0 + e < N) {
          float t = v[e] + (bias ? __ldg(bias + c0 + e) : 0.f);
          y[(int64_t)row * ldy + c0 + e] = act ? fmaxf(t, 0.f) : t;
        }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_d);
}

template <int BN, int NPASS>
static int launch_gemm_tc_ts(const float* x, int64_t ldx, const uint8_t* blob, int n_pad, const float* bias, float* y,
                             int64_t ldy, int M, int N, int act, int swap, cudaStream_t stream) {
  constexpr size_t smem = 4 * (size_t)BN * 128 + 1024;
  cudaError_t e = cudaFuncSetAttribute(gemm_tc_ts_kernel<BN, NPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("dsc_linear_tc(ts): %s", cudaGetErrorString(e)); return DSC_ERR_CUDA; }
  dim3 grid(n_pad / BN, (M + TC_BM - 1) / TC_BM);
  gemm_tc_ts_kernel<BN, NPASS><<<grid, 128, smem, stream>>>(x, ldx, blob, n_pad, bias, y, ldy, M, N, act, swap);
  return check_launch("dsc_linear_tc(ts)");
}

template <int BN, int NPASS, int KB>
static int launch_gemm_tc(const float* x, int64_t ldx, const uint8_t* blob, int n_pad, const float* bias, float* y,
                          int64_t ldy, int M, int K, int N, int act, int row_mod, int row_skip, cudaStream_t stream) {
  constexpr size_t smem = 2 * KB * (size_t)TC_BM * 128 + 2 * KB * (size_t)BN * 128 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, NPASS, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("dsc_linear_tc: %s", cudaGetErrorString(e)); return DSC_ERR_CUDA; }
    attr_set = true;
  }
  dim3 grid(n_pad / BN, (M + TC_BM - 1) / TC_BM);
  gemm_tc_kernel<BN, NPASS, KB><<<grid, 128, smem, stream>>>(x, ldx, blob, n_pad, bias, y, ldy, M, K, N, act, row_mod, row_skip);
  return check_launch("dsc_linear_tc");
}

int linear_k128_persistent(const float* x, int64_t ldx, const uint8_t* blob, int n_pad, const float* bias, float* y, int64_t ldy,
                           int M, int K, int N, int act, int npass, cudaStream_t s);   // dsc_gemm_k128.cu

int linear_tc(const float*, int64_t, const float*, int64_t, const float*, float*, int64_t, int, int, int, int, int, int,
              int prec, cudaStream_t) {
  set_error("dsc_linear: prec=%d needs pre-packed weights: call dsc_pack_weight once, then dsc_linear_tc", prec);
  return DSC_ERR_UNSUPPORTED;
}

}  // namespace dsc

using namespace dsc;

extern "C" int64_t dsc_packed_weight_bytes(int K, int N) {
  if (K <= 0 || N <= 0 || (K % 64) != 0) return 0;
  int64_t n_pad = ((int64_t)N + 127) / 128 * 128;
  return 2 * (int64_t)(K / 64) * n_pad * 128;
}

extern "C" int dsc_pack_weight(const float* w, int64_t ldw, int K, int N, void* blob, void* stream) {
  DSC_REQUIRE(w && blob, "dsc_pack_weight: null pointer");
  DSC_REQUIRE(K > 0 && (K % 64) == 0 && N > 0 && ldw >= N, "dsc_pack_weight: K must be a positive multiple of 64, ldw >= N");
  DSC_REQUIRE((reinterpret_cast<uintptr_t>(blob) & 127u) == 0, "dsc_pack_weight: blob must be 128-byte aligned");
  int n_pad = (N + 127) / 128 * 128;
  int64_t total = (int64_t)n_pad * K;
  int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  pack_weight_kernel<<<blocks, 256, 0, as_stream(stream)>>>(w, ldw, K, N, n_pad, reinterpret_cast<uint8_t*>(blob));
  return check_launch("dsc_pack_weight");
}

extern "C" int dsc_linear_tc(const float* x, int64_t ldx, const void* packed_w, const float* bias,
                             float* y, int64_t ldy, int M, int K, int N, int act,
                             int row_mod, int row_skip, int prec, void* stream) {
  DSC_REQUIRE(x && packed_w && y, "dsc_linear_tc: null pointer");
  DSC_REQUIRE(M >= 0 && N > 0 && K > 0 && (K % 128) == 0, "dsc_linear_tc: K=%d must be a positive multiple of 128", K);
  DSC_REQUIRE((ldx & 3) == 0 && ldx >= K && aligned16(x), "dsc_linear_tc: x rows must be 16-byte aligned");
  DSC_REQUIRE((reinterpret_cast<uintptr_t>(packed_w) & 127u) == 0, "dsc_linear_tc: packed weights must be 128-byte aligned");
  DSC_REQUIRE(ldy >= N && (act == 0 || act == 1), "dsc_linear_tc: bad ldy/act");
  const int ts_mode = prec & 16, ts_swap = prec & 32;       // bring-up knobs: A operand staged in TMEM
  const int prec_flags = prec;                              // | 64: keep the tiled kernel (A/B timing, tests)
  prec &= 15;
  DSC_REQUIRE(prec == 1 || prec == 2, "dsc_linear_tc: prec must be 1 (bf16x3) or 2 (bf16)");
  if (M == 0) return DSC_OK;
  const int n_pad = (N + 127) / 128 * 128;
  const uint8_t* blob = reinterpret_cast<const uint8_t*>(packed_w);
  cudaStream_t s = as_stream(stream);
  if (ts_mode) {
    DSC_REQUIRE(K == 128 && row_mod == 0, "dsc_linear_tc(ts): K must be 128");
    return prec == 1 ? launch_gemm_tc_ts<128, 3>(x, ldx, blob, n_pad, bias, y, ldy, M, N, act, ts_swap, s)
                     : launch_gemm_tc_ts<128, 1>(x, ldx, blob, n_pad, bias, y, ldy, M, N, act, ts_swap, s);
  }
  // many row tiles (the channel codec's 73,408-row layers): 64 KB CTAs, three per SM, so that the phases of different
  // tiles overlap; few tiles (a greedy step's 2,368 rows): one CTA per SM with the whole K = 128 chunk in one pass
  const bool many = (int64_t)((M + TC_BM - 1) / TC_BM) * (n_pad / 128) >= 2 * kSMs;
  // at least two row tiles per SM: the persistent kernel (K = 128: weights resident; K = 256 .. 512: weight chunks ringed)
  if ((K == 128 || (K <= 512 && n_pad <= 256)) && row_mod == 0 && (M + TC_BM - 1) / TC_BM >= 2 * kSMs && !(prec_flags & 64))
    return linear_k128_persistent(x, ldx, blob, n_pad, bias, y, ldy, M, K, N, act, prec == 1 ? 3 : 1, s);
  if (many)
    return prec == 1 ? launch_gemm_tc<128, 3, 1>(x, ldx, blob, n_pad, bias, y, ldy, M, K, N, act, row_mod, row_skip, s)
                     : launch_gemm_tc<128, 1, 1>(x, ldx, blob, n_pad, bias, y, ldy, M, K, N, act, row_mod, row_skip, s);
  const bool wide = (n_pad % 256) == 0;
  if (prec == 1) {
    return wide ? launch_gemm_tc<256, 3, 2>(x, ldx, blob, n_pad, bias, y, ldy, M, K, N, act, row_mod, row_skip, s)
                : launch_gemm_tc<128, 3, 2>(x, ldx, blob, n_pad, bias, y, ldy, M, K, N, act, row_mod, row_skip, s);
  }
  return wide ? launch_gemm_tc<256, 1, 2>(x, ldx, blob, n_pad, bias, y, ldy, M, K, N, act, row_mod, row_skip, s)
              : launch_gemm_tc<128, 1, 2>(x, ldx, blob, n_pad, bias, y, ldy, M, K, N, act, row_mod, row_skip, s);
}
