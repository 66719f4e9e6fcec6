// Persistent tcgen05 Dense kernel for K = 128 and many rows (the channel codec's 73,408-row layers 128 -> 256 and
// 128 -> 512, models/transceiver.py:89-90, 103-105): y = act(x @ W + bias), fp32 in HBM, bf16x3 (or bf16) on the tensor cores.
//
// The tiled kernel of dsc_gemm_tc.cu gives every 128 x 128 output tile its own CTA, so a 256-wide layer reads and converts
// every x row twice and nothing overlaps inside a CTA: 1.3-1.5 TB/s.  This layer is HBM-bound (64 KB of x in, 128 KB of y
// out per 128-row tile against 1.7 us of UMMAs), so here ONE CTA per SM keeps the packed weights of a 256-column slab
// resident in shared memory (128 KB) and walks its row tiles with three warp roles:
//   warps 0-7  loader: x rows (fp32) -> registers -> bf16 hi/lo K-major swizzled A operand in shared memory (64 KB, single
//              buffer); the NEXT tile's rows are already in flight in registers while this tile's UMMAs run;
//   warp  8    UMMA issuer: 24 (8 for bf16) SS-mode UMMAs M = 128, N <= 256 into one of two TMEM accumulators;
//   warps 9-16 epilogue: TMEM -> registers -> bias / ReLU -> y (two warps per TMEM lane quarter, alternate 32-column groups),
//              overlapping the next tile's load + UMMAs.
// K = 128 nk with nk > 1 (512 -> 128, 256 -> 16 of the same codec): the same roles walk (row tile, K chunk) pairs; a slab is
// then 128 columns wide and its weights (64 KB per K chunk) stream from L2 through a two-stage ring filled by a producer
// warp, the accumulator collects the chunks.
// Algorithmic bytes per row: 4 x (K + N) (x read once per column slab, y written once).
#include "dsc_common.cuh"
#include "dsc_tc.cuh"

namespace dsc {

using namespace tc;

namespace k128 {
constexpr int kLoaders = 8, kIssuer = 8, kEpi = 8, kProducer = kLoaders + 1 + kEpi, kThreads = 32 * (kProducer + 1);   // 8 loader warps, issuer, 8 epilogue warps, weight producer
constexpr uint32_t A_PLANE = 128 * 128;       // [part][kb] plane of the A operand: 128 rows x 128 B
struct Bars { uint64_t w_full[2], w_free[2], a_full, a_free, acc_full[2], acc_free[2]; };
}  // namespace k128

template <int NPASS>
__global__ void __launch_bounds__(k128::kThreads, 1)
gemm_k128_persistent_kernel(const float* __restrict__ x, int64_t ldx, const uint8_t* __restrict__ blob, int n_pad, int n0, int bn,
                            const float* __restrict__ bias, float* __restrict__ y, int64_t ldy, int M, int N, int act, int nk) {
  using namespace k128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int parts = (NPASS == 3) ? 2 : 1;
  const uint32_t b_plane = (uint32_t)bn * 128u;
  uint8_t* sB = smem;                                   // [part][kb][bn rows][128 B]
  uint8_t* sA = smem + parts * 2 * 256 * 128;           // [part][kb][128 rows][128 B]
  float* stage = reinterpret_cast<float*>(sA + parts * 2 * A_PLANE);   // epilogue: 8 warps x [32 rows][32] floats, XOR-swizzled
  __shared__ __align__(8) k128::Bars bars;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float bias_s[256];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles = (M + 127) / 128;
  const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  for (int i = tid; i < 256; i += kThreads) bias_s[i] = (bias != nullptr && i < bn && n0 + i < N) ? __ldg(bias + n0 + i) : 0.f;
  if (tid == 0) {
    for (int st = 0; st < 2; ++st) { mbar_init(&bars.w_full[st], 1); mbar_init(&bars.w_free[st], 1); }
    mbar_init(&bars.a_full, kLoaders);
    mbar_init(&bars.a_free, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&bars.acc_full[b], 1); mbar_init(&bars.acc_free[b], kEpi); }
    fence_barrier_init();
  }
  if (warp == kIssuer) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < kLoaders) {
    // ------------------------------------------------------------------ loader: lane l holds k = 4l..4l+3 of row warp + 8*it
    const int kb_l = lane >> 4;
    const uint32_t k_in = (uint32_t)((lane & 15) << 2);
    float4 v[16];
    const int n_chunks = my_tiles * nk;                     // (row tile, K chunk) pairs of this CTA, chunk-minor
    auto load_chunk = [&](int g) {
      const int i = g / nk, kc = g - i * nk;
      const int m0 = (blockIdx.x + i * gridDim.x) * 128;
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int gr = m0 + warp + kLoaders * it;
        v[it] = (gr < M) ? ld_stream(reinterpret_cast<const float4*>(x + (int64_t)gr * ldx + kc * 128) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    if (n_chunks > 0) load_chunk(0);
    for (int g = 0; g < n_chunks; ++g) {
      if (g > 0) mbar_wait(&bars.a_free, (uint32_t)(g - 1) & 1u);       // the UMMAs of the previous chunk have read sA
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        uint32_t h0, l0, h1, l1;
        split2(v[it].x, v[it].y, h0, l0);
        split2(v[it].z, v[it].w, h1, l1);
        const uint32_t off = kb_l * A_PLANE + sw128_offset((uint32_t)(warp + kLoaders * it), k_in);
        *reinterpret_cast<uint2*>(sA + off) = make_uint2(h0, h1);
        if (NPASS == 3) *reinterpret_cast<uint2*>(sA + 2 * A_PLANE + off) = make_uint2(l0, l1);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.a_full);
      if (g + 1 < n_chunks) load_chunk(g + 1);                          // in flight while this chunk's UMMAs run
    }
  } else if (warp == kProducer) {
    // ------------------------------------------------------------------ weights: once (nk = 1) or a two-stage ring over the K chunks
    if (lane == 0) {
      const int n_loads = (nk == 1) ? (my_tiles > 0 ? 1 : 0) : my_tiles * nk;
      const size_t kplane = (size_t)n_pad * 128;                        // one (part, K-block) plane of the blob
      for (int g = 0; g < n_loads; ++g) {
        const int st = g & 1, kc = g % nk;
        mbar_wait(&bars.w_free[st], ((uint32_t)(g >> 1) - 1u) & 1u);
        mbar_expect_tx(&bars.w_full[st], parts * 2 * b_plane);
        uint8_t* dst = sB + (nk == 1 ? 0 : st * (parts * 2 * b_plane));
        for (int p = 0; p < parts; ++p)
          for (int kb = 0; kb < 2; ++kb)
            bulk_g2s(dst + (p * 2 + kb) * b_plane, blob + ((size_t)(p * 2 * nk + 2 * kc + kb)) * kplane + (size_t)n0 * 128, b_plane, &bars.w_full[st]);
      }
    }
    __syncwarp();
  } else if (warp == kIssuer) {
    // ------------------------------------------------------------------ UMMA issuer
    const bool leader = elect_one();
    const uint32_t a_base = smem_u32(sA), b_base0 = smem_u32(sB);
    const uint32_t idesc = idesc_bf16_f32(128, bn);
    const int n_chunks = my_tiles * nk;
    if (nk == 1 && my_tiles > 0) mbar_wait(&bars.w_full[0], 0);
    for (int g = 0; g < n_chunks; ++g) {
      const int i = g / nk, kc = g - i * nk;
      const uint32_t b = (uint32_t)i & 1u, use = (uint32_t)i >> 1, st = (uint32_t)g & 1u;
      mbar_wait(&bars.a_full, (uint32_t)g & 1u);
      if (nk > 1) mbar_wait(&bars.w_full[st], ((uint32_t)g >> 1) & 1u);
      if (kc == 0) mbar_wait(&bars.acc_free[b], (use - 1) & 1u);        // the epilogue drained this accumulator
      tc_fence_after();
      const uint32_t b_base = b_base0 + (nk == 1 ? 0u : st * (parts * 2 * b_plane));
      if (leader) {
#pragma unroll
        for (int pass = 0; pass < NPASS; ++pass) {
          const uint32_t pa = (pass == 1) ? 1u : 0u, pb = (pass == 2) ? 1u : 0u;   // hi*hi, lo*hi, hi*lo
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_ss(tmem_base + b * 256u, smem_desc_sw128(a_base + (pa * 2 + kb) * A_PLANE + ks * 32u),
                      smem_desc_sw128(b_base + (pb * 2 + kb) * b_plane + ks * 32u), idesc,
                      (kc > 0 || pass > 0 || kb > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&bars.a_free);
        if (nk > 1) umma_commit(&bars.w_free[st]);
        if (kc == nk - 1) umma_commit(&bars.acc_full[b]);
      }
      __syncwarp();
    }
  } else if (warp < kProducer) {
    // ------------------------------------------------------------------ epilogue: warps 9..16, TMEM lane quarter = warp & 3, two
    // warps per quarter taking alternate 32-column groups (one warp per scheduler was latency-bound: ~9 us per tile)
    const int quarter = warp & 3, e_half = (warp - kIssuer - 1) >> 2;
    const bool vec_ok = ((ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15u) == 0);
    for (int i = 0; i < my_tiles; ++i) {
      const uint32_t b = (uint32_t)i & 1u, use = (uint32_t)i >> 1;
      const int t = blockIdx.x + i * gridDim.x;
      mbar_wait(&bars.acc_full[b], use & 1u);
      tc_fence_after();
      const uint32_t acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + b * 256u;
#pragma unroll 1
      for (int j = e_half; j < bn / 32; j += 2) {
        float v[32];
        tmem_ld32(acc + (uint32_t)(j * 32), v);
        tmem_ld_wait();
        if (j + 2 >= bn / 32) {                                         // this warp's last read of the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars.acc_free[b]);
        }
        // bias / ReLU in registers, then through a per-warp XOR-swizzled shared-memory tile so that every store instruction
        // writes whole 128-byte row segments (lane -> row 4k + (lane >> 3), columns 4 * (lane & 7) ..).  Written straight from
        // the accumulator layout (lane = row) an instruction touches 32 rows with 16 bytes each: the same speed while y stays
        // in L2 (73,408 x 256), 7-25 % slower once y goes to HBM (N = 512, or 4 x the rows).
        const int c0 = n0 + j * 32;
        float* tile = stage + (warp - kIssuer - 1) * (32 * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b4 = *reinterpret_cast<const float4*>(&bias_s[j * 32 + q * 4]);
          float4 o;
          o.x = v[q * 4] + b4.x;
          o.y = v[q * 4 + 1] + b4.y;
          o.z = v[q * 4 + 2] + b4.z;
          o.w = v[q * 4 + 3] + b4.w;
          if (act == 1) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          *reinterpret_cast<float4*>(tile + lane * 32 + ((q ^ (lane & 7)) << 2)) = o;
        }
        __syncwarp();
        const int rr = lane >> 3, cg = lane & 7;
        const int col = c0 + cg * 4;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int r = 4 * k + rr;
          const int grow = t * 128 + quarter * 32 + r;
          const float4 o = *reinterpret_cast<const float4*>(tile + r * 32 + ((cg ^ (r & 7)) << 2));
          if (grow < M) {
            float* dst = y + (int64_t)grow * ldy + col;
            if (vec_ok && col + 3 < N) {
              st_stream(reinterpret_cast<float4*>(dst), o);
            } else {
              if (col < N) dst[0] = o.x;
              if (col + 1 < N) dst[1] = o.y;
              if (col + 2 < N) dst[2] = o.z;
              if (col + 3 < N) dst[3] = o.w;
            }
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kIssuer) tmem_dealloc<512>(tmem_base);
}

template <int NPASS>
static int launch_k128(const float* x, int64_t ldx, const uint8_t* blob, int n_pad, const float* bias, float* y, int64_t ldy,
                       int M, int K, int N, int act, cudaStream_t s) {
  constexpr int parts = (NPASS == 3) ? 2 : 1;
  constexpr size_t smem = (size_t)parts * 2 * 256 * 128 + (size_t)parts * 2 * 128 * 128 + k128::kEpi * 32 * 32 * 4 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_k128_persistent_kernel<NPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("dsc_linear_tc: %s", cudaGetErrorString(e)); return DSC_ERR_CUDA; }
    attr_set = true;
  }
  const int n_tiles = (M + 127) / 128, nk = K / 128;
  const int grid = n_tiles < kSMs ? n_tiles : kSMs;
  const int slab = (nk == 1) ? 256 : 128;                     // K > 128: the ring holds two 64 KB chunks of a 128-column slab
  for (int n0 = 0; n0 < n_pad; n0 += slab) {                  // one launch per column slab (x is re-read per slab)
    const int bn = (n_pad - n0) < slab ? (n_pad - n0) : slab;
    gemm_k128_persistent_kernel<NPASS><<<grid, k128::kThreads, smem, s>>>(x, ldx, blob, n_pad, n0, bn, bias, y, ldy, M, N, act, nk);
  }
  return check_launch("dsc_linear_tc");
}

int linear_k128_persistent(const float* x, int64_t ldx, const uint8_t* blob, int n_pad, const float* bias, float* y, int64_t ldy,
                           int M, int K, int N, int act, int npass, cudaStream_t s) {
  return npass == 3 ? launch_k128<3>(x, ldx, blob, n_pad, bias, y, ldy, M, K, N, act, s)
                    : launch_k128<1>(x, ldx, blob, n_pad, bias, y, ldy, M, K, N, act, s);
}

}  // namespace dsc
