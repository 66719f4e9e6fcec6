// Fused star-cycle phase kernels on tcgen05 (sm_100a): persistent, warp-specialised, weights resident in
// shared memory, the activation operand staged in tensor memory (TS-mode UMMA), accumulators in TMEM.
//
//  star_sat_kernel  (K2+K3): per tile of 4 sentences (128 rows = TMEM lanes; a warp quarter is one sentence,
//      lane 31 its relay row)  QKV = X @ [Wq|Wk|Wv]_satellite by head pairs (N = 96 per UMMA group, double-
//      buffered accumulators) -> satellite attention over the five keys {h[i+1], h[i], h[i-1], e[i], s} with
//      neighbour rows fetched by warp shuffle -> ATT rows (fp32) to HBM.  The [rows,384] QKV tensor of the
//      unfused path never exists.
//  star_mix_kernel  (K2+K4): ATT @ Wo_sat + b, ReLU -> X' (written back and re-staged as the next operand
//      in TMEM) -> K|V = X' @ [Wk|Wv]_relay (N = 256) -> relay attention of each sentence's s row over its
//      32 tile rows plus the cached h2 keys -> per-sentence attention output (before the relay dense).
//
// Data layout in HBM ("interleaved tile"): thread t of the CTA owns row t (TMEM lane t), so every tensor the
// warps stream per row is stored [tile][k/4][row][4 floats]: for a fixed k/4 the 32 lanes of a warp read or
// write 512 contiguous bytes (4 L1 wavefronts instead of 32 for a row-major [row][k] layout).
//   XI / ATTI [n_tiles][32][128][4], KVEI [n_tiles][64][128][4], KV2I [n_sent][64][32][4].
// The relay node lives in its own compact buffer S [n_sent][128] (row 31 of a tile is not stored in XI).
//
// Arithmetic: prec 1 = bf16x3 split (fp32-class), prec 2 = single bf16 pass; fp32 accumulation, fp32 softmax.
#include "dsc_star_common.cuh"

namespace dsc {

using namespace tc;

// ===================================================================================== satellite phase
// Packed weight column order (host side, see modules.star_cycles): for head pair g: [q(2 heads x 16) | k | v].
template <int NPASS>
__global__ void __launch_bounds__(kThreads, 1)
star_sat_kernel(const float* __restrict__ XI, const float* __restrict__ Sbuf, const float* __restrict__ KVEI,
                const uint8_t* __restrict__ wblob, float* __restrict__ ATTI, int n_tiles, int dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sW = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) Bars bars;
  __shared__ uint32_t tmem_base_s;
  constexpr uint32_t W_PLANE = 384 * 128;                          // one (part, kb) plane
  constexpr int parts = (NPASS == 3) ? 2 : 1;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bars.w_full, 1);
    mbar_init(&bars.a_full, kLoaderWarps * 32);
    mbar_init(&bars.a_free, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars.acc_full[b], 1); mbar_init(&bars.acc_free[b], kEpiWarps * 32);
      mbar_init(&bars.kve_full[b], kLoaderWarps * 32); mbar_init(&bars.kve_free[b], kEpiWarps * 32);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < kLoaderWarps) {
    // ------------------------------------------------------------ loaders: X rows -> bf16 hi/lo -> TMEM
    const int quarter = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int row_in_tile = quarter * 32 + lane;
    // Each loader thread also stages the e-keys of its row for head `2g + half` into a TMEM slot (double
    // buffered over the head pairs g), so that the epilogue warps never wait on an HBM load.
    uint32_t hi[32], lo[32];
    uint32_t kv[32];                                              // k[16] | v[16] of (row, head) as raw fp32 bits
    auto load_kve = [&](int t, int g) {
      const uint4* base = reinterpret_cast<const uint4*>(KVEI + (int64_t)t * 32768) + row_in_tile;
      const int head = 2 * g + half;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 a = __ldg(base + (head * 4 + q) * 128), c = __ldg(base + (32 + head * 4 + q) * 128);
        kv[4*q] = a.x; kv[4*q+1] = a.y; kv[4*q+2] = a.z; kv[4*q+3] = a.w;
        kv[16+4*q] = c.x; kv[16+4*q+1] = c.y; kv[16+4*q+2] = c.z; kv[16+4*q+3] = c.w;
      }
    };
    auto load_x = [&](int t) {
      if (lane == 31)   // relay row: compact buffer S[sentence][128]
        load_half_row(reinterpret_cast<const float4*>(Sbuf + ((int64_t)t * 4 + quarter) * 128 + half * 64), 1, hi, lo);
      else
        load_half_row(reinterpret_cast<const float4*>(XI + (int64_t)t * 16384) + (half * 16) * 128 + row_in_tile, 128, hi, lo);
    };
    // Register budget (96/thread): the operand registers (hi/lo, 64) and the e-key registers (32) are never live
    // together - X(t+1) is fetched after the last e-key slot of tile t is staged, the e-keys of (t, 0) right
    // after the operand of tile t is in TMEM.
    int it = 0;
    if ((int)blockIdx.x < n_tiles) load_x(blockIdx.x);
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      mbar_wait(&bars.a_free, (it - 1) & 1);                       // previous tile's UMMAs are done with the operand
      tc_fence_after();
      store_half_row<NPASS>(lane_addr, COL_A_HI, COL_A_LO, half, hi, lo);
      tc_fence_before();
      mbar_arrive(&bars.a_full);
      if (!(dbg & 2)) load_kve(t, 0);
#pragma unroll 1
      for (int g = 0; g < 4; ++g) {
        const int b = g & 1, use = it * 2 + (g >> 1);
        mbar_wait(&bars.kve_free[b], (use - 1) & 1);
        tc_fence_after();
        const uint32_t slot = lane_addr + (b ? COL_KVE1 : COL_KVE0);
        tmem_st16(slot + half * 16, kv);
        tmem_st16(slot + 32 + half * 16, kv + 16);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars.kve_full[b]);
        if (g < 3 && !(dbg & 2)) load_kve(t, g + 1);
      }
      if (t + (int)gridDim.x < n_tiles) load_x(t + gridDim.x);     // next tile's operand
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------ UMMA issuer
    if (lane == 0) {
      mbar_expect_tx(&bars.w_full, parts * 2 * W_PLANE);
      for (int p = 0; p < parts * 2; ++p) bulk_g2s(sW + p * W_PLANE, wblob + (size_t)p * W_PLANE, W_PLANE, &bars.w_full);
      mbar_wait(&bars.w_full, 0);
      const uint32_t b_base = smem_u32(sW);
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        mbar_wait(&bars.a_full, it & 1);
        tc_fence_after();
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
          const int b = g & 1, use = it * 2 + (g >> 1);
          mbar_wait(&bars.acc_free[b], (use - 1) & 1);
          tc_fence_after();
          if (!(dbg & 4)) issue_group<NPASS, 96>(tmem_base, b ? COL_ACC1 : COL_ACC0, COL_A_HI, COL_A_LO, b_base, W_PLANE, (uint32_t)g * 96u);
          umma_commit(&bars.acc_full[b]);
        }
        umma_commit(&bars.a_free);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue: satellite attention of one head per warp
    const int ew = warp - kLoaderWarps;
    const int quarter = ew & 3, hh = ew >> 2;                      // sentence within the tile, head within the pair
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int up = (lane >= 30) ? 0 : lane + 1;                    // roll(h,-1)[i] = h[(i+1) mod 31]
    const int dn = (lane == 0) ? 30 : lane - 1;                    // roll(h,+1)[i] = h[(i-1) mod 31]
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int row_in_tile = quarter * 32 + lane;
      float4* att = reinterpret_cast<float4*>(ATTI + (int64_t)t * 16384) + row_in_tile;
#pragma unroll 1
      for (int g = 0; g < 4; ++g) {
        const int b = g & 1, use = it * 2 + (g >> 1);
        const int head = g * 2 + hh;
        mbar_wait(&bars.kve_full[b], use & 1);                     // e-keys of this row/head, staged by the loaders
        mbar_wait(&bars.acc_full[b], use & 1);
        tc_fence_after();
        const uint32_t col = lane_addr + (b ? COL_ACC1 : COL_ACC0) + hh * 16;
        const uint32_t kcol = lane_addr + (b ? COL_KVE1 : COL_KVE0) + hh * 16;
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f, l4 = 0.f;
        if (dbg & 1) {
          tc_fence_before();
          mbar_arrive(&bars.acc_free[b]);
          mbar_arrive(&bars.kve_free[b]);
          continue;
        }
        {
          float q[16], k[16], ke[16];
          tmem_ld16(col, q);
          tmem_ld16(col + 32, k);
          tmem_ld16(kcol, ke);
          tmem_ld_wait();
#pragma unroll
          for (int d = 0; d < 16; ++d) {
            const float ku = __shfl_sync(0xffffffffu, k[d], up);
            const float kd = __shfl_sync(0xffffffffu, k[d], dn);
            const float ks = __shfl_sync(0xffffffffu, k[d], 31);
            l0 = fmaf(q[d], ku, l0);
            l1 = fmaf(q[d], k[d], l1);
            l2 = fmaf(q[d], kd, l2);
            l3 = fmaf(q[d], ke[d], l3);
            l4 = fmaf(q[d], ks, l4);
          }
        }
        float v[16], ve[16];
        tmem_ld16(col + 64, v);
        tmem_ld16(kcol + 32, ve);
        l0 *= 0.25f; l1 *= 0.25f; l2 *= 0.25f; l3 *= 0.25f; l4 *= 0.25f;
        const float mx = fmaxf(fmaxf(fmaxf(l0, l1), fmaxf(l2, l3)), l4);
        l0 = expf(l0 - mx); l1 = expf(l1 - mx); l2 = expf(l2 - mx); l3 = expf(l3 - mx); l4 = expf(l4 - mx);
        const float inv = 1.0f / (l0 + l1 + l2 + l3 + l4);
        l0 *= inv; l1 *= inv; l2 *= inv; l3 *= inv; l4 *= inv;
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&bars.acc_free[b]);
        mbar_arrive(&bars.kve_free[b]);
        float o[16];
#pragma unroll
        for (int d = 0; d < 16; ++d) {
          const float vu = __shfl_sync(0xffffffffu, v[d], up);
          const float vd = __shfl_sync(0xffffffffu, v[d], dn);
          const float vs = __shfl_sync(0xffffffffu, v[d], 31);
          float acc = l0 * vu;
          acc = fmaf(l1, v[d], acc);
          acc = fmaf(l2, vd, acc);
          acc = fmaf(l3, ve[d], acc);
          acc = fmaf(l4, vs, acc);
          o[d] = (lane == 31) ? 0.f : acc;                         // relay row carries no satellite output
        }
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) att[(head * 4 + qd) * 128] = make_float4(o[4*qd], o[4*qd+1], o[4*qd+2], o[4*qd+3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc<512>(tmem_base);
}


// ===================================================================================== mix phase
// X' = relu(ATT @ Wo_sat + b) (rows 0..30; row 31 keeps s) -> K|V = X' @ [Wk|Wv]_relay -> relay attention.
// wblob = [Wo_sat planes (4 x 16 KB)] then [Wkv_relay planes (4 x 32 KB)]  (parts = 1: 2 + 2 planes).
template <int NPASS>
__global__ void __launch_bounds__(kThreads, 1)
star_mix_kernel(const float* __restrict__ ATTI, float* __restrict__ XI, float* __restrict__ Xrow,
                const float* __restrict__ Sbuf, const uint8_t* __restrict__ wo_blob,
                const uint8_t* __restrict__ wkv_blob, const float* __restrict__ bias_o, const float* __restrict__ Qr,
                const float* __restrict__ KV2I, int n2, float* __restrict__ ATTR, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sWo = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int parts = (NPASS == 3) ? 2 : 1;
  constexpr uint32_t WO_PLANE = 128 * 128, WKV_PLANE = 256 * 128;
  uint8_t* sWkv = sWo + parts * 2 * WO_PLANE;
  __shared__ __align__(8) Bars bars;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bars.w_full, 1);
    mbar_init(&bars.a_full, kLoaderWarps * 32);
    mbar_init(&bars.a_free, 1);
    mbar_init(&bars.a2_full, kEpiWarps * 32);
    mbar_init(&bars.o_full, 1);
    mbar_init(&bars.o_free, kEpiWarps * 32);
    mbar_init(&bars.kv_full, 1);
    mbar_init(&bars.kv_free, kEpiWarps * 32);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < kLoaderWarps) {
    const int quarter = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int row_in_tile = quarter * 32 + lane;
    uint32_t hi[32], lo[32];
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      load_half_row(reinterpret_cast<const float4*>(ATTI + (int64_t)t * 16384) + (half * 16) * 128 + row_in_tile, 128, hi, lo);
      mbar_wait(&bars.a_free, (it - 1) & 1);
      tc_fence_after();
      store_half_row<NPASS>(lane_addr, MIX_A_HI, MIX_A_LO, half, hi, lo);
      tc_fence_before();
      mbar_arrive(&bars.a_full);
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0) {
      mbar_expect_tx(&bars.w_full, parts * 2 * (WO_PLANE + WKV_PLANE));
      for (int p = 0; p < parts * 2; ++p) {
        bulk_g2s(sWo + p * WO_PLANE, wo_blob + (size_t)p * WO_PLANE, WO_PLANE, &bars.w_full);
        bulk_g2s(sWkv + p * WKV_PLANE, wkv_blob + (size_t)p * WKV_PLANE, WKV_PLANE, &bars.w_full);
      }
      mbar_wait(&bars.w_full, 0);
      const uint32_t wo_base = smem_u32(sWo), wkv_base = smem_u32(sWkv);
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        mbar_wait(&bars.a_full, it & 1);
        mbar_wait(&bars.o_free, (it - 1) & 1);
        tc_fence_after();
        issue_group<NPASS, 128>(tmem_base, MIX_ACC_O, MIX_A_HI, MIX_A_LO, wo_base, WO_PLANE, 0u);
        umma_commit(&bars.o_full);
        mbar_wait(&bars.a2_full, it & 1);
        mbar_wait(&bars.kv_free, (it - 1) & 1);
        tc_fence_after();
        issue_group<NPASS, 256>(tmem_base, MIX_ACC_KV, MIX_A_HI, MIX_A_LO, wkv_base, WKV_PLANE, 0u);
        umma_commit(&bars.kv_full);
        umma_commit(&bars.a_free);
      }
    }
    __syncwarp();
  } else {
    const int ew = warp - kLoaderWarps;
    const int quarter = ew & 3, hh = ew >> 2;                      // sentence within the tile, column half
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int row_in_tile = quarter * 32 + lane;
      const int64_t sent = (int64_t)t * 4 + quarter;
      // ---------------- phase 1: X' = relu(acc + b); relay row keeps s; re-stage as the next operand
      {
        float4* xi = reinterpret_cast<float4*>(XI + (int64_t)t * 16384) + (hh * 16) * 128 + row_in_tile;
        float4* xr = Xrow ? reinterpret_cast<float4*>(Xrow + ((int64_t)t * 128 + row_in_tile) * 128 + hh * 64) : nullptr;
        const float4* srow = reinterpret_cast<const float4*>(Sbuf + sent * 128 + hh * 64);
        mbar_wait(&bars.o_full, it & 1);
        tc_fence_after();
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float v[32];
          tmem_ld32(lane_addr + MIX_ACC_O + hh * 64 + j * 32, v);
          tmem_ld_wait();
          if (lane == 31) {
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) {
              float4 s4 = __ldg(srow + j * 8 + q4);
              v[4*q4] = s4.x; v[4*q4+1] = s4.y; v[4*q4+2] = s4.z; v[4*q4+3] = s4.w;
              if (xr) xr[j * 8 + q4] = s4;
            }
          } else {
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) {
              float4 b4 = __ldg(reinterpret_cast<const float4*>(bias_o + hh * 64 + j * 32) + q4);
              v[4*q4]   = fmaxf(v[4*q4]   + b4.x, 0.f);
              v[4*q4+1] = fmaxf(v[4*q4+1] + b4.y, 0.f);
              v[4*q4+2] = fmaxf(v[4*q4+2] + b4.z, 0.f);
              v[4*q4+3] = fmaxf(v[4*q4+3] + b4.w, 0.f);
              const float4 o4 = make_float4(v[4*q4], v[4*q4+1], v[4*q4+2], v[4*q4+3]);
              if (xr) xr[j * 8 + q4] = o4; else xi[(j * 8 + q4) * 128] = o4;
            }
          }
#pragma unroll
          for (int q2 = 0; q2 < 16; ++q2) split2(v[2*q2], v[2*q2+1], hi[j * 16 + q2], lo[j * 16 + q2]);
        }
        tc_fence_before();
        mbar_arrive(&bars.o_free);
        store_half_row<NPASS>(lane_addr, MIX_A_HI, MIX_A_LO, hh, hi, lo);
        tc_fence_before();
        mbar_arrive(&bars.a2_full);
      }
      // ---------------- phase 2: relay attention, 4 heads per warp, lane = key row (and h2 key `lane` if < n2)
      {
        const float* qr = Qr + sent * 128 + hh * 64;
        // KV2I [sentence][64 k4][32 rows][4]: k columns are k4 0..31, v columns k4 32..63; lane = h2 row
        const float4* kv2 = reinterpret_cast<const float4*>(KV2I + sent * 8192) + (hh * 16) * 32 + lane;
        const bool has2 = lane < n2;
        mbar_wait(&bars.kv_full, it & 1);
        tc_fence_after();
        float w1[4], w2[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          float k[16];
          tmem_ld16(lane_addr + MIX_ACC_KV + hh * 64 + h * 16, k);
          tmem_ld_wait();
          float d1 = 0.f, d2 = 0.f;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 qq = __ldg(reinterpret_cast<const float4*>(qr + h * 16) + q4);
            d1 = fmaf(qq.x, k[4*q4], d1); d1 = fmaf(qq.y, k[4*q4+1], d1);
            d1 = fmaf(qq.z, k[4*q4+2], d1); d1 = fmaf(qq.w, k[4*q4+3], d1);
            if (has2) {
              const float4 kk = __ldg(kv2 + (h * 4 + q4) * 32);
              d2 = fmaf(qq.x, kk.x, d2); d2 = fmaf(qq.y, kk.y, d2); d2 = fmaf(qq.z, kk.z, d2); d2 = fmaf(qq.w, kk.w, d2);
            }
          }
          d1 *= 0.25f;
          d2 = has2 ? d2 * 0.25f : -3.4e38f;
          const float mx = warp_max(fmaxf(d1, d2));
          const float e1 = expf(d1 - mx), e2 = has2 ? expf(d2 - mx) : 0.f;
          const float inv = 1.0f / warp_sum(e1 + e2);
          w1[h] = e1 * inv;
          w2[h] = e2 * inv;
        }
        float p[64];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          float v[16];
          tmem_ld16(lane_addr + MIX_ACC_KV + 128 + hh * 64 + h * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            float4 vv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has2) vv = __ldg(kv2 + (32 + h * 4 + q4) * 32);
            p[h * 16 + 4*q4]     = fmaf(w1[h], v[4*q4],     w2[h] * vv.x);
            p[h * 16 + 4*q4 + 1] = fmaf(w1[h], v[4*q4 + 1], w2[h] * vv.y);
            p[h * 16 + 4*q4 + 2] = fmaf(w1[h], v[4*q4 + 2], w2[h] * vv.z);
            p[h * 16 + 4*q4 + 3] = fmaf(w1[h], v[4*q4 + 3], w2[h] * vv.w);
          }
        }
        tc_fence_before();
        mbar_arrive(&bars.kv_free);
        // reduce-scatter the 64 partial sums over the 32 lanes: lane ends with dims 2*lane, 2*lane+1
#pragma unroll
        for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
          const bool upper = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < n; ++i) {
            const float send = upper ? p[i] : p[i + n];
            const float keep = upper ? p[i + n] : p[i];
            p[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        *reinterpret_cast<float2*>(ATTR + sent * 128 + hh * 64 + 2 * lane) = make_float2(p[0], p[1]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc<512>(tmem_base);
}


// ===================================================================================== layout helpers
// row-major [n_groups*R rows][width] -> interleaved [group][width/4][R][4].  One CTA per (group, 32-column slab).
template <int R>
__global__ void __launch_bounds__(256)
interleave_kernel(const float* __restrict__ src, int64_t src_group_stride, float* __restrict__ dst, int width) {
  __shared__ float4 tile[R][9];
  const int g = blockIdx.x, slab = blockIdx.y;
  const float* in = src + (int64_t)g * src_group_stride + slab * 32;
  for (int idx = threadIdx.x; idx < R * 8; idx += 256) {
    const int row = idx >> 3, c4 = idx & 7;
    tile[row][c4] = __ldg(reinterpret_cast<const float4*>(in + (int64_t)row * width) + c4);
  }
  __syncthreads();
  float4* out = reinterpret_cast<float4*>(dst + (int64_t)g * R * width) + (int64_t)slab * 8 * R;
  for (int idx = threadIdx.x; idx < R * 8; idx += 256) {
    const int k4 = idx / R, row = idx % R;
    out[k4 * R + row] = tile[row][k4];
  }
}

// scatter one h2 row per sentence into KV2I: vals [n_sent][256] -> kv2i[sent][k4][row_index][4]
__global__ void __launch_bounds__(256)
kv2_put_kernel(const float* __restrict__ vals, float* __restrict__ kv2i, int row_index, int n_sent) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_sent * 64) return;
  const int sent = idx >> 6, k4 = idx & 63;
  const float4 v = __ldg(reinterpret_cast<const float4*>(vals) + idx);
  reinterpret_cast<float4*>(kv2i)[((int64_t)sent * 64 + k4) * 32 + row_index] = v;
}

}  // namespace dsc

using namespace dsc;

template <int NPASS>
static int launch_star_sat(const float* x, const float* sbuf, const float* kv_e, const void* w, float* att, int n_tiles, int dbg, cudaStream_t s) {
  constexpr size_t smem = (size_t)(NPASS == 3 ? 4 : 2) * 384 * 128 + 1024;
  cudaError_t e = cudaFuncSetAttribute(star_sat_kernel<NPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("dsc_star_sat_tc: %s", cudaGetErrorString(e)); return DSC_ERR_CUDA; }
  int grid = n_tiles < kSMs ? n_tiles : kSMs;
  star_sat_kernel<NPASS><<<grid, kThreads, smem, s>>>(x, sbuf, kv_e, reinterpret_cast<const uint8_t*>(w), att, n_tiles, dbg);
  return check_launch("dsc_star_sat_tc");
}

extern "C" int dsc_star_sat_tc(const float* x_tile, const float* s_relay, const float* kv_e, const void* packed_wqkv_grouped,
                               float* att, int n_sent, int prec, void* stream) {
  DSC_REQUIRE(x_tile && s_relay && kv_e && packed_wqkv_grouped && att, "dsc_star_sat_tc: null pointer");
  DSC_REQUIRE(n_sent >= 0 && (n_sent % 4) == 0, "dsc_star_sat_tc: n_sent must be a multiple of 4 (one tile = 4 sentences)");
  DSC_REQUIRE(aligned16(x_tile) && aligned16(kv_e) && aligned16(att) && ((uintptr_t)packed_wqkv_grouped & 127u) == 0,
              "dsc_star_sat_tc: misaligned pointer");
  const int dbg = prec >> 8;                    // profiling knobs (tools/time_star.py): 1 no attention math, 2 no e-key loads, 4 no UMMA
  prec &= 255;
  DSC_REQUIRE(prec == 1 || prec == 2, "dsc_star_sat_tc: prec must be 1 (bf16x3) or 2 (bf16)");
  if (n_sent == 0) return DSC_OK;
  return prec == 1 ? launch_star_sat<3>(x_tile, s_relay, kv_e, packed_wqkv_grouped, att, n_sent / 4, dbg, as_stream(stream))
                   : launch_star_sat<1>(x_tile, s_relay, kv_e, packed_wqkv_grouped, att, n_sent / 4, dbg, as_stream(stream));
}

template <int NPASS>
static int launch_star_mix(const float* att, float* x, float* xrow, const float* sbuf, const void* wo, const void* wkv,
                           const float* bias_o, const float* qr, const float* kv2, int n2, float* attr, int n_tiles,
                           cudaStream_t s) {
  constexpr size_t smem = (size_t)(NPASS == 3 ? 4 : 2) * (128 + 256) * 128 + 1024;
  cudaError_t e = cudaFuncSetAttribute(star_mix_kernel<NPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("dsc_star_mix_tc: %s", cudaGetErrorString(e)); return DSC_ERR_CUDA; }
  int grid = n_tiles < kSMs ? n_tiles : kSMs;
  star_mix_kernel<NPASS><<<grid, kThreads, smem, s>>>(att, x, xrow, sbuf, reinterpret_cast<const uint8_t*>(wo),
                                                      reinterpret_cast<const uint8_t*>(wkv), bias_o, qr, kv2, n2, attr,
                                                      n_tiles);
  return check_launch("dsc_star_mix_tc");
}

extern "C" int dsc_star_mix_tc(const float* att, float* x_tile, float* x_rowmajor, const float* s_relay,
                               const void* packed_wo, const void* packed_wkv_relay,
                               const float* bias_o, const float* q_relay, const float* kv2, int n2,
                               float* att_relay, int n_sent, int prec, void* stream) {
  DSC_REQUIRE(att && x_tile && s_relay && packed_wo && packed_wkv_relay && bias_o && q_relay && att_relay, "dsc_star_mix_tc: null pointer");
  DSC_REQUIRE(n_sent >= 0 && (n_sent % 4) == 0, "dsc_star_mix_tc: n_sent must be a multiple of 4");
  DSC_REQUIRE(n2 >= 0 && n2 <= 32 && (n2 == 0 || kv2), "dsc_star_mix_tc: bad h2 key count");
  DSC_REQUIRE(aligned16(att) && aligned16(x_tile) && aligned16(q_relay) && aligned16(att_relay) && aligned16(bias_o) &&
              aligned16(s_relay) && (!x_rowmajor || aligned16(x_rowmajor)) && (!kv2 || aligned16(kv2)), "dsc_star_mix_tc: misaligned pointer");
  if (n2 == 0) kv2 = att;   // never dereferenced (lane < n2 is false); keeps the pointer arithmetic defined
  DSC_REQUIRE(prec == 1 || prec == 2, "dsc_star_mix_tc: prec must be 1 (bf16x3) or 2 (bf16)");
  if (n_sent == 0) return DSC_OK;
  return prec == 1 ? launch_star_mix<3>(att, x_tile, x_rowmajor, s_relay, packed_wo, packed_wkv_relay, bias_o, q_relay, kv2, n2,
                                        att_relay, n_sent / 4, as_stream(stream))
                   : launch_star_mix<1>(att, x_tile, x_rowmajor, s_relay, packed_wo, packed_wkv_relay, bias_o, q_relay, kv2, n2,
                                        att_relay, n_sent / 4, as_stream(stream));
}

extern "C" int dsc_star_interleave(const float* src, int64_t src_group_stride, float* dst, int n_groups, int group_rows,
                                   int width, void* stream) {
  DSC_REQUIRE(src && dst && n_groups >= 0, "dsc_star_interleave: bad argument");
  DSC_REQUIRE((group_rows == 128 || group_rows == 32) && width > 0 && (width % 32) == 0, "dsc_star_interleave: group_rows must be 128 or 32, width a multiple of 32");
  DSC_REQUIRE(aligned16(src) && aligned16(dst) && (src_group_stride & 3) == 0, "dsc_star_interleave: misaligned pointer");
  if (n_groups == 0) return DSC_OK;
  dim3 grid(n_groups, width / 32);
  if (group_rows == 128) interleave_kernel<128><<<grid, 256, 0, as_stream(stream)>>>(src, src_group_stride, dst, width);
  else interleave_kernel<32><<<grid, 256, 0, as_stream(stream)>>>(src, src_group_stride, dst, width);
  return check_launch("dsc_star_interleave");
}

extern "C" int dsc_star_kv2_put(const float* vals, float* kv2i, int row_index, int n_sent, void* stream) {
  DSC_REQUIRE(vals && kv2i && row_index >= 0 && row_index < 32 && n_sent >= 0, "dsc_star_kv2_put: bad argument");
  DSC_REQUIRE(aligned16(vals) && aligned16(kv2i), "dsc_star_kv2_put: misaligned pointer");
  if (n_sent == 0) return DSC_OK;
  kv2_put_kernel<<<(n_sent * 64 + 255) / 256, 256, 0, as_stream(stream)>>>(vals, kv2i, row_index, n_sent);
  return check_launch("dsc_star_kv2_put");
}
