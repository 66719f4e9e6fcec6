// Backward kernels: the gradient of every forward op group, needed by the FGM evaluators
// (d loss / d channel symbols, utlis/eval.py:25-44, 197-224) and by the training steps (utlis/trainer.py:12-64,
// utlis/gan_train.py:8-50).  fp32 throughout (the reference's GradientTape is fp32).  Each kernel recomputes the
// cheap forward quantities it needs (softmax weights, LayerNorm statistics) instead of storing them.
#include "dsc_common.cuh"
#include <float.h>

namespace dsc {

// ================================================================== generic fp32 GEMM with transposes, split-K
// C[M,N] (+)= op(A)[M,K] * op(B)[K,N];  A_(m,k) = tA ? A[k*lda+m] : A[m*lda+k];  B_(k,n) = tB ? B[n*ldb+k] : B[k*ldb+n].
// 64x64x16 CTA tile, 256 threads, 4x4 register tile.  gridDim.z > 1: split-K, partial sums meet by atomicAdd.
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
gemm_nt_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
               float* __restrict__ C, int64_t ldc, int M, int N, int K, int k_per_split, int atomic) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = k_begin; k0 < k_end; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      int m, k;
      if (TA) { k = idx >> 6; m = idx & 63; } else { m = idx >> 4; k = idx & 15; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < k_end) v = TA ? __ldg(A + (int64_t)gk * lda + gm) : __ldg(A + (int64_t)gm * lda + gk);
      As[k][m] = v;
      int n, kb;
      if (TB) { n = idx >> 4; kb = idx & 15; } else { kb = idx >> 6; n = idx & 63; }
      const int gn = n0 + n, gkb = k0 + kb;
      float w = 0.f;
      if (gn < N && gkb < k_end) w = TB ? __ldg(B + (int64_t)gn * ldb + gkb) : __ldg(B + (int64_t)gkb * ldb + gn);
      Bs[kb][n] = w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + ty * 4 + i;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c >= N) continue;
      float* dst = C + (int64_t)r * ldc + c;
      if (atomic) atomicAdd(dst, acc[i][j]); else *dst = acc[i][j];
    }
  }
}

__global__ void __launch_bounds__(256)
zero_rows_kernel(float* __restrict__ C, int64_t ldc, int M, int N) {
  const int64_t total = (int64_t)M * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    C[(i / N) * ldc + (i % N)] = 0.f;
}

// ================================================================== bias / activation backward
// dz = dy * (y > 0) when act == 1 (dz may alias dy; skipped when dz == NULL), dbias[c] += sum_r dz[r][c].
__global__ void __launch_bounds__(256)
bias_act_backward_kernel(const float* __restrict__ dy, int64_t ld_dy, const float* __restrict__ y, int64_t ld_y, int act,
                         float* __restrict__ dz, int64_t ld_dz, float* __restrict__ dbias, int M, int N) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float acc = 0.f;
  if (c < N) {
    for (int r = blockIdx.y * 8 + ty; r < M; r += gridDim.y * 8) {
      float g = dy[(int64_t)r * ld_dy + c];
      if (act == 1 && !(y[(int64_t)r * ld_y + c] > 0.f)) g = 0.f;
      if (dz != nullptr) dz[(int64_t)r * ld_dz + c] = g;
      acc += g;
    }
  }
  if (dbias == nullptr) return;
  part[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[w][tx];
    atomicAdd(dbias + c, s);
  }
}

// ================================================================== residual + LayerNorm (x2) backward
// forward: v = x + res; y1 = LN_a(v); out = gb ? LN_b(2*y1) : y1.  One warp per row, float4 per lane.
// dv (= dx = dres) is written; dgamma/dbeta are accumulated (block partials, then atomicAdd; caller zeroes them).
__device__ __forceinline__ void ln_stats(const float4 v, float& mean, float& inv, float4& xh) {
  mean = warp_sum(v.x + v.y + v.z + v.w) * (1.f / 128.f);
  const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
  const float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.f / 128.f);
  inv = 1.0f / sqrtf(var + 1e-6f);
  xh = make_float4(dx * inv, dy * inv, dz * inv, dw * inv);
}
// dxhat -> dinput of the normalisation: inv * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat))
__device__ __forceinline__ float4 ln_back(const float4 dxh, const float4 xh, float inv) {
  const float m1 = warp_sum(dxh.x + dxh.y + dxh.z + dxh.w) * (1.f / 128.f);
  const float m2 = warp_sum(dxh.x * xh.x + dxh.y * xh.y + dxh.z * xh.z + dxh.w * xh.w) * (1.f / 128.f);
  return make_float4(inv * (dxh.x - m1 - xh.x * m2), inv * (dxh.y - m1 - xh.y * m2),
                     inv * (dxh.z - m1 - xh.z * m2), inv * (dxh.w - m1 - xh.w * m2));
}
__device__ __forceinline__ void acc4(float4& a, const float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ float4 mul4(const float4 a, const float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }

__global__ void __launch_bounds__(256)
add_layernorm_backward_kernel(const float* __restrict__ x, int64_t xgs, const float* __restrict__ res, int64_t rgs,
                              const float* __restrict__ ga, const float* __restrict__ ba,
                              const float* __restrict__ gb, const float* __restrict__ bb,
                              const float* __restrict__ dout, int64_t dogs, float* __restrict__ dv, int64_t dvgs,
                              float* __restrict__ dga, float* __restrict__ dba, float* __restrict__ dgb, float* __restrict__ dbb,
                              int n_rows, int group_rows) {
  __shared__ float4 red[4][8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  const float4 g1 = __ldg(reinterpret_cast<const float4*>(ga) + lane);
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(ba) + lane);
  float4 g2 = g1;
  if (gb != nullptr) g2 = __ldg(reinterpret_cast<const float4*>(gb) + lane);
  float4 s_ga = make_float4(0.f, 0.f, 0.f, 0.f), s_ba = s_ga, s_gb = s_ga, s_bb = s_ga;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps_per_grid) {
    const int g = r / group_rows, m = r - g * group_rows;
    float4 v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)g * xgs + (int64_t)m * 128) + lane);
    if (res != nullptr) acc4(v, __ldg(reinterpret_cast<const float4*>(res + (int64_t)g * rgs + (int64_t)m * 128) + lane));
    float4 d = __ldg(reinterpret_cast<const float4*>(dout + (int64_t)g * dogs + (int64_t)m * 128) + lane);
    float mean1, inv1;
    float4 xh1;
    ln_stats(v, mean1, inv1, xh1);
    if (gb != nullptr) {
      float4 u = make_float4(2.f * (xh1.x * g1.x + b1.x), 2.f * (xh1.y * g1.y + b1.y), 2.f * (xh1.z * g1.z + b1.z), 2.f * (xh1.w * g1.w + b1.w));
      float mean2, inv2;
      float4 xh2;
      ln_stats(u, mean2, inv2, xh2);
      acc4(s_gb, mul4(d, xh2));
      acc4(s_bb, d);
      float4 du = ln_back(mul4(d, g2), xh2, inv2);
      d = make_float4(2.f * du.x, 2.f * du.y, 2.f * du.z, 2.f * du.w);       // d(y1)
    }
    acc4(s_ga, mul4(d, xh1));
    acc4(s_ba, d);
    const float4 o = ln_back(mul4(d, g1), xh1, inv1);
    reinterpret_cast<float4*>(dv + (int64_t)g * dvgs + (int64_t)m * 128)[lane] = o;
  }
  red[0][warp][lane] = s_ga; red[1][warp][lane] = s_ba; red[2][warp][lane] = s_gb; red[3][warp][lane] = s_bb;
  __syncthreads();
  if (warp < 4) {
    float* dst = warp == 0 ? dga : warp == 1 ? dba : warp == 2 ? dgb : dbb;
    if (dst != nullptr) {
      float4 s = red[warp][0][lane];
#pragma unroll
      for (int w = 1; w < 8; ++w) acc4(s, red[warp][w][lane]);
      atomicAdd(dst + 4 * lane, s.x); atomicAdd(dst + 4 * lane + 1, s.y);
      atomicAdd(dst + 4 * lane + 2, s.z); atomicAdd(dst + 4 * lane + 3, s.w);
    }
  }
}

// ================================================================== small-L multi-head attention backward
// CTA per sentence, warp per head.  smem: K, V [lk][128]; Q, dO [lq][128]; per-(head, query) max, 1/sum, D.
// pass 1 (lane = query): softmax statistics, D_i = sum_j P_ij dP_ij, dq_i.  pass 2 (lane = key): dk_j, dv_j.
__global__ void __launch_bounds__(256)
mha_attention_backward_kernel(const float* __restrict__ q, int64_t ldq, int64_t qbs,
                              const float* __restrict__ k, const float* __restrict__ v, int64_t ldkv, int64_t kvbs,
                              const float* __restrict__ dout, int64_t ldo, int64_t obs,
                              const float* __restrict__ mask, int64_t mbs, int64_t mqs,
                              const int32_t* __restrict__ key_ids, int64_t kis, int causal, int q_off,
                              float* __restrict__ dq, int64_t lddq, int64_t dqbs,
                              float* __restrict__ dk, float* __restrict__ dv, int64_t lddkv, int64_t dkvbs,
                              int lq, int lk) {
  extern __shared__ __align__(16) float smem[];
  float* Ks = smem;
  float* Vs = Ks + (size_t)lk * 128;
  float* Qs = Vs + (size_t)lk * 128;
  float* Os = Qs + (size_t)lq * 128;
  float* st_m = Os + (size_t)lq * 128;          // [8][64]
  float* st_l = st_m + 8 * 64;
  float* st_d = st_l + 8 * 64;
  __shared__ float padm[64];
  const int b = blockIdx.x;
  for (int idx = threadIdx.x; idx < lk * 32; idx += blockDim.x) {
    const int j = idx >> 5, c = idx & 31;
    reinterpret_cast<float4*>(Ks)[idx] = __ldg(reinterpret_cast<const float4*>(k + (int64_t)b * kvbs + (int64_t)j * ldkv) + c);
    reinterpret_cast<float4*>(Vs)[idx] = __ldg(reinterpret_cast<const float4*>(v + (int64_t)b * kvbs + (int64_t)j * ldkv) + c);
  }
  for (int idx = threadIdx.x; idx < lq * 32; idx += blockDim.x) {
    const int i = idx >> 5, c = idx & 31;
    reinterpret_cast<float4*>(Qs)[idx] = __ldg(reinterpret_cast<const float4*>(q + (int64_t)b * qbs + (int64_t)i * ldq) + c);
    reinterpret_cast<float4*>(Os)[idx] = __ldg(reinterpret_cast<const float4*>(dout + (int64_t)b * obs + (int64_t)i * ldo) + c);
  }
  if (threadIdx.x < 64)
    padm[threadIdx.x] = (key_ids && threadIdx.x < lk && key_ids[(int64_t)b * kis + threadIdx.x] == 0) ? 1.f : 0.f;
  __syncthreads();
  const int head = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* mbase = mask ? mask + (int64_t)b * mbs : nullptr;
  auto dot16 = [&](const float* a, const float* c) {
    float d = 0.f;
#pragma unroll
    for (int t = 0; t < 16; ++t) d = fmaf(a[t], c[t], d);
    return d;
  };
  auto logit = [&](int i, int j, const float* qi) {
    float d = dot16(qi, Ks + j * 128 + head * 16) * 0.25f;
    float m = padm[j];
    if (mbase) m = fmaxf(m, __ldg(mbase + (int64_t)i * mqs + j));
    if (causal && j > q_off + i) m = 1.f;
    return d + m * -1e9f;
  };
  // ---- pass 1
  for (int i = lane; i < lq; i += 32) {
    float qi[16], oi[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) { qi[t] = Qs[i * 128 + head * 16 + t]; oi[t] = Os[i * 128 + head * 16 + t]; }
    float mx = -3.4e38f;
    for (int j = 0; j < lk; ++j) mx = fmaxf(mx, logit(i, j, qi));
    float sum = 0.f, dsum = 0.f;
    for (int j = 0; j < lk; ++j) {
      const float w = expf(logit(i, j, qi) - mx);
      sum += w;
      dsum = fmaf(w, dot16(oi, Vs + j * 128 + head * 16), dsum);
    }
    const float inv = 1.0f / sum;
    const float D = dsum * inv;
    st_m[head * 64 + i] = mx; st_l[head * 64 + i] = inv; st_d[head * 64 + i] = D;
    float g[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) g[t] = 0.f;
    for (int j = 0; j < lk; ++j) {
      const float p = expf(logit(i, j, qi) - mx) * inv;
      const float ds = p * (dot16(oi, Vs + j * 128 + head * 16) - D) * 0.25f;
      const float* kj = Ks + j * 128 + head * 16;
#pragma unroll
      for (int t = 0; t < 16; ++t) g[t] = fmaf(ds, kj[t], g[t]);
    }
    float4* dst = reinterpret_cast<float4*>(dq + (int64_t)b * dqbs + (int64_t)i * lddq + head * 16);
#pragma unroll
    for (int t = 0; t < 4; ++t) dst[t] = make_float4(g[4*t], g[4*t+1], g[4*t+2], g[4*t+3]);
  }
  __syncwarp();
  // ---- pass 2
  for (int j = lane; j < lk; j += 32) {
    float gk[16], gv[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) { gk[t] = 0.f; gv[t] = 0.f; }
    const float* vj = Vs + j * 128 + head * 16;
    for (int i = 0; i < lq; ++i) {
      const float* qi = Qs + i * 128 + head * 16;
      const float* oi = Os + i * 128 + head * 16;
      float qr[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) qr[t] = qi[t];
      const float p = expf(logit(i, j, qr) - st_m[head * 64 + i]) * st_l[head * 64 + i];
      const float ds = p * (dot16(oi, vj) - st_d[head * 64 + i]) * 0.25f;
#pragma unroll
      for (int t = 0; t < 16; ++t) { gk[t] = fmaf(ds, qr[t], gk[t]); gv[t] = fmaf(p, oi[t], gv[t]); }
    }
    float4* dkp = reinterpret_cast<float4*>(dk + (int64_t)b * dkvbs + (int64_t)j * lddkv + head * 16);
    float4* dvp = reinterpret_cast<float4*>(dv + (int64_t)b * dkvbs + (int64_t)j * lddkv + head * 16);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      dkp[t] = make_float4(gk[4*t], gk[4*t+1], gk[4*t+2], gk[4*t+3]);
      dvp[t] = make_float4(gv[4*t], gv[4*t+1], gv[4*t+2], gv[4*t+3]);
    }
  }
}

// ================================================================== star satellite attention backward
// CTA per sentence, warp per head, lane = tile row (31 = relay).  qkv [S*32,384], kv_e [S*32,256], datt [S*32,128]
// -> dqkv [S*32,384] (row 31: dq = 0, dk/dv = sums over the 31 queries), dkv_e [S*32,256] (row 31 zero).
__global__ void __launch_bounds__(256)
star_satellite_attn_backward_kernel(const float* __restrict__ qkv, const float* __restrict__ kv_e,
                                    const float* __restrict__ datt, float* __restrict__ dqkv, float* __restrict__ dkv_e) {
  const int s = blockIdx.x, head = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)s * 32 + lane;
  const bool sat = lane < 31;
  const int up = (lane >= 30) ? 0 : lane + 1;
  const int dn = (lane == 0) ? 30 : lane - 1;
  float q[16], k[16], v[16], ke[16], ve[16], go[16];
  auto ld16 = [&](const float* p, float* o) {
#pragma unroll
    for (int t = 0; t < 4; ++t) { float4 a = __ldg(reinterpret_cast<const float4*>(p) + t); o[4*t] = a.x; o[4*t+1] = a.y; o[4*t+2] = a.z; o[4*t+3] = a.w; }
  };
  auto st16 = [&](float* p, const float* o) {
#pragma unroll
    for (int t = 0; t < 4; ++t) reinterpret_cast<float4*>(p)[t] = make_float4(o[4*t], o[4*t+1], o[4*t+2], o[4*t+3]);
  };
  ld16(qkv + row * 384 + head * 16, q);
  ld16(qkv + row * 384 + 128 + head * 16, k);
  ld16(qkv + row * 384 + 256 + head * 16, v);
  ld16(kv_e + row * 256 + head * 16, ke);
  ld16(kv_e + row * 256 + 128 + head * 16, ve);
  ld16(datt + row * 128 + head * 16, go);
  float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f, l4 = 0.f;       // logits
  float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f, p4 = 0.f;       // dP_j = go . v_j
#pragma unroll
  for (int d = 0; d < 16; ++d) {
    const float ku = __shfl_sync(0xffffffffu, k[d], up), kd = __shfl_sync(0xffffffffu, k[d], dn), ks = __shfl_sync(0xffffffffu, k[d], 31);
    const float vu = __shfl_sync(0xffffffffu, v[d], up), vd = __shfl_sync(0xffffffffu, v[d], dn), vs = __shfl_sync(0xffffffffu, v[d], 31);
    l0 = fmaf(q[d], ku, l0); l1 = fmaf(q[d], k[d], l1); l2 = fmaf(q[d], kd, l2); l3 = fmaf(q[d], ke[d], l3); l4 = fmaf(q[d], ks, l4);
    p0 = fmaf(go[d], vu, p0); p1 = fmaf(go[d], v[d], p1); p2 = fmaf(go[d], vd, p2); p3 = fmaf(go[d], ve[d], p3); p4 = fmaf(go[d], vs, p4);
  }
  l0 *= 0.25f; l1 *= 0.25f; l2 *= 0.25f; l3 *= 0.25f; l4 *= 0.25f;
  const float mx = fmaxf(fmaxf(fmaxf(l0, l1), fmaxf(l2, l3)), l4);
  float w0 = expf(l0 - mx), w1 = expf(l1 - mx), w2 = expf(l2 - mx), w3 = expf(l3 - mx), w4 = expf(l4 - mx);
  const float inv = 1.0f / (w0 + w1 + w2 + w3 + w4);
  w0 *= inv; w1 *= inv; w2 *= inv; w3 *= inv; w4 *= inv;
  if (!sat) { w0 = w1 = w2 = w3 = w4 = 0.f; }                     // the relay row is not a query
  const float D = w0 * p0 + w1 * p1 + w2 * p2 + w3 * p3 + w4 * p4;
  const float s0 = w0 * (p0 - D) * 0.25f, s1 = w1 * (p1 - D) * 0.25f, s2 = w2 * (p2 - D) * 0.25f,
              s3 = w3 * (p3 - D) * 0.25f, s4 = w4 * (p4 - D) * 0.25f;
  float gq[16], gk[16], gv[16], gke[16], gve[16];
#pragma unroll
  for (int d = 0; d < 16; ++d) {
    const float ku = __shfl_sync(0xffffffffu, k[d], up), kd = __shfl_sync(0xffffffffu, k[d], dn), ks = __shfl_sync(0xffffffffu, k[d], 31);
    gq[d] = s0 * ku + s1 * k[d] + s2 * kd + s3 * ke[d] + s4 * ks;
    // key of row i is used by query i (slot 1), by query i-1 as its "up" key (slot 0) and by query i+1 as its "down" key (slot 2)
    const float c0 = s0 * q[d], c2 = s2 * q[d], c4 = s4 * q[d];
    const float from_dn = __shfl_sync(0xffffffffu, c0, dn), from_up = __shfl_sync(0xffffffffu, c2, up);
    const float tot_s = warp_sum(c4);
    gk[d] = sat ? (s1 * q[d] + from_dn + from_up) : tot_s;
    const float e0 = w0 * go[d], e2 = w2 * go[d], e4 = w4 * go[d];
    const float vfrom_dn = __shfl_sync(0xffffffffu, e0, dn), vfrom_up = __shfl_sync(0xffffffffu, e2, up);
    const float vtot_s = warp_sum(e4);
    gv[d] = sat ? (w1 * go[d] + vfrom_dn + vfrom_up) : vtot_s;
    gke[d] = s3 * q[d];
    gve[d] = w3 * go[d];
  }
  st16(dqkv + row * 384 + head * 16, gq);
  st16(dqkv + row * 384 + 128 + head * 16, gk);
  st16(dqkv + row * 384 + 256 + head * 16, gv);
  st16(dkv_e + row * 256 + head * 16, gke);
  st16(dkv_e + row * 256 + 128 + head * 16, gve);
}

// ================================================================== star relay attention backward
// CTA per sentence, warp per head, lane = key j and j + 32 (key 0 = relay row 31, key j in 1..31 = satellite j-1,
// key 32 + r = h2 row r).  dqkv_r [S*32,384] fully written (q columns are zero except the relay row),
// dkv2 [S, kv2_rows, 256]: rows < n2 written, rows >= n2 zero-filled.
__global__ void __launch_bounds__(256)
star_relay_attn_backward_kernel(const float* __restrict__ qkv_r, const float* __restrict__ kv2, int kv2_rows, int n2,
                                const float* __restrict__ dout, float* __restrict__ dqkv_r, float* __restrict__ dkv2) {
  const int s = blockIdx.x, head = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* tile = qkv_r + (int64_t)s * 32 * 384;
  float* dtile = dqkv_r + (int64_t)s * 32 * 384;
  const int trow = (lane == 0) ? 31 : lane - 1;                   // tile row of key `lane`
  float qh[16], go[16], k1[16], v1[16], k2[16], v2[16];
  auto ld16 = [&](const float* p, float* o) {
#pragma unroll
    for (int t = 0; t < 4; ++t) { float4 a = __ldg(reinterpret_cast<const float4*>(p) + t); o[4*t] = a.x; o[4*t+1] = a.y; o[4*t+2] = a.z; o[4*t+3] = a.w; }
  };
  auto st16 = [&](float* p, const float* o) {
#pragma unroll
    for (int t = 0; t < 4; ++t) reinterpret_cast<float4*>(p)[t] = make_float4(o[4*t], o[4*t+1], o[4*t+2], o[4*t+3]);
  };
  ld16(tile + 31 * 384 + head * 16, qh);
  ld16(dout + (int64_t)s * 128 + head * 16, go);
  ld16(tile + (int64_t)trow * 384 + 128 + head * 16, k1);
  ld16(tile + (int64_t)trow * 384 + 256 + head * 16, v1);
  const bool has2 = lane < n2;
  if (has2) {
    ld16(kv2 + ((int64_t)s * kv2_rows + lane) * 256 + head * 16, k2);
    ld16(kv2 + ((int64_t)s * kv2_rows + lane) * 256 + 128 + head * 16, v2);
  } else {
#pragma unroll
    for (int t = 0; t < 16; ++t) { k2[t] = 0.f; v2[t] = 0.f; }
  }
  float d1 = 0.f, d2 = 0.f, e1 = 0.f, e2 = 0.f;
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    d1 = fmaf(qh[t], k1[t], d1); d2 = fmaf(qh[t], k2[t], d2);
    e1 = fmaf(go[t], v1[t], e1); e2 = fmaf(go[t], v2[t], e2);
  }
  d1 *= 0.25f;
  d2 = has2 ? d2 * 0.25f : -3.4e38f;
  const float mx = warp_max(fmaxf(d1, d2));
  float w1 = expf(d1 - mx), w2 = has2 ? expf(d2 - mx) : 0.f;
  const float inv = 1.0f / warp_sum(w1 + w2);
  w1 *= inv; w2 *= inv;
  const float D = warp_sum(w1 * e1 + w2 * e2);
  const float s1 = w1 * (e1 - D) * 0.25f, s2 = w2 * (e2 - D) * 0.25f;
  float gq[16], g[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) gq[t] = warp_sum(s1 * k1[t] + s2 * k2[t]);
  if (lane != 0) {
#pragma unroll
    for (int t = 0; t < 16; ++t) gq[t] = 0.f;                       // only the relay row carries a query
  }
  st16(dtile + (int64_t)trow * 384 + head * 16, gq);
#pragma unroll
  for (int t = 0; t < 16; ++t) g[t] = s1 * qh[t];
  st16(dtile + (int64_t)trow * 384 + 128 + head * 16, g);
#pragma unroll
  for (int t = 0; t < 16; ++t) g[t] = w1 * go[t];
  st16(dtile + (int64_t)trow * 384 + 256 + head * 16, g);
  if (dkv2 != nullptr && lane < kv2_rows) {
    float* o = dkv2 + ((int64_t)s * kv2_rows + lane) * 256;
#pragma unroll
    for (int t = 0; t < 16; ++t) g[t] = s2 * qh[t];                 // zero when lane >= n2 (w2 = 0)
    st16(o + head * 16, g);
#pragma unroll
    for (int t = 0; t < 16; ++t) g[t] = w2 * go[t];
    st16(o + 128 + head * 16, g);
  }
}

// ================================================================== embedding / star pack backward
__global__ void __launch_bounds__(256)
embed_backward_kernel(const int32_t* __restrict__ ids, int64_t ids_stride, const float* __restrict__ dout, int64_t ld,
                      float* __restrict__ dtable, int vocab, int n_rows, int len) {
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  const float scale = 11.313708498984761f;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps_per_grid) {
    const int s = r / len, i = r - s * len;
    int id = __ldg(ids + (int64_t)s * ids_stride + i);
    id = min(max(id, 0), vocab - 1);
    const float4 g = __ldg(reinterpret_cast<const float4*>(dout + (int64_t)r * ld) + lane);
    float* dst = dtable + (int64_t)id * 128 + 4 * lane;
    atomicAdd(dst, g.x * scale); atomicAdd(dst + 1, g.y * scale); atomicAdd(dst + 2, g.z * scale); atomicAdd(dst + 3, g.w * scale);
  }
}

__global__ void __launch_bounds__(128)
star_pack_backward_kernel(const float* __restrict__ dtile, float* __restrict__ dsrc) {
  const int s = blockIdx.x, c = threadIdx.x;
  const float* in = dtile + (int64_t)s * 32 * 128;
  float* o = dsrc + (int64_t)s * 31 * 128;
  const float m = in[31 * 128 + c] * (1.0f / 31.0f);
#pragma unroll
  for (int i = 0; i < 31; ++i) o[i * 128 + c] = in[i * 128 + c] + m;
}

// ================================================================== masked CE backward
// dlogits[r][j] = g[r] * (softmax(logits[r])[j] - [j == t]) for t != 0, else 0.  CTA per row: ONE read of the row (16-byte
// loads into shared memory with an online log-sum-exp) and one write, instead of three scalar passes over 87 KB.
__global__ void __launch_bounds__(256)
masked_ce_backward_kernel(const float* __restrict__ logits, int64_t ld, const int32_t* __restrict__ target,
                          const float* __restrict__ grow, float* __restrict__ dlogits, int64_t ldd, int N) {
  extern __shared__ float srow[];                 // the row, N floats
  __shared__ float red_m[8], red_s[8];
  const int r = blockIdx.x;
  const float* row = logits + (int64_t)r * ld;
  float* drow = dlogits + (int64_t)r * ldd;
  const int t = target[r];
  const float g = grow[r];
  const RowSpan dsp = row_span(drow, N);
  float4* dbody = reinterpret_cast<float4*>(drow + dsp.head);
  if (t == 0 || g == 0.f) {
    if ((int)threadIdx.x < dsp.head) drow[threadIdx.x] = 0.f;
    for (int j4 = threadIdx.x; j4 < dsp.n4; j4 += 256) dbody[j4] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = dsp.head + 4 * dsp.n4 + threadIdx.x; j < N; j += 256) drow[j] = 0.f;
    return;
  }
  const RowSpan sp = row_span(row, N);
  float m = -FLT_MAX, s = 0.f;
  if ((int)threadIdx.x < sp.head) { const float v = __ldg(row + threadIdx.x); srow[threadIdx.x] = v; lse_take(v, m, s); }
  for (int j4 = threadIdx.x; j4 < sp.n4; j4 += 256) {
    const float4 a = ld_stream(sp.body + j4);
    const int j = sp.head + 4 * j4;
    srow[j] = a.x; srow[j + 1] = a.y; srow[j + 2] = a.z; srow[j + 3] = a.w;
    const float m4 = fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w));
    if (m4 > m) { s *= expf(m - m4); m = m4; }
    s += (expf(a.x - m) + expf(a.y - m)) + (expf(a.z - m) + expf(a.w - m));
  }
  for (int j = sp.head + 4 * sp.n4 + threadIdx.x; j < N; j += 256) { const float v = __ldg(row + j); srow[j] = v; lse_take(v, m, s); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, s, o);
    lse_merge(om, os, m, s);
  }
  if ((threadIdx.x & 31) == 0) { red_m[threadIdx.x >> 5] = m; red_s[threadIdx.x >> 5] = s; }
  __syncthreads();
  m = red_m[0]; s = red_s[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) lse_merge(red_m[w], red_s[w], m, s);
  const float sc = g / s;
  auto out = [&](int j) { return expf(srow[j] - m) * sc - (j == t ? g : 0.f); };
  if ((int)threadIdx.x < dsp.head) drow[threadIdx.x] = out(threadIdx.x);
  for (int j4 = threadIdx.x; j4 < dsp.n4; j4 += 256) {
    const int j = dsp.head + 4 * j4;
    st_stream(dbody + j4, make_float4(out(j), out(j + 1), out(j + 2), out(j + 3)));
  }
  for (int j = dsp.head + 4 * dsp.n4 + threadIdx.x; j < N; j += 256) drow[j] = out(j);
}

// ================================================================== power norm / channel backward
__global__ void __launch_bounds__(256)
unit_dot_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int64_t elems_per_unit) {
  __shared__ float part[8];
  const int u = blockIdx.y;
  const float4* pa = reinterpret_cast<const float4*>(a + (int64_t)u * elems_per_unit);
  const float4* pb = reinterpret_cast<const float4*>(b + (int64_t)u * elems_per_unit);
  const int64_t n4 = elems_per_unit >> 2;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 x = ld_stream(pa + i), y = ld_stream(pb + i);
    acc += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(out + u, v);
  }
}

// out = x * r, r = (factor * ss / n)^(-1/2)  =>  dx = r * dy - x * r^3 * (factor / n) * dot(x, dy)
__global__ void __launch_bounds__(256)
power_normalize_backward_kernel(const float4* __restrict__ x, const float* __restrict__ sumsq, const float* __restrict__ dot,
                                float factor, const float4* __restrict__ dy, float4* __restrict__ dx, int64_t n4, int64_t unit4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(i / unit4);
    const float n = (float)(unit4 * 4);
    const float r = 1.0f / sqrtf(factor * __ldg(sumsq + u) / n);
    const float c = r * r * r * (factor / n) * __ldg(dot + u);
    const float4 xv = ld_stream(x + i), g = ld_stream(dy + i);
    dx[i] = make_float4(r * g.x - xv.x * c, r * g.y - xv.y * c, r * g.z - xv.z * c, r * g.w - xv.w * c);
  }
}

// y = channel(xs, ps):  AWGN  y = xs + n + p_scale*p      -> dxs = dy, dp = p_scale * dy
//                       fading y = D(xs*h + n), D = identity | conj(h)/den  -> dxs = conj(h) * (h/den) * dy  (complex)
__global__ void __launch_bounds__(256)
channel_backward_kernel(const float4* __restrict__ dy, const float2* __restrict__ h, const float* __restrict__ n_std,
                        int detector, const float* __restrict__ p_scale, float4* __restrict__ dx, float4* __restrict__ dp,
                        int64_t n4, int64_t unit4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(i / unit4);
    float4 g = ld_stream(dy + i);
    if (h == nullptr) {
      if (dx != nullptr) dx[i] = g;
      if (dp != nullptr) {
        const float ps = p_scale ? __ldg(p_scale + u) : 1.0f;
        dp[i] = make_float4(ps * g.x, ps * g.y, ps * g.z, ps * g.w);
      }
      continue;
    }
    const float2 hh = __ldg(h + u);
    if (detector != 0) {
      // est = y * conj(h) / den  =>  dy = dest * h / den
      const float ns = __ldg(n_std + u);
      float den = hh.x * hh.x + hh.y * hh.y;
      if (detector == 2) den += ns * ns * 2.0f;
      const float r0 = (g.x * hh.x - g.y * hh.y) / den, i0 = (g.x * hh.y + g.y * hh.x) / den;
      const float r1 = (g.z * hh.x - g.w * hh.y) / den, i1 = (g.z * hh.y + g.w * hh.x) / den;
      g = make_float4(r0, i0, r1, i1);
    }
    // y = x * h  =>  dx = dy * conj(h)
    if (dx != nullptr)
      dx[i] = make_float4(g.x * hh.x + g.y * hh.y, g.y * hh.x - g.x * hh.y, g.z * hh.x + g.w * hh.y, g.w * hh.x - g.z * hh.y);
    if (dp != nullptr) dp[i] = make_float4(0.f, 0.f, 0.f, 0.f);     // p is ignored by the fading channel (:35-83)
  }
}

// ================================================================== dropout (forward and backward are the same map)
// out = keep ? x / (1 - rate) : 0, keep decided by Philox4x32-10(seed, offset, element group): the mask is
// regenerated in the backward pass instead of being stored.
__device__ __forceinline__ uint4 philox4(uint64_t ctr_lo, uint64_t ctr_hi, uint64_t seed) {
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
__global__ void __launch_bounds__(256)
dropout_kernel(const float4* __restrict__ x, float4* __restrict__ out, float rate, uint64_t seed, uint64_t offset,
               const int64_t* __restrict__ step_dev, int64_t n4) {
  if (step_dev != nullptr) offset += (uint64_t)(*step_dev) << 20;          // a fresh mask per (graph-replayed) step
  const float keep_scale = 1.0f / (1.0f - rate);
  const uint32_t thresh = (uint32_t)(rate * 4294967296.0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 r = philox4((uint64_t)i, offset, seed);
    const float4 v = ld_stream(x + i);
    out[i] = make_float4(r.x >= thresh ? v.x * keep_scale : 0.f, r.y >= thresh ? v.y * keep_scale : 0.f,
                         r.z >= thresh ? v.z * keep_scale : 0.f, r.w >= thresh ? v.w * keep_scale : 0.f);
  }
}

// ================================================================== Adam (tf.keras.optimizers.Adam update rule)
// g = grad*grad_scale (+ grad2*grad2_scale); m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// p -= lr * sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps)
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, const float* __restrict__ g2, float* __restrict__ m,
            float* __restrict__ v, float lr_t, float lr, int step, const int64_t* __restrict__ step_dev, int steps_per_iter,
            float b1, float b2, float eps, float grad_scale, float grad2_scale, int64_t n) {
  if (step_dev != nullptr) {             // graph replay: the iteration count lives on the device
    const double t = (double)step + (double)(*step_dev) * (double)steps_per_iter;
    lr_t = (float)((double)lr * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t)));
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale;
    if (g2 != nullptr) gi = fmaf(g2[i], grad2_scale, gi);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

static int stream_blocks(int64_t n, int per_block = 256) {
  int64_t want = (n + per_block - 1) / per_block;
  return (int)(want < (int64_t)kSMs * 8 ? (want > 0 ? want : 1) : (int64_t)kSMs * 8);
}

}  // namespace dsc

using namespace dsc;

namespace dsc {
bool gemm_nt_tc_eligible(const float* A, int64_t lda, const float* B, int64_t ldb, int M, int N, int K);   // dsc_gemm_nt_tc.cu
int gemm_nt_tc(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int M, int N, int K,
               int accumulate, cudaStream_t s);
}

extern "C" int dsc_gemm(const float* A, int64_t lda, int trans_a, const float* B, int64_t ldb, int trans_b,
                        float* C, int64_t ldc, int M, int N, int K, int accumulate, void* stream) {
  DSC_REQUIRE(A && B && C, "dsc_gemm: null pointer");
  DSC_REQUIRE(M >= 0 && N >= 0 && K >= 0 && ldc >= N, "dsc_gemm: bad shape");
  DSC_REQUIRE(lda >= (trans_a ? M : K) && ldb >= (trans_b ? K : N), "dsc_gemm: leading dimension too small");
  if (M == 0 || N == 0) return DSC_OK;
  cudaStream_t s = as_stream(stream);
  // large A * B^T products (both operands contiguous along K) go to the tensor cores in bf16x3
  if (!trans_a && trans_b && gemm_nt_tc_eligible(A, lda, B, ldb, M, N, K))
    return gemm_nt_tc(A, lda, B, ldb, C, ldc, M, N, K, accumulate, s);
  const int tm = (M + 63) / 64, tn = (N + 63) / 64;
  int splits = 1;
  if (K >= 512 && tm * tn < 2 * kSMs) {
    splits = (2 * kSMs + tm * tn - 1) / (tm * tn);
    const int max_splits = K / 128;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  int kps = ((K + splits - 1) / splits + 15) / 16 * 16;
  if (kps < 16) kps = 16;
  splits = K > 0 ? (K + kps - 1) / kps : 1;
  const int atomic = (splits > 1 || accumulate) ? 1 : 0;
  if (splits > 1 && !accumulate) {
    zero_rows_kernel<<<stream_blocks((int64_t)M * N), 256, 0, s>>>(C, ldc, M, N);
  }
  if (K == 0) {
    if (!accumulate) zero_rows_kernel<<<stream_blocks((int64_t)M * N), 256, 0, s>>>(C, ldc, M, N);
    return check_launch("dsc_gemm");
  }
  dim3 grid(tn, tm, splits);
  if (trans_a && trans_b) gemm_nt_kernel<true, true><<<grid, 256, 0, s>>>(A, lda, B, ldb, C, ldc, M, N, K, kps, atomic);
  else if (trans_a) gemm_nt_kernel<true, false><<<grid, 256, 0, s>>>(A, lda, B, ldb, C, ldc, M, N, K, kps, atomic);
  else if (trans_b) gemm_nt_kernel<false, true><<<grid, 256, 0, s>>>(A, lda, B, ldb, C, ldc, M, N, K, kps, atomic);
  else gemm_nt_kernel<false, false><<<grid, 256, 0, s>>>(A, lda, B, ldb, C, ldc, M, N, K, kps, atomic);
  return check_launch("dsc_gemm");
}

extern "C" int dsc_bias_act_backward(const float* dy, int64_t ld_dy, const float* y, int64_t ld_y, int act,
                                     float* dz, int64_t ld_dz, float* dbias, int M, int N, void* stream) {
  DSC_REQUIRE(dy && M >= 0 && N > 0 && (act == 0 || act == 1), "dsc_bias_act_backward: bad argument");
  DSC_REQUIRE(act == 0 || (y && dz), "dsc_bias_act_backward: relu needs y and dz");
  if (M == 0) return DSC_OK;
  cudaStream_t s = as_stream(stream);
  if (dbias) {
    cudaError_t e = cudaMemsetAsync(dbias, 0, sizeof(float) * N, s);
    if (e != cudaSuccess) { set_error("dsc_bias_act_backward: %s", cudaGetErrorString(e)); return DSC_ERR_CUDA; }
  }
  const int cb = (N + 31) / 32;
  int rb = (M + 63) / 64;
  const int cap = (4 * kSMs + cb - 1) / cb;
  if (rb > cap) rb = cap;
  if (rb < 1) rb = 1;
  bias_act_backward_kernel<<<dim3(cb, rb), 256, 0, s>>>(dy, ld_dy, y, ld_y, act, act ? dz : nullptr, ld_dz, dbias, M, N);
  return check_launch("dsc_bias_act_backward");
}

extern "C" int dsc_add_layernorm_backward(const float* x, int64_t xgs, const float* res, int64_t rgs,
                                          const float* gamma_a, const float* beta_a, const float* gamma_b, const float* beta_b,
                                          const float* dout, int64_t dogs, float* dv, int64_t dvgs,
                                          float* dgamma_a, float* dbeta_a, float* dgamma_b, float* dbeta_b,
                                          int n_rows, int group_rows, void* stream) {
  DSC_REQUIRE(x && gamma_a && beta_a && dout && dv, "dsc_add_layernorm_backward: null pointer");
  DSC_REQUIRE((gamma_b == nullptr) == (beta_b == nullptr), "dsc_add_layernorm_backward: gamma_b/beta_b must come together");
  DSC_REQUIRE(group_rows > 0 && n_rows >= 0, "dsc_add_layernorm_backward: bad sizes");
  DSC_REQUIRE(((xgs | rgs | dogs | dvgs) & 3) == 0 && aligned16(x) && aligned16(dout) && aligned16(dv) && (!res || aligned16(res)),
              "dsc_add_layernorm_backward: rows must be 16-byte aligned");
  if (n_rows == 0) return DSC_OK;
  int blocks = min((n_rows + 7) / 8, kSMs * 4);
  add_layernorm_backward_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, xgs, res, rgs, gamma_a, beta_a, gamma_b, beta_b, dout, dogs,
                                                                        dv, dvgs, dgamma_a, dbeta_a, dgamma_b, dbeta_b, n_rows, group_rows);
  return check_launch("dsc_add_layernorm_backward");
}

extern "C" int dsc_mha_attention_backward(const float* q, int64_t ldq, int64_t qbs, const float* k, const float* v,
                                          int64_t ldkv, int64_t kvbs, const float* dout, int64_t ldo, int64_t obs,
                                          const float* mask, int64_t mbs, int64_t mqs,
                                          const int32_t* key_ids, int64_t kis, int causal, int q_off,
                                          float* dq, int64_t lddq, int64_t dqbs, float* dk, float* dv, int64_t lddkv, int64_t dkvbs,
                                          int n, int lq, int lk, void* stream) {
  DSC_REQUIRE(q && k && v && dout && dq && dk && dv, "dsc_mha_attention_backward: null pointer");
  DSC_REQUIRE(n >= 0 && lq > 0 && lq <= 64 && lk > 0 && lk <= 64, "dsc_mha_attention_backward: lq, lk must be in 1..64");
  DSC_REQUIRE(((ldq | qbs | ldkv | kvbs | ldo | obs | lddq | dqbs | lddkv | dkvbs) & 3) == 0 && aligned16(q) && aligned16(k) &&
              aligned16(v) && aligned16(dout) && aligned16(dq) && aligned16(dk) && aligned16(dv),
              "dsc_mha_attention_backward: rows must be 16-byte aligned");
  if (n == 0) return DSC_OK;
  static bool attr_set = false;
  const size_t smem = ((size_t)(2 * lk + 2 * lq) * 128 + 3 * 8 * 64) * sizeof(float);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mha_attention_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(((size_t)4 * 64 * 128 + 3 * 8 * 64) * sizeof(float)));
    if (e != cudaSuccess) { set_error("dsc_mha_attention_backward: %s", cudaGetErrorString(e)); return DSC_ERR_CUDA; }
    attr_set = true;
  }
  mha_attention_backward_kernel<<<n, 256, smem, as_stream(stream)>>>(q, ldq, qbs, k, v, ldkv, kvbs, dout, ldo, obs, mask, mbs, mqs,
                                                                     key_ids, kis, causal, q_off, dq, lddq, dqbs, dk, dv, lddkv, dkvbs, lq, lk);
  return check_launch("dsc_mha_attention_backward");
}

extern "C" int dsc_star_satellite_attn_backward(const float* qkv, const float* kv_e, const float* datt,
                                                float* dqkv, float* dkv_e, int n_sent, void* stream) {
  DSC_REQUIRE(qkv && kv_e && datt && dqkv && dkv_e && n_sent >= 0, "dsc_star_satellite_attn_backward: bad argument");
  DSC_REQUIRE(aligned16(qkv) && aligned16(kv_e) && aligned16(datt) && aligned16(dqkv) && aligned16(dkv_e),
              "dsc_star_satellite_attn_backward: misaligned pointer");
  if (n_sent == 0) return DSC_OK;
  star_satellite_attn_backward_kernel<<<n_sent, 256, 0, as_stream(stream)>>>(qkv, kv_e, datt, dqkv, dkv_e);
  return check_launch("dsc_star_satellite_attn_backward");
}

extern "C" int dsc_star_relay_attn_backward(const float* qkv_r, const float* kv2, int kv2_rows, int n2, const float* dout,
                                            float* dqkv_r, float* dkv2, int n_sent, void* stream) {
  DSC_REQUIRE(qkv_r && dout && dqkv_r && n_sent >= 0, "dsc_star_relay_attn_backward: bad argument");
  DSC_REQUIRE(n2 >= 0 && n2 <= 32 && kv2_rows >= 0 && kv2_rows <= 32 && (n2 == 0 || (kv2 && dkv2 && n2 <= kv2_rows)),
              "dsc_star_relay_attn_backward: bad h2 key count");
  DSC_REQUIRE(aligned16(qkv_r) && aligned16(dout) && aligned16(dqkv_r) && (!kv2 || aligned16(kv2)) && (!dkv2 || aligned16(dkv2)),
              "dsc_star_relay_attn_backward: misaligned pointer");
  if (n_sent == 0) return DSC_OK;
  star_relay_attn_backward_kernel<<<n_sent, 256, 0, as_stream(stream)>>>(qkv_r, n2 ? kv2 : nullptr, kv2_rows, n2, dout, dqkv_r,
                                                                         n2 ? dkv2 : nullptr);
  return check_launch("dsc_star_relay_attn_backward");
}

extern "C" int dsc_embed_backward(const int32_t* ids, int64_t ids_stride, const float* dout, int64_t ld_dout,
                                  float* dtable, int vocab, int n_sent, int len, void* stream) {
  DSC_REQUIRE(ids && dout && dtable && n_sent >= 0 && len > 0 && vocab > 0, "dsc_embed_backward: bad argument");
  DSC_REQUIRE((ld_dout & 3) == 0 && aligned16(dout), "dsc_embed_backward: rows must be 16-byte aligned");
  const int n_rows = n_sent * len;
  if (n_rows == 0) return DSC_OK;
  embed_backward_kernel<<<min((n_rows + 7) / 8, kSMs * 8), 256, 0, as_stream(stream)>>>(ids, ids_stride, dout, ld_dout, dtable, vocab,
                                                                                         n_rows, len);
  return check_launch("dsc_embed_backward");
}

extern "C" int dsc_star_pack_backward(const float* dtile, float* dsrc, int n_sent, void* stream) {
  DSC_REQUIRE(dtile && dsrc && n_sent >= 0, "dsc_star_pack_backward: bad argument");
  if (n_sent == 0) return DSC_OK;
  star_pack_backward_kernel<<<n_sent, 128, 0, as_stream(stream)>>>(dtile, dsrc);
  return check_launch("dsc_star_pack_backward");
}

extern "C" int dsc_masked_ce_backward(const float* logits, int64_t ld, const int32_t* target, const float* grad_rows,
                                      float* dlogits, int64_t ld_d, int M, int N, void* stream) {
  DSC_REQUIRE(logits && target && grad_rows && dlogits && M >= 0 && N > 0, "dsc_masked_ce_backward: bad argument");
  if (M == 0) return DSC_OK;
  const size_t smem = sizeof(float) * (size_t)N;
  DSC_REQUIRE(smem <= 200 * 1024, "dsc_masked_ce_backward: a row of %d logits does not fit in shared memory", N);
  static size_t attr_bytes = 0;
  if (smem > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(masked_ce_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("dsc_masked_ce_backward: %s", cudaGetErrorString(e)); return DSC_ERR_CUDA; }
    attr_bytes = smem;
  }
  masked_ce_backward_kernel<<<M, 256, smem, as_stream(stream)>>>(logits, ld, target, grad_rows, dlogits, ld_d, N);
  return check_launch("dsc_masked_ce_backward");
}

extern "C" int dsc_unit_dot(const float* a, const float* b, float* out, int n_units, int64_t elems_per_unit, void* stream) {
  DSC_REQUIRE(a && b && out && n_units >= 0 && elems_per_unit > 0 && (elems_per_unit & 3) == 0 && aligned16(a) && aligned16(b),
              "dsc_unit_dot: bad argument");
  if (n_units == 0) return DSC_OK;
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * n_units, as_stream(stream));
  if (e != cudaSuccess) { set_error("dsc_unit_dot: %s", cudaGetErrorString(e)); return DSC_ERR_CUDA; }
  int64_t want = (elems_per_unit / 4 + 255) / 256;
  int chunks = (int)(want < 32 ? want : 32);
  unit_dot_kernel<<<dim3(chunks, n_units), 256, 0, as_stream(stream)>>>(a, b, out, elems_per_unit);
  return check_launch("dsc_unit_dot");
}

extern "C" int dsc_power_normalize_backward(const float* x, const float* sumsq, const float* dot, float factor,
                                            const float* dy, float* dx, int n_units, int64_t elems_per_unit, void* stream) {
  DSC_REQUIRE(x && sumsq && dot && dy && dx, "dsc_power_normalize_backward: null pointer");
  DSC_REQUIRE(n_units >= 0 && elems_per_unit > 0 && (elems_per_unit & 3) == 0 && aligned16(x) && aligned16(dy) && aligned16(dx),
              "dsc_power_normalize_backward: bad sizes or alignment");
  if (n_units == 0) return DSC_OK;
  const int64_t unit4 = elems_per_unit / 4, n4 = unit4 * n_units;
  power_normalize_backward_kernel<<<stream_blocks(n4), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(x), sumsq, dot, factor, reinterpret_cast<const float4*>(dy), reinterpret_cast<float4*>(dx), n4, unit4);
  return check_launch("dsc_power_normalize_backward");
}

extern "C" int dsc_channel_backward(const float* dy, const float* h, const float* n_std, int detector, const float* p_scale,
                                    float* dx, float* dp, int n_units, int64_t elems_per_unit, void* stream) {
  DSC_REQUIRE(dy && n_std && (dx || dp), "dsc_channel_backward: null pointer");
  DSC_REQUIRE(n_units >= 0 && elems_per_unit > 0 && (elems_per_unit & 3) == 0 && aligned16(dy) && (!dx || aligned16(dx)) && (!dp || aligned16(dp)),
              "dsc_channel_backward: bad sizes or alignment");
  if (detector < 0 || detector > 2) { set_error("detector must in LS and MMSE"); return DSC_ERR_BAD_ARG; }
  if (n_units == 0) return DSC_OK;
  const int64_t unit4 = elems_per_unit / 4, n4 = unit4 * n_units;
  channel_backward_kernel<<<stream_blocks(n4), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(dy), reinterpret_cast<const float2*>(h), n_std, detector, p_scale,
      reinterpret_cast<float4*>(dx), reinterpret_cast<float4*>(dp), n4, unit4);
  return check_launch("dsc_channel_backward");
}

extern "C" int dsc_dropout(const float* x, float* out, float rate, uint64_t seed, uint64_t offset, const int64_t* step_dev,
                           int64_t n, void* stream) {
  DSC_REQUIRE(x && out && n >= 0 && (n & 3) == 0 && aligned16(x) && aligned16(out), "dsc_dropout: n must be a multiple of 4, 16-byte aligned tensors");
  DSC_REQUIRE(rate >= 0.f && rate < 1.f, "dsc_dropout: rate must be in [0, 1)");
  if (n == 0) return DSC_OK;
  dropout_kernel<<<stream_blocks(n / 4), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(out),
                                                                    rate, seed, offset, step_dev, n / 4);
  return check_launch("dsc_dropout");
}

extern "C" int dsc_adam_step(float* param, const float* grad, const float* grad2, float* m, float* v, float lr, float beta1,
                             float beta2, float eps, int step, const int64_t* step_dev, int steps_per_iter,
                             float grad_scale, float grad2_scale, int64_t n, void* stream) {
  DSC_REQUIRE(param && grad && m && v && n >= 0 && step >= 1, "dsc_adam_step: bad argument");
  if (n == 0) return DSC_OK;
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, step)) / (1.0 - pow((double)beta1, step));
  adam_kernel<<<stream_blocks(n), 256, 0, as_stream(stream)>>>(param, grad, grad2, m, v, (float)lr_t, lr, step, step_dev, steps_per_iter,
                                                                beta1, beta2, eps, grad_scale, grad2_scale, n);
  return check_launch("dsc_adam_step");
}
