// Attention kernels for d_model = 128, 8 heads of depth 16.
//  * star satellite attention: warp per token, 5 keys gathered by index (cyclic neighbours, the
//    token's own embedding, the relay), fused softmax(5) + PV.  HBM/L2-bound: 11 rows of 512 B per token.
//  * star relay attention: CTA per sentence, warp per head, <= 62 keys.
//  * generic small-L multi-head attention (baseline Transformer, multi_tar): CTA per sentence,
//    K/V of the sentence staged in shared memory, warp per head, lane per query.
#include "dsc_common.cuh"

namespace dsc {

// ---------------------------------------------------------------- K3 satellite attention
// lane l owns columns 4l..4l+3, i.e. head l/4; a head's dot product is reduced over its 4 lanes.
__device__ __forceinline__ float head_dot(const float4& a, const float4& b) {
  float d = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  d += __shfl_xor_sync(0xffffffffu, d, 2);
  return d;
}

__global__ void __launch_bounds__(256)
star_satellite_attn_kernel(const float* __restrict__ qkv, const float* __restrict__ kv_e,
                           float* __restrict__ att, int n_sent) {
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  const int n_rows = n_sent * DSC_TILE_ROWS;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps_per_grid) {
    const int i = r & 31;
    float4* dst = reinterpret_cast<float4*>(att + (int64_t)r * 128) + lane;
    if (i == DSC_SEQ) { *dst = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
    const int base = r - i;
    const int i_up = (i + 1 == DSC_SEQ) ? 0 : i + 1;          // roll(h,-1)[i] = h[(i+1) mod L]
    const int i_dn = (i == 0) ? DSC_SEQ - 1 : i - 1;          // roll(h,+1)[i] = h[(i-1) mod L]
    const float4* row_i = reinterpret_cast<const float4*>(qkv + (int64_t)r * 384);
    const float4* row_u = reinterpret_cast<const float4*>(qkv + (int64_t)(base + i_up) * 384);
    const float4* row_d = reinterpret_cast<const float4*>(qkv + (int64_t)(base + i_dn) * 384);
    const float4* row_s = reinterpret_cast<const float4*>(qkv + (int64_t)(base + DSC_SEQ) * 384);
    const float4* row_e = reinterpret_cast<const float4*>(kv_e + (int64_t)r * 256);
    const float4 q = __ldg(row_i + lane);
    float4 k[5], v[5];
    k[0] = __ldg(row_u + 32 + lane); v[0] = __ldg(row_u + 64 + lane);
    k[1] = __ldg(row_i + 32 + lane); v[1] = __ldg(row_i + 64 + lane);
    k[2] = __ldg(row_d + 32 + lane); v[2] = __ldg(row_d + 64 + lane);
    k[3] = __ldg(row_e + lane);      v[3] = __ldg(row_e + 32 + lane);
    k[4] = __ldg(row_s + 32 + lane); v[4] = __ldg(row_s + 64 + lane);
    float l[5], mx = -3.4e38f;
#pragma unroll
    for (int j = 0; j < 5; ++j) { l[j] = head_dot(q, k[j]) * 0.25f; mx = fmaxf(mx, l[j]); }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 5; ++j) { l[j] = expf(l[j] - mx); sum += l[j]; }
    const float inv = 1.0f / sum;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      float w = l[j] * inv;
      o.x = fmaf(w, v[j].x, o.x); o.y = fmaf(w, v[j].y, o.y);
      o.z = fmaf(w, v[j].z, o.z); o.w = fmaf(w, v[j].w, o.w);
    }
    *dst = o;
  }
}

// ---------------------------------------------------------------- K4 relay attention
__global__ void __launch_bounds__(256)
star_relay_attn_kernel(const float* __restrict__ qkv_r, const float* __restrict__ kv2, int kv2_rows, int n2,
                       float* __restrict__ out, int n_sent) {
  __shared__ float w_s[8][64];
  const int s = blockIdx.x;
  if (s >= n_sent) return;
  const int head = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* tile = qkv_r + (int64_t)s * DSC_TILE_ROWS * 384;
  const float* k2 = kv2 ? kv2 + (int64_t)s * kv2_rows * 256 : nullptr;
  const int nk = DSC_TILE_ROWS + n2;          // key j < 32: tile row (j==0 -> relay row 31, else satellite j-1)
  float qh[16];
  {
    const float4* qp = reinterpret_cast<const float4*>(tile + DSC_SEQ * 384 + head * 16);
#pragma unroll
    for (int t = 0; t < 4; ++t) { float4 a = __ldg(qp + t); qh[4*t] = a.x; qh[4*t+1] = a.y; qh[4*t+2] = a.z; qh[4*t+3] = a.w; }
  }
  float lg[2];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    int j = lane + 32 * half;
    float d = -3.4e38f;
    if (j < nk) {
      const float* kr = (j < 32) ? tile + (int64_t)((j == 0) ? DSC_SEQ : j - 1) * 384 + 128 + head * 16
                                 : k2 + (int64_t)(j - 32) * 256 + head * 16;
      const float4* kp = reinterpret_cast<const float4*>(kr);
      d = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float4 a = __ldg(kp + t);
        d = fmaf(qh[4*t], a.x, d); d = fmaf(qh[4*t+1], a.y, d); d = fmaf(qh[4*t+2], a.z, d); d = fmaf(qh[4*t+3], a.w, d);
      }
      d *= 0.25f;
    }
    lg[half] = d;
  }
  float mx = warp_max(fmaxf(lg[0], lg[1]));
  float e0 = (lane < nk) ? expf(lg[0] - mx) : 0.f;
  float e1 = (lane + 32 < nk) ? expf(lg[1] - mx) : 0.f;
  float inv = 1.0f / warp_sum(e0 + e1);
  w_s[head][lane] = e0 * inv;
  w_s[head][lane + 32] = e1 * inv;
  __syncwarp();
  // PV: lane = (half, d): d = lane & 15 owns output column head*16+d, half splits the keys by parity
  const int d = lane & 15, half = lane >> 4;
  float acc = 0.f;
  for (int j = half; j < nk; j += 2) {
    const float* vr = (j < 32) ? tile + (int64_t)((j == 0) ? DSC_SEQ : j - 1) * 384 + 256 + head * 16
                               : k2 + (int64_t)(j - 32) * 256 + 128 + head * 16;
    acc = fmaf(w_s[head][j], __ldg(vr + d), acc);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 16);
  if (half == 0) out[(int64_t)s * 128 + head * 16 + d] = acc;
}

// ---------------------------------------------------------------- K5 generic small-L MHA
// dynamic smem: K [lk][128], V [lk][128]
__global__ void __launch_bounds__(256)
mha_attention_kernel(const float* __restrict__ q, int64_t ldq, int64_t qbs,
                     const float* __restrict__ k, const float* __restrict__ v, int64_t ldkv, int64_t kvbs,
                     float* __restrict__ out, int64_t ldo, int64_t obs,
                     const float* __restrict__ mask, int64_t mbs, int64_t mqs,
                     const int32_t* __restrict__ key_ids, int64_t kis, int causal, int q_off,
                     int lq, int lk) {
  extern __shared__ __align__(16) float smem[];
  float* Ks = smem;
  float* Vs = smem + (size_t)lk * 128;
  __shared__ float padm[64];
  const int b = blockIdx.x;
  const float* kb = k + (int64_t)b * kvbs;
  const float* vb = v + (int64_t)b * kvbs;
  for (int idx = threadIdx.x; idx < lk * 32; idx += blockDim.x) {
    int j = idx >> 5, c = idx & 31;
    reinterpret_cast<float4*>(Ks)[idx] = __ldg(reinterpret_cast<const float4*>(kb + (int64_t)j * ldkv) + c);
    reinterpret_cast<float4*>(Vs)[idx] = __ldg(reinterpret_cast<const float4*>(vb + (int64_t)j * ldkv) + c);
  }
  if (threadIdx.x < 64)
    padm[threadIdx.x] = (key_ids && threadIdx.x < lk && key_ids[(int64_t)b * kis + threadIdx.x] == 0) ? 1.f : 0.f;
  __syncthreads();
  const int head = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < lq; i += 32) {
    float qh[16];
    const float4* qp = reinterpret_cast<const float4*>(q + (int64_t)b * qbs + (int64_t)i * ldq + head * 16);
#pragma unroll
    for (int t = 0; t < 4; ++t) { float4 a = __ldg(qp + t); qh[4*t] = a.x; qh[4*t+1] = a.y; qh[4*t+2] = a.z; qh[4*t+3] = a.w; }
    const float* mrow = mask ? mask + (int64_t)b * mbs + (int64_t)i * mqs : nullptr;
    auto logit = [&](int j) {
      const float4* kp = reinterpret_cast<const float4*>(Ks + j * 128 + head * 16);
      float d = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float4 a = kp[t];
        d = fmaf(qh[4*t], a.x, d); d = fmaf(qh[4*t+1], a.y, d); d = fmaf(qh[4*t+2], a.z, d); d = fmaf(qh[4*t+3], a.w, d);
      }
      d *= 0.25f;
      float m = padm[j];
      if (mrow) m = fmaxf(m, __ldg(mrow + j));
      if (causal && j > q_off + i) m = 1.f;
      return d + m * -1e9f;
    };
    float mx = -3.4e38f;
    for (int j = 0; j < lk; ++j) mx = fmaxf(mx, logit(j));
    float sum = 0.f, acc[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) acc[t] = 0.f;
    for (int j = 0; j < lk; ++j) {
      float w = expf(logit(j) - mx);
      sum += w;
      const float4* vp = reinterpret_cast<const float4*>(Vs + j * 128 + head * 16);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float4 a = vp[t];
        acc[4*t] = fmaf(w, a.x, acc[4*t]); acc[4*t+1] = fmaf(w, a.y, acc[4*t+1]);
        acc[4*t+2] = fmaf(w, a.z, acc[4*t+2]); acc[4*t+3] = fmaf(w, a.w, acc[4*t+3]);
      }
    }
    const float inv = 1.0f / sum;
    float4* op = reinterpret_cast<float4*>(out + (int64_t)b * obs + (int64_t)i * ldo + head * 16);
#pragma unroll
    for (int t = 0; t < 4; ++t)
      op[t] = make_float4(acc[4*t] * inv, acc[4*t+1] * inv, acc[4*t+2] * inv, acc[4*t+3] * inv);
  }
}

// ---------------------------------------------------------------- K5, single query (greedy decode step)
// One warp per (sentence, head): lane = key (two keys per lane when lk > 32).  No shared-memory staging: each key /
// value row slice (16 floats) is read once.  Same masking and arithmetic as mha_attention_kernel.
__global__ void __launch_bounds__(256)
mha_decode_attention_kernel(const float* __restrict__ q, int64_t qbs, const float* __restrict__ k, const float* __restrict__ v,
                            int64_t ldkv, int64_t kvbs, float* __restrict__ out, int64_t obs,
                            const float* __restrict__ mask, int64_t mbs,
                            const int32_t* __restrict__ key_ids, int64_t kis, int causal, int q_off, int n, int lk) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= n * 8) return;
  const int b = gw >> 3, head = gw & 7;
  float qh[16];
  {
    const float4* qp = reinterpret_cast<const float4*>(q + (int64_t)b * qbs + head * 16);
#pragma unroll
    for (int t = 0; t < 4; ++t) { const float4 a = __ldg(qp + t); qh[4*t] = a.x; qh[4*t+1] = a.y; qh[4*t+2] = a.z; qh[4*t+3] = a.w; }
  }
  float lg[2], w[2];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int j = lane + 32 * half;
    float d = -3.4e38f;
    if (j < lk) {
      const float4* kp = reinterpret_cast<const float4*>(k + (int64_t)b * kvbs + (int64_t)j * ldkv + head * 16);
      d = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float4 a = __ldg(kp + t);
        d = fmaf(qh[4*t], a.x, d); d = fmaf(qh[4*t+1], a.y, d); d = fmaf(qh[4*t+2], a.z, d); d = fmaf(qh[4*t+3], a.w, d);
      }
      d *= 0.25f;
      float m = (key_ids != nullptr && key_ids[(int64_t)b * kis + j] == 0) ? 1.f : 0.f;
      if (mask != nullptr) m = fmaxf(m, __ldg(mask + (int64_t)b * mbs + j));
      if (causal && j > q_off) m = 1.f;
      d += m * -1e9f;
    }
    lg[half] = d;
  }
  const float mx = warp_max(fmaxf(lg[0], lg[1]));
  w[0] = (lane < lk) ? expf(lg[0] - mx) : 0.f;
  w[1] = (lane + 32 < lk) ? expf(lg[1] - mx) : 0.f;
  const float inv = 1.0f / warp_sum(w[0] + w[1]);
  float p[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) p[t] = 0.f;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int j = lane + 32 * half;
    if (j < lk) {
      const float4* vp = reinterpret_cast<const float4*>(v + (int64_t)b * kvbs + (int64_t)j * ldkv + head * 16);
      const float ww = w[half] * inv;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float4 a = __ldg(vp + t);
        p[4*t] = fmaf(ww, a.x, p[4*t]); p[4*t+1] = fmaf(ww, a.y, p[4*t+1]);
        p[4*t+2] = fmaf(ww, a.z, p[4*t+2]); p[4*t+3] = fmaf(ww, a.w, p[4*t+3]);
      }
    }
  }
  // sum over the 32 lanes: fold the half-warps, then reduce-scatter 16 values over 16 lanes (lane l ends with dim l & 15)
#pragma unroll
  for (int i = 0; i < 16; ++i) p[i] += __shfl_xor_sync(0xffffffffu, p[i], 16);
#pragma unroll
  for (int off = 8, m = 8; off >= 1; off >>= 1, m >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < m; ++i) {
      const float send = upper ? p[i] : p[i + m];
      const float keep = upper ? p[i + m] : p[i];
      p[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  if (lane < 16) out[(int64_t)b * obs + head * 16 + lane] = p[0];
}

}  // namespace dsc

using namespace dsc;

extern "C" int dsc_star_satellite_attn(const float* qkv, const float* kv_e, float* att, int n_sent, void* stream) {
  DSC_REQUIRE(qkv && kv_e && att && n_sent >= 0, "dsc_star_satellite_attn: bad argument");
  DSC_REQUIRE(aligned16(qkv) && aligned16(kv_e) && aligned16(att), "dsc_star_satellite_attn: misaligned pointer");
  if (n_sent == 0) return DSC_OK;
  int n_rows = n_sent * DSC_TILE_ROWS;
  int blocks = min((n_rows + 7) / 8, kSMs * 16);
  star_satellite_attn_kernel<<<blocks, 256, 0, as_stream(stream)>>>(qkv, kv_e, att, n_sent);
  return check_launch("dsc_star_satellite_attn");
}

extern "C" int dsc_star_relay_attn(const float* qkv_r, const float* kv2, int kv2_rows, int n2,
                                   float* out, int n_sent, void* stream) {
  DSC_REQUIRE(qkv_r && out && n_sent >= 0, "dsc_star_relay_attn: bad argument");
  DSC_REQUIRE(n2 >= 0 && n2 <= 32 && (n2 == 0 || (kv2 && n2 <= kv2_rows)), "dsc_star_relay_attn: bad h2 key count");
  DSC_REQUIRE(aligned16(qkv_r) && (!kv2 || aligned16(kv2)), "dsc_star_relay_attn: misaligned pointer");
  if (n_sent == 0) return DSC_OK;
  star_relay_attn_kernel<<<n_sent, 256, 0, as_stream(stream)>>>(qkv_r, n2 ? kv2 : nullptr, kv2_rows, n2, out, n_sent);
  return check_launch("dsc_star_relay_attn");
}

extern "C" int dsc_mha_attention(const float* q, int64_t ldq, int64_t qbs, const float* k, const float* v,
                                 int64_t ldkv, int64_t kvbs, float* out, int64_t ldo, int64_t obs,
                                 const float* mask, int64_t mbs, int64_t mqs,
                                 const int32_t* key_ids, int64_t kis, int causal, int q_off,
                                 int n, int lq, int lk, void* stream) {
  DSC_REQUIRE(q && k && v && out, "dsc_mha_attention: null pointer");
  DSC_REQUIRE(n >= 0 && lq > 0 && lq <= 64 && lk > 0 && lk <= 64, "dsc_mha_attention: lq, lk must be in 1..64");
  DSC_REQUIRE(((ldq | qbs | ldkv | kvbs | ldo | obs) & 3) == 0 && aligned16(q) && aligned16(k) && aligned16(v) && aligned16(out),
              "dsc_mha_attention: rows must be 16-byte aligned");
  if (n == 0) return DSC_OK;
  if (lq == 1) {                       // greedy decode step: one query per sentence, warp per (sentence, head)
    mha_decode_attention_kernel<<<(n * 8 * 32 + 255) / 256, 256, 0, as_stream(stream)>>>(
        q, qbs, k, v, ldkv, kvbs, out, obs, mask, mbs, key_ids, kis, causal, q_off, n, lk);
    return check_launch("dsc_mha_attention");
  }
  static bool attr_set = false;
  size_t smem = (size_t)lk * 128 * 2 * sizeof(float);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mha_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 128 * 2 * 4);
    if (e != cudaSuccess) { set_error("dsc_mha_attention: %s", cudaGetErrorString(e)); return DSC_ERR_CUDA; }
    attr_set = true;
  }
  mha_attention_kernel<<<n, 256, smem, as_stream(stream)>>>(q, ldq, qbs, k, v, ldkv, kvbs, out, ldo, obs,
                                                            mask, mbs, mqs, key_ids, kis, causal, q_off, lq, lk);
  return check_launch("dsc_mha_attention");
}
