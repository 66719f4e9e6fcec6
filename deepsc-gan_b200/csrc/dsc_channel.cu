// K9: the fused channel kernel.  Power-normalise + noise (injected or Philox) + perturbation +
// fading (complex multiply) + equalise in one pass.  Pure HBM stream: each thread moves one float4
// (= two complex symbols) per tensor; algorithmic bytes per element are 4 (x) + 4 (y) [+4 noise]
// [+4 perturbation] [+4 x_norm].  Grid sized in multiples of the SM count.
#include "dsc_common.cuh"

namespace dsc {

struct Philox {
  // Philox4x32-10 (Salmon et al., SC'11): counter = (group index, stream offset), key = seed.
  static __device__ __forceinline__ uint4 gen(uint64_t ctr_lo, uint64_t ctr_hi, uint64_t seed) {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
  // Box-Muller on the SFU: lg2 / rsqrt / sin / cos approximations (absolute error < 4e-6 on a unit normal, far inside the
  // 1e-3 budget of the channel symbols); the exact libm forms made the kernel instruction-bound at 41 % of HBM.
  static __device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
    const float u1 = ((float)a + 1.0f) * 2.3283064365386963e-10f;     // (0, 1]
    const float t = (float)b * 2.3283064365386963e-10f - 0.5f;        // [-0.5, 0.5]: angle 2*pi*t + pi
    const float r = sqrtf(-1.3862943611198906f * __log2f(u1));        // sqrt(-2 ln u1)
    float s, c;
    __sincosf(6.283185307179586f * t, &s, &c);
    return make_float2(-r * c, -r * s);                                 // cos(a + pi) = -cos a, sin(a + pi) = -sin a
  }
};

__global__ void __launch_bounds__(256)
channel_kernel(const float4* __restrict__ x, const float* __restrict__ x_sumsq, float x_factor,
               const float4* __restrict__ noise, uint64_t seed, uint64_t offset,
               const float4* __restrict__ p, const float* __restrict__ p_sumsq, float p_factor,
               const float* __restrict__ p_scale, const float2* __restrict__ h,
               const float* __restrict__ n_std, int detector,
               float4* __restrict__ y, float4* __restrict__ x_norm, int n_units, int unit4) {
  // grid (chunks, units): the per-unit scalars are loaded once per CTA and unit, no index division in the stream loop
  for (int u = blockIdx.y; u < n_units; u += gridDim.y) {
    const float elems = (float)unit4 * 4.0f;
    const float sc = x_sumsq ? 1.0f / sqrtf(x_factor * __ldg(x_sumsq + u) / elems) : 1.0f;
    const float ns = y ? __ldg(n_std + u) : 0.f;
    float ps = 0.f;
    if (p != nullptr) {
      ps = p_scale ? __ldg(p_scale + u) : 1.0f;
      if (p_sumsq != nullptr) ps *= 1.0f / sqrtf(p_factor * __ldg(p_sumsq + u) / elems);
    }
    float2 hh = make_float2(1.f, 0.f);
    float inv_den = 1.f;
    if (h != nullptr) {
      hh = __ldg(h + u);
      float den = hh.x * hh.x + hh.y * hh.y;
      if (detector == 2) den += ns * ns * 2.0f;
      inv_den = 1.0f / den;
    }
    const int64_t base = (int64_t)u * unit4;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < unit4; j += gridDim.x * blockDim.x) {
      const int64_t i = base + j;
      float4 v = ld_stream(x + i);
      if (x_sumsq != nullptr) { v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc; }
      if (x_norm != nullptr) x_norm[i] = v;
      if (y == nullptr) continue;                 // dsc_power_normalize: normalise only
      float4 z;
      if (noise != nullptr) {
        z = ld_stream(noise + i);
      } else {
        uint4 r = Philox::gen((uint64_t)i, offset, seed);
        float2 a = Philox::box_muller(r.x, r.y), b = Philox::box_muller(r.z, r.w);
        z = make_float4(a.x, a.y, b.x, b.y);
      }
      float4 o;
      if (h == nullptr) {
        o = make_float4(v.x + ns * z.x, v.y + ns * z.y, v.z + ns * z.z, v.w + ns * z.w);
        if (p != nullptr) {
          float4 pp = ld_stream(p + i);
          o.x = fmaf(ps, pp.x, o.x); o.y = fmaf(ps, pp.y, o.y); o.z = fmaf(ps, pp.z, o.z); o.w = fmaf(ps, pp.w, o.w);
        }
      } else {
        // y = x*h + n  (two complex symbols: (x,y) and (z,w))
        float yr0 = v.x * hh.x - v.y * hh.y + ns * z.x, yi0 = v.x * hh.y + v.y * hh.x + ns * z.y;
        float yr1 = v.z * hh.x - v.w * hh.y + ns * z.z, yi1 = v.z * hh.y + v.w * hh.x + ns * z.w;
        if (detector != 0) {
          float er0 = (yr0 * hh.x + yi0 * hh.y) * inv_den, ei0 = (yi0 * hh.x - yr0 * hh.y) * inv_den;
          float er1 = (yr1 * hh.x + yi1 * hh.y) * inv_den, ei1 = (yi1 * hh.x - yr1 * hh.y) * inv_den;
          yr0 = er0; yi0 = ei0; yr1 = er1; yi1 = ei1;
        }
        o = make_float4(yr0, yi0, yr1, yi1);
      }
      st_stream(y + i, o);
    }
  }
}

// grid for n_units units of unit4 float4 each: ~8 CTAs of 256 threads per SM in total, at least one chunk per unit
static inline dim3 channel_grid(int n_units, int64_t unit4) {
  int64_t per_unit = (unit4 + 255) / 256;
  int64_t want_x = ((int64_t)kSMs * 8 + n_units - 1) / n_units;
  int gx = (int)(per_unit < want_x ? per_unit : want_x);
  if (gx < 1) gx = 1;
  int gy = n_units < 65535 ? n_units : 65535;
  return dim3(gx, gy);
}

}  // namespace dsc

using namespace dsc;

extern "C" int dsc_channel(const float* x, const float* x_sumsq, float x_factor,
                           const float* noise, uint64_t seed, uint64_t offset,
                           const float* p, const float* p_sumsq, float p_factor, const float* p_scale,
                           const float* h, const float* n_std, int detector,
                           float* y, float* x_norm, int n_units, int64_t elems_per_unit, void* stream) {
  DSC_REQUIRE(x && y && n_std, "dsc_channel: x, y and n_std are required");
  DSC_REQUIRE(n_units >= 0 && elems_per_unit > 0 && (elems_per_unit & 3) == 0 && elems_per_unit < ((int64_t)1 << 32),
              "dsc_channel: elems_per_unit must be a positive multiple of 4 below 2^32");
  DSC_REQUIRE(aligned16(x) && aligned16(y) && (!noise || aligned16(noise)) && (!p || aligned16(p)) && (!x_norm || aligned16(x_norm)),
              "dsc_channel: tensors must be 16-byte aligned");
  if (detector < 0 || detector > 2) { set_error("detector must in LS and MMSE"); return DSC_ERR_BAD_ARG; }
  if (n_units == 0) return DSC_OK;
  const int64_t unit4 = elems_per_unit / 4;
  channel_kernel<<<channel_grid(n_units, unit4), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(x), x_sumsq, x_factor, reinterpret_cast<const float4*>(noise), seed, offset,
      reinterpret_cast<const float4*>(p), p_sumsq, p_factor, p_scale, reinterpret_cast<const float2*>(h), n_std,
      detector, reinterpret_cast<float4*>(y), reinterpret_cast<float4*>(x_norm), n_units, (int)unit4);
  return check_launch("dsc_channel");
}

extern "C" int dsc_power_normalize(const float* x, const float* sumsq, float factor, float* out,
                                   int n_units, int64_t elems_per_unit, void* stream) {
  DSC_REQUIRE(x && sumsq && out, "dsc_power_normalize: null pointer");
  DSC_REQUIRE(n_units >= 0 && elems_per_unit > 0 && (elems_per_unit & 3) == 0 && elems_per_unit < ((int64_t)1 << 32) &&
              aligned16(x) && aligned16(out), "dsc_power_normalize: bad sizes or alignment");
  if (n_units == 0) return DSC_OK;
  const int64_t unit4 = elems_per_unit / 4;
  channel_kernel<<<channel_grid(n_units, unit4), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(x), sumsq, factor, nullptr, 0, 0, nullptr, nullptr, 1.f, nullptr, nullptr,
      sumsq /* unused */, 0, nullptr, reinterpret_cast<float4*>(out), n_units, (int)unit4);
  return check_launch("dsc_power_normalize");
}
