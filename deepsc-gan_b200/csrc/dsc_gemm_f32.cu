// fp32 FFMA GEMM for the Dense layers (prec 0): y = act(x @ w + bias).
// Used for parity-exact fp32 arithmetic and for the odd shapes (K = 16, N = 16) that are too
// small for a tensor-core tile.  128x128x16 CTA tile, 256 threads, 8x8 register tile per thread,
// register-prefetch double buffering.  Compute-bound on the FP32 pipe (no tensor cores).
#include "dsc_common.cuh"

namespace dsc {

constexpr int BM = 128, BN = 128, BK = 16, APAD = 4;

__global__ void __launch_bounds__(256, 2)
gemm_f32_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
                const float* __restrict__ bias, float* __restrict__ C, int64_t ldc,
                int M, int K, int N, int act, int row_mod, int row_skip) {
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  // global->register staging: A tile 128x16 (2 float4 per thread), B tile 16x128 (2 float4 per thread)
  float4 ra[2], rb[2];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int idx = tid + i * 256;
      int r = idx >> 2, k4 = (idx & 3) << 2;
      int gr = m0 + r;
      ra[i] = (gr < M) ? __ldg(reinterpret_cast<const float4*>(A + (int64_t)gr * lda + k0 + k4))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
      int kr = idx >> 5, c4 = (idx & 31) << 2;
      int gc = n0 + c4;
      rb[i] = (gc < ldb && gc < ((N + 3) & ~3))
                  ? __ldg(reinterpret_cast<const float4*>(B + (int64_t)(k0 + kr) * ldb + gc))
                  : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int idx = tid + i * 256;
      int r = idx >> 2, k4 = (idx & 3) << 2;
      As[buf][k4 + 0][r] = ra[i].x;
      As[buf][k4 + 1][r] = ra[i].y;
      As[buf][k4 + 2][r] = ra[i].z;
      As[buf][k4 + 3][r] = ra[i].w;
      int kr = idx >> 5, c4 = (idx & 31) << 2;
      *reinterpret_cast<float4*>(&Bs[buf][kr][c4]) = rb[i];
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = K / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r >= M) continue;
    if (row_mod > 0 && (r % row_mod) == row_skip) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      int c = n0 + jh * 64 + tx * 4;
      if (c >= N) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t = acc[i][jh * 4 + j];
        if (bias != nullptr && c + j < N) t += __ldg(bias + c + j);
        if (act == 1) t = fmaxf(t, 0.f);
        v[j] = t;
      }
      float* dst = C + (int64_t)r * ldc + c;
      if (vec_ok && c + 3 < N) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c + j < N) dst[j] = v[j];
      }
    }
  }
}

int linear_f32(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias,
               float* y, int64_t ldy, int M, int K, int N, int act, int row_mod, int row_skip,
               cudaStream_t stream) {
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  gemm_f32_kernel<<<grid, 256, 0, stream>>>(x, ldx, w, ldw, bias, y, ldy, M, K, N, act, row_mod, row_skip);
  return check_launch("dsc_linear(f32)");
}

}  // namespace dsc
