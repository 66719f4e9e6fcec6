// tcgen05 GEMM for the backward pass (K17): C[M,N] (+)= A[M,K] * B[N,K]^T with BOTH operands row-major fp32 in HBM and
// contiguous along K - the shape of dX = dY @ W^T (utlis/trainer.py / gan_train.py tapes through Dense) and, after a
// transposing copy of the two operands, of dW^T = dY^T @ X.  No pre-packed operand: a CTA converts its 128 x 128 K-chunk
// of A and of B to bf16 hi/lo planes (K-major, 128-byte swizzle) in shared memory, one elected lane issues the three
// bf16x3 passes into a 128 x 128 fp32 accumulator in TMEM.  Split-K over gridDim.z (partials meet by atomicAdd on a
// zeroed C) fills the SMs when M*N is small and K is vocabulary-sized.  The fp32 FFMA kernel it replaces ran the eight
// vocabulary-sized GEMMs of a GAN training step at 19-25 TFLOP/s (4.0 ms of the 9.1 ms step).
#include "dsc_common.cuh"
#include "dsc_tc.cuh"

namespace dsc {

using namespace tc;

namespace nt {
constexpr int BM = 128, BN = 128, KC = 128, kThreads = 256;
constexpr uint32_t PLANE = 128 * 128;              // one (part, kb) plane: 128 rows x 128 B

// rows [r0, r0 + 128) x k [k0, k0 + 128) of a row-major fp32 matrix -> planes [part][kb][128 rows][128 B] at dst.
// Warp w takes rows w, w+8, ...; a lane moves two consecutive k per K-block (8-byte loads: rows of a [*, 22234] matrix
// are only 8-byte aligned), zero beyond the matrix.  All 32 loads of a thread are issued before the first conversion:
// the staging is latency-bound otherwise (one DRAM round trip per unrolled group).
__device__ __forceinline__ void load_operand(const float* __restrict__ src, int64_t ld, int rows, int r0, int k0, int k_end,
                                             int warp, int lane, float2 (&v)[32]) {
  // branch-free and consumed late: K is even (eligibility), so a lane's pair is inside or outside as a whole;
  // out-of-range lanes read element 0 of the matrix, and the masking happens in store_operand, so that nothing here
  // depends on a pending load and all 32 go out back to back
  const int k_in = lane << 1;
  const bool k_ok0 = k0 + k_in < k_end, k_ok1 = k0 + 64 + k_in < k_end;
#pragma unroll
  for (int it = 0; it < 16; ++it) {
    const int gr = r0 + warp + 8 * it;
    const bool r_ok = gr < rows;
    const float* row = src + (r_ok ? (int64_t)gr * ld : 0) + k0 + k_in;
    v[2 * it] = __ldg(reinterpret_cast<const float2*>((r_ok && k_ok0) ? row : src));
    v[2 * it + 1] = __ldg(reinterpret_cast<const float2*>((r_ok && k_ok1) ? row + 64 : src));
  }
}
__device__ __forceinline__ void store_operand(const float2 (&v)[32], int rows, int r0, int k0, int k_end, uint8_t* dst,
                                              int warp, int lane) {
  const uint32_t k_in = (uint32_t)(lane << 1);
  const bool k_ok[2] = {k0 + (int)k_in < k_end, k0 + 64 + (int)k_in < k_end};
#pragma unroll
  for (int it = 0; it < 16; ++it) {
    const uint32_t r = (uint32_t)(warp + 8 * it);
    const bool r_ok = r0 + (int)r < rows;
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
      const bool ok = r_ok && k_ok[kb];
      uint32_t h, l;
      split2(ok ? v[2 * it + kb].x : 0.f, ok ? v[2 * it + kb].y : 0.f, h, l);
      const uint32_t off = kb * PLANE + sw128_offset(r, k_in);
      *reinterpret_cast<uint32_t*>(dst + off) = h;
      *reinterpret_cast<uint32_t*>(dst + 2 * PLANE + off) = l;
    }
  }
}
}  // namespace nt

__global__ void __launch_bounds__(nt::kThreads, 1)
gemm_nt_tc_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
                  float* __restrict__ C, int64_t ldc, int M, int N, int K, int chunks_per_split, int atomic) {
  using namespace nt;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sA = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = sA + 4 * PLANE;
  __shared__ __align__(8) uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int n_chunks_total = (K + KC - 1) / KC;
  const int c_begin = blockIdx.z * chunks_per_split;
  const int c_end = min(n_chunks_total, c_begin + chunks_per_split);
  if (tid == 0) { mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<128>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  constexpr uint32_t IDESC = idesc_bf16_f32(BM, BN);

  // The loads of chunk c + 1 are issued before the wait on chunk c's UMMAs, so the HBM round trip of the next operands
  // overlaps the tensor work and the conversion of the current ones.
  float2 va[32], vb[32];
  if (c_begin < c_end) {
    load_operand(A, lda, M, m0, c_begin * KC, K, warp, lane, va);
    load_operand(B, ldb, N, n0, c_begin * KC, K, warp, lane, vb);
  }
  for (int c = c_begin; c < c_end; ++c) {
    if (c > c_begin) {                     // the UMMAs of the previous chunk have consumed the planes
      mbar_wait(&bar_mma, (uint32_t)(c - c_begin - 1) & 1u);
      tc_fence_after();
    }
    store_operand(va, M, m0, c * KC, K, sA, warp, lane);
    store_operand(vb, N, n0, c * KC, K, sB, warp, lane);
    if (c + 1 < c_end) {
      load_operand(A, lda, M, m0, (c + 1) * KC, K, warp, lane, va);
      load_operand(B, ldb, N, n0, (c + 1) * KC, K, warp, lane, vb);
    }
    fence_async_smem();
    __syncthreads();
    if (warp == 0) {
      const bool leader = elect_one();
      tc_fence_after();
      const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
      if (leader) {
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t pa = (pass == 1) ? 1u : 0u, pb = (pass == 2) ? 1u : 0u;       // hi*hi, lo*hi, hi*lo
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_ss(tmem_d, smem_desc_sw128(a_base + (pa * 2 + kb) * PLANE + ks * 32u),
                      smem_desc_sw128(b_base + (pb * 2 + kb) * PLANE + ks * 32u), IDESC,
                      (c > c_begin || pass > 0 || kb > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&bar_mma);
      }
      __syncwarp();
    }
  }
  if (c_end > c_begin) {
    mbar_wait(&bar_mma, (uint32_t)(c_end - c_begin - 1) & 1u);
    tc_fence_after();
    // epilogue through shared memory (the operand planes are dead): a thread owns one accumulator ROW, so direct stores
    // would touch 32 rows per instruction; staged as a [128][129] fp32 tile, every warp then writes whole rows, 128
    // contiguous bytes per instruction (plain stores, or atomics when partial sums meet)
    float* tile = reinterpret_cast<float*>(sA);
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane;
#pragma unroll 1
    for (int j = 0; j < 2; ++j) {
      float v[32];
      const int c0 = half * 64 + j * 32;
      tmem_ld32(tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) tile[r * 129 + c0 + e] = v[e];
    }
    __syncthreads();
#pragma unroll 1
    for (int it = 0; it < 16; ++it) {
      const int rr = warp + 8 * it, row = m0 + rr;
      if (row >= M) break;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = n0 + j * 32 + lane;
        if (col < N) {
          float* dst = C + (int64_t)row * ldc + col;
          const float val = tile[rr * 129 + j * 32 + lane];
          if (atomic) atomicAdd(dst, val); else *dst = val;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tmem_d);
}

// dst[c][r] = src[r][c]: 32 x 32 tiles through shared memory, coalesced both ways
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ src, int64_t ld_src, float* __restrict__ dst, int64_t ld_dst, int rows, int cols) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int tiles_c = (cols + 31) / 32;
  const int64_t n_tiles = (int64_t)((rows + 31) / 32) * tiles_c;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int r0 = (int)(t / tiles_c) * 32, c0 = (int)(t % tiles_c) * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + ty + 8 * i, c = c0 + tx;
      tile[ty + 8 * i][tx] = (r < rows && c < cols) ? __ldg(src + (int64_t)r * ld_src + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + ty + 8 * i, r = r0 + tx;
      if (c < cols && r < rows) dst[(int64_t)c * ld_dst + r] = tile[tx][ty + 8 * i];
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
zero_matrix_kernel(float* __restrict__ C, int64_t ldc, int rows, int cols) {
  const int64_t total = (int64_t)rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    C[(i / cols) * ldc + (i % cols)] = 0.f;
}

// whether the tensor-core path takes C (+)= A B^T: worth it only when the product is large, and the 8-byte loads of the
// staging need even leading dimensions and 8-byte aligned bases
bool gemm_nt_tc_eligible(const float* A, int64_t lda, const float* B, int64_t ldb, int M, int N, int K) {
  if ((double)M * (double)N * (double)K < 1.6e7) return false;   // from [1984,128] x [128,128]^T up: 16 CTAs of one chunk beat the FFMA kernel
  if ((lda & 1) || (ldb & 1) || (K & 1)) return false;
  return ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 7u) == 0;
}

int gemm_nt_tc(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int M, int N, int K,
               int accumulate, cudaStream_t s) {
  using namespace nt;
  constexpr size_t smem = 8 * (size_t)PLANE + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_nt_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("dsc_gemm_nt_tc: %s", cudaGetErrorString(e)); return DSC_ERR_CUDA; }
    attr_set = true;
  }
  const int tm = (M + BM - 1) / BM, tn = (N + BN - 1) / BN, chunks = (K + KC - 1) / KC;
  // split-K only when the output tiles do not fill the SMs and the partial sums are few (every split adds M*N atomics:
  // at the 174 vocabulary tiles of dW a 5-way split cost more in atomics than its shorter tail wave saved)
  int splits = 1;
  if (tm * tn < kSMs) {
    splits = (kSMs + tm * tn - 1) / (tm * tn);
    if (splits > chunks) splits = chunks;
    while (splits > 1 && (double)M * N * splits > 8.0e6) --splits;
  }
  const int cps = (chunks + splits - 1) / splits;
  splits = (chunks + cps - 1) / cps;
  const int atomic = (splits > 1 || accumulate) ? 1 : 0;
  if (splits > 1 && !accumulate) {
    const int64_t want = ((int64_t)M * N + 255) / 256;
    zero_matrix_kernel<<<(int)(want < kSMs * 8 ? want : kSMs * 8), 256, 0, s>>>(C, ldc, M, N);
  }
  gemm_nt_tc_kernel<<<dim3(tn, tm, splits), kThreads, smem, s>>>(A, lda, B, ldb, C, ldc, M, N, K, cps, atomic);
  return check_launch("dsc_gemm_nt_tc");
}

}  // namespace dsc

using namespace dsc;

extern "C" int dsc_gemm_nt_tc(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                              int M, int N, int K, int accumulate, void* stream) {
  DSC_REQUIRE(A && B && C, "dsc_gemm_nt_tc: null pointer");
  DSC_REQUIRE(M >= 0 && N >= 0 && K > 0 && lda >= K && ldb >= K, "dsc_gemm_nt_tc: bad shape");
  DSC_REQUIRE(ldc >= N, "dsc_gemm_nt_tc: ldc too small");
  DSC_REQUIRE(!(lda & 1) && !(ldb & 1) && !(K & 1) && ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 7u) == 0,
              "dsc_gemm_nt_tc: operands must be 8-byte aligned with even K and leading dimensions");
  if (M == 0 || N == 0) return DSC_OK;
  return gemm_nt_tc(A, lda, B, ldb, C, ldc, M, N, K, accumulate, as_stream(stream));
}

extern "C" int dsc_transpose(const float* src, int64_t ld_src, float* dst, int64_t ld_dst, int rows, int cols, void* stream) {
  DSC_REQUIRE(src && dst && rows >= 0 && cols >= 0 && ld_src >= cols && ld_dst >= rows, "dsc_transpose: bad argument");
  if (rows == 0 || cols == 0) return DSC_OK;
  const int64_t tiles = (int64_t)((rows + 31) / 32) * ((cols + 31) / 32);
  transpose_kernel<<<(int)(tiles < (int64_t)kSMs * 16 ? tiles : (int64_t)kSMs * 16), 256, 0, as_stream(stream)>>>(
      src, ld_src, dst, ld_dst, rows, cols);
  return check_launch("dsc_transpose");
}
