// Micro-benchmark entry point (not on the product path): issue rate of tcgen05.mma kind::f16, M = 128, K = 16, with the
// A operand in tensor memory (TS mode) or in shared memory (SS mode), for N = 32..256.  Used to decide the operand
// placement / tile shapes of the fused kernels.  One CTA, one issuing thread, `iters` x 8 back-to-back UMMAs
// into one accumulator, timed with clock64 between the first issue and the completion of the last commit.
#include "dsc_common.cuh"
#include "dsc_tc.cuh"

namespace dsc {
using namespace tc;

template <bool ELECT>
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(int ts_mode, int n, int iters, long long* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  // operands: zero-filled shared memory (A: 128 rows x 128 B x 2 K-blocks, B: 256 rows likewise), zero TMEM A columns
  for (int i = tid; i < (2 * 128 * 128 + 2 * 256 * 128) / 16; i += 128) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  {
    uint32_t z[16] = {0};
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 4; ++c) tmem_st16(lane_addr + 256 + c * 16, z);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (ELECT ? (warp == 0) : (tid == 0)) {      // ELECT: whole warp converged + elect.sync; else the divergent `tid == 0` form
    const bool leader = ELECT ? elect_one() : true;
    tc_fence_after();
    const uint32_t a_base = smem_u32(sm), b_base = a_base + 2 * 128 * 128;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t db = smem_desc_sw128(b_base + kb * 256 * 128 + ks * 32);
          if (leader) {
            if (ts_mode) umma_ts(tmem_base, tmem_base + 256 + (kb * 4 + ks) * 8, db, idesc, 1u);
            else umma_ss(tmem_base, smem_desc_sw128(a_base + kb * 128 * 128 + ks * 32), db, idesc, 1u);
          }
        }
    }
    if (leader) umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (leader) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}
}  // namespace dsc

using namespace dsc;

extern "C" int dsc_umma_probe(int ts_mode, int n, int iters, long long* cycles_dev, void* stream) {
  DSC_REQUIRE(ts_mode >= 0 && ts_mode <= 3 && cycles_dev && n >= 16 && n <= 256 && (n % 16) == 0 && iters > 0, "dsc_umma_probe: bad argument");
  constexpr size_t smem = 2 * 128 * 128 + 2 * 256 * 128 + 1024;
  cudaError_t e = cudaFuncSetAttribute(umma_probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(umma_probe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("dsc_umma_probe: %s", cudaGetErrorString(e)); return DSC_ERR_CUDA; }
  if (ts_mode & 2) umma_probe_kernel<true><<<1, 128, smem, as_stream(stream)>>>(ts_mode & 1, n, iters, cycles_dev);
  else umma_probe_kernel<false><<<1, 128, smem, as_stream(stream)>>>(ts_mode & 1, n, iters, cycles_dev);
  return check_launch("dsc_umma_probe");
}
