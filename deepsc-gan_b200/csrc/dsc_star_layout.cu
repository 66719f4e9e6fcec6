// Layout helpers of the fused star-cycle kernel (dsc_star_fused.cu): row-major tiles -> the "interleaved tile" layout
// [tile][k/4][row][4 floats] (thread t of a CTA owns TMEM lane t = row t, so a warp reading element group k/4 of its 32
// rows touches 512 contiguous bytes), and the scatter of one cached target-key row per sentence into it.
#include "dsc_common.cuh"

namespace dsc {

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
