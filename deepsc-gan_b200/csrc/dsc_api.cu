// C-ABI glue: identity, errors, and the Dense / vocab-projection entry points that dispatch on `prec`.
#include "dsc_common.cuh"
#include <stdarg.h>

namespace dsc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int linear_f32(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias,
               float* y, int64_t ldy, int M, int K, int N, int act, int row_mod, int row_skip,
               cudaStream_t stream);
int linear_tc(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias,
              float* y, int64_t ldy, int M, int K, int N, int act, int row_mod, int row_skip,
              int prec, cudaStream_t stream);
int vocab_argmax_tc(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias,
                    int32_t* ids, int64_t ids_stride, float* workspace, int64_t workspace_floats,
                    int M, int N, int prec, cudaStream_t stream);
int64_t vocab_argmax_tc_workspace(int M, int N);

}  // namespace dsc

using namespace dsc;

extern "C" int dsc_version(void) { return 100; }   // 0.1.0

extern "C" const char* dsc_last_error(void) { return g_err; }

extern "C" int dsc_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { set_error("dsc_device_arch: no CUDA device"); return DSC_ERR_CUDA; }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return major * 10 + minor;
}

extern "C" int dsc_linear(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias,
                          float* y, int64_t ldy, int M, int K, int N, int act,
                          int row_mod, int row_skip, int prec, void* stream) {
  DSC_REQUIRE(x && w && y, "dsc_linear: null pointer");
  DSC_REQUIRE(M >= 0 && K > 0 && N > 0, "dsc_linear: bad sizes M=%d K=%d N=%d", M, K, N);
  DSC_REQUIRE((K % 16) == 0, "dsc_linear: K=%d must be a multiple of 16", K);
  DSC_REQUIRE((ldx & 3) == 0 && ldx >= K && aligned16(x), "dsc_linear: x rows must be 16-byte aligned (ldx=%lld)", (long long)ldx);
  DSC_REQUIRE((ldw & 3) == 0 && ldw >= ((N + 3) & ~3) && aligned16(w), "dsc_linear: ldw=%lld must be a multiple of 4 and >= round_up(N,4)", (long long)ldw);
  DSC_REQUIRE(ldy >= N, "dsc_linear: ldy < N");
  DSC_REQUIRE(act == 0 || act == 1, "dsc_linear: act must be 0 or 1");
  if (M == 0) return DSC_OK;
  if (prec == 0) return linear_f32(x, ldx, w, ldw, bias, y, ldy, M, K, N, act, row_mod, row_skip, as_stream(stream));
  if (prec == 1 || prec == 2)
    return linear_tc(x, ldx, w, ldw, bias, y, ldy, M, K, N, act, row_mod, row_skip, prec, as_stream(stream));
  set_error("dsc_linear: prec must be 0, 1 or 2");
  return DSC_ERR_BAD_ARG;
}

extern "C" int64_t dsc_vocab_argmax_workspace(int M, int N) {
  if (M <= 0 || N <= 0) return 0;
  return (int64_t)M * (((int64_t)N + 3) & ~3LL);     // stage-A: materialised logits rows (ld = round_up(N,4))
}

extern "C" int dsc_vocab_argmax(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias,
                                int32_t* ids, int64_t ids_stride, float* logits, int64_t ld_logits,
                                float* workspace, int64_t workspace_floats,
                                int M, int N, int prec, void* stream) {
  DSC_REQUIRE(x && w && ids, "dsc_vocab_argmax: null pointer");
  DSC_REQUIRE(M >= 0 && N > 0, "dsc_vocab_argmax: bad sizes");
  if (M == 0) return DSC_OK;
  float* lg = logits;
  int64_t ld = ld_logits;
  if (lg == nullptr) {
    ld = ((int64_t)N + 3) & ~3LL;
    DSC_REQUIRE(workspace && workspace_floats >= (int64_t)M * ld, "dsc_vocab_argmax: workspace too small");
    lg = workspace;
  }
  int rc = dsc_linear(x, ldx, w, ldw, bias, lg, ld, M, DSC_D_MODEL, N, 0, 0, 0, prec, stream);
  if (rc != DSC_OK) return rc;
  return dsc_argmax_rows(lg, ld, ids, ids_stride, M, N, stream);
}
