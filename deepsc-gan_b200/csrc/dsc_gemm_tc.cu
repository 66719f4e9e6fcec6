// tcgen05 Dense path (prec 1 = bf16x3 split, prec 2 = bf16).  Placeholder until the UMMA kernel lands.
#include "dsc_common.cuh"
namespace dsc {
int linear_tc(const float*, int64_t, const float*, int64_t, const float*, float*, int64_t, int, int, int, int, int, int,
              int prec, cudaStream_t) {
  set_error("dsc_linear: tensor-core path (prec=%d) not built in this revision", prec);
  return DSC_ERR_UNSUPPORTED;
}
}  // namespace dsc
