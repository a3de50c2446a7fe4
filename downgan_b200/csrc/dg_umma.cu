// tcgen05 / TMEM / TMA implicit-GEMM convolution path (bf16 operands, fp32 accumulate).
#include "dg_common.cuh"

namespace dg {
bool umma_supported(const ConvOp&) { return false; }
int conv_umma(const ConvOp&, cudaStream_t) {
  set_error("conv_umma: not built");
  return DG_ERR_STATE;
}
}  // namespace dg

extern "C" int dg_has_tcgen05(void) { return 0; }
