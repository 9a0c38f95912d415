// gemm_tc.cu — bf16 tcgen05 / TMEM / TMA Linear (placeholder until the kernel lands).
#include "common.cuh"
namespace cswin {
int linear_fwd_tc(const cswin_linear_args_t*, cudaStream_t, bool* handled) {
  *handled = false;
  return CSWIN_OK;
}
}  // namespace cswin
