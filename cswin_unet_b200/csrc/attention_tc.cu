// attention_tc.cu — bf16 tcgen05 / TMEM / TMA LePE stripe attention (placeholder until the kernel lands).
#include "common.cuh"
namespace cswin {
int lepe_attention_fwd_tc(const cswin_lepe_branch_t*, int, int, int, float, cudaStream_t, bool* handled) {
  *handled = false;
  return CSWIN_OK;
}
int lepe_attention_bwd_simt(const cswin_lepe_branch_grad_t*, int, int, int, float, int, cudaStream_t) {
  set_error("lepe_attention_bwd: not implemented yet");
  return CSWIN_ERR_UNSUPPORTED;
}
}  // namespace cswin
