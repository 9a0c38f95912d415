// api.cu — extern "C" boundary of libcswin_b200.so (declared in include/cswin_b200.h).
// Argument validation, dtype / kernel-family selection, thread-local error text.  No torch types, no allocation,
// no synchronisation: every entry point only enqueues kernels on the caller's stream.
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace cswin {

std::atomic<uint64_t> g_launches{0};
std::atomic<uint64_t> g_tc_launches{0};
std::atomic<unsigned long long*> g_trace{nullptr};
std::atomic<int> g_gemm_smem_cap_kb{0};

namespace {
thread_local char t_err[512] = "";
std::mutex g_dev_mu;
int g_sm_count[64] = {0};
}  // namespace

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("CSWIN_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

int sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (g_sm_count[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    g_sm_count[dev] = n;
  }
  return g_sm_count[dev];
}

static bool valid_dtype(int dtype) { return dtype == CSWIN_F32 || dtype == CSWIN_BF16; }

// bf16 calls that fall outside a tcgen05 kernel's envelope (window > 256 tokens, head_dim != 32, unaligned rows) run on the
// general SIMT kernels: same results, up to ~45x slower.  That must not happen silently: every such launch is counted
// (cswin_simt_fallback_count) and the FIRST one of each op prints one line to stderr (CSWIN_QUIET_FALLBACK=1 silences it).
std::atomic<uint64_t> g_simt_fallbacks{0};
static void note_simt_fallback(int op, const char* what) {
  static std::atomic<int> warned[4] = {{0}, {0}, {0}, {0}};
  g_simt_fallbacks.fetch_add(1, std::memory_order_relaxed);
  static const bool quiet = [] { const char* e = getenv("CSWIN_QUIET_FALLBACK"); return e && e[0] == '1'; }();
  if (!quiet && warned[op & 3].exchange(1) == 0)
    fprintf(stderr, "[cswin_b200] warning: bf16 %s is outside the tcgen05 kernel's envelope (see include/cswin_b200.h) and runs on the "
                    "general SIMT kernel — correct but much slower; further occurrences are only counted (cswin_simt_fallback_count).\n", what);
}

}  // namespace cswin

using namespace cswin;

extern "C" {

int cswin_abi_version(void) { return CSWIN_ABI_VERSION; }
const char* cswin_last_error(void) { return t_err; }
uint64_t cswin_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
void cswin_debug_set_trace(void* device_buffer) { g_trace.store((unsigned long long*)device_buffer); }
int cswin_set_option(int32_t option, int32_t value) {
  if (option == CSWIN_OPT_GEMM_SMEM_CAP_KB && value >= 0 && value <= 227) { g_gemm_smem_cap_kb.store(value); return CSWIN_OK; }
  set_error("set_option: unknown option %d or value %d out of range", option, value);
  return CSWIN_ERR_INVALID;
}
uint64_t cswin_tc_launch_count(void) { return g_tc_launches.load(std::memory_order_relaxed); }
uint64_t cswin_simt_fallback_count(void) { return g_simt_fallbacks.load(std::memory_order_relaxed); }

int cswin_lepe_attention_fwd(const cswin_lepe_branch_t* branches, int32_t n_branches, int32_t B, int32_t reso,
                             float scale, int32_t dtype, cswin_stream_t stream) {
  CSWIN_REQUIRE(valid_dtype(dtype), CSWIN_ERR_INVALID, "lepe_attention_fwd: bad dtype %d", dtype);
  CSWIN_REQUIRE(branches && (n_branches == 1 || n_branches == 2), CSWIN_ERR_INVALID, "lepe_attention_fwd: n_branches must be 1 or 2");
  CSWIN_REQUIRE(B >= 0 && reso > 0, CSWIN_ERR_INVALID, "lepe_attention_fwd: bad B=%d reso=%d", B, reso);
  if (B == 0) return CSWIN_OK;
  static const bool fwd_simt = [] { const char* e = getenv("CSWIN_ATTN_FWD_SIMT"); return e && e[0] == '1'; }();   // A/B debugging aid
  if (dtype == CSWIN_BF16 && !fwd_simt) {
    bool handled = false;
    int rc = lepe_attention_fwd_tc(branches, n_branches, B, reso, scale, (cudaStream_t)stream, &handled);
    if (rc != CSWIN_OK || handled) return rc;
    note_simt_fallback(0, "cswin_lepe_attention_fwd");
  }
  return lepe_attention_fwd_simt(branches, n_branches, B, reso, scale, dtype, (cudaStream_t)stream);
}

int cswin_lepe_attention_bwd(const cswin_lepe_branch_grad_t* branches, int32_t n_branches, int32_t B, int32_t reso,
                             float scale, int32_t dtype, cswin_stream_t stream) {
  CSWIN_REQUIRE(valid_dtype(dtype), CSWIN_ERR_INVALID, "lepe_attention_bwd: bad dtype %d", dtype);
  CSWIN_REQUIRE(branches && (n_branches == 1 || n_branches == 2), CSWIN_ERR_INVALID, "lepe_attention_bwd: n_branches must be 1 or 2");
  CSWIN_REQUIRE(B >= 0 && reso > 0, CSWIN_ERR_INVALID, "lepe_attention_bwd: bad B=%d reso=%d", B, reso);
  if (B == 0) return CSWIN_OK;
  static const bool force_simt = [] { const char* e = getenv("CSWIN_ATTN_BWD_SIMT"); return e && e[0] == '1'; }();   // A/B debugging aid
  if (dtype == CSWIN_BF16 && !force_simt) {
    bool handled = false;
    int rc = lepe_attention_bwd_tc(branches, n_branches, B, reso, scale, (cudaStream_t)stream, &handled);
    if (rc != CSWIN_OK || handled) return rc;
    note_simt_fallback(1, "cswin_lepe_attention_bwd");
  }
  return lepe_attention_bwd_simt(branches, n_branches, B, reso, scale, dtype, (cudaStream_t)stream);
}

int cswin_lepe_param_grad(const cswin_lepe_branch_grad_t* branches, int32_t n_branches, int32_t B, int32_t reso,
                          int32_t dtype, cswin_stream_t stream, int32_t* handled) {
  CSWIN_REQUIRE(valid_dtype(dtype) && handled, CSWIN_ERR_INVALID, "lepe_param_grad: bad dtype %d / null handled", dtype);
  CSWIN_REQUIRE(branches && (n_branches == 1 || n_branches == 2), CSWIN_ERR_INVALID, "lepe_param_grad: n_branches must be 1 or 2");
  CSWIN_REQUIRE(B >= 0 && reso > 0, CSWIN_ERR_INVALID, "lepe_param_grad: bad B=%d reso=%d", B, reso);
  *handled = 0;
  if (B == 0 || dtype != CSWIN_BF16) return CSWIN_OK;
  bool h = false;
  const int rc = lepe_param_grad_tc(branches, n_branches, B, reso, (cudaStream_t)stream, &h);
  *handled = h ? 1 : 0;
  return rc;
}

int cswin_zoom_cubic_fwd(const float* in, int32_t n, int32_t H, int32_t W, double* work, float* out, int64_t out_slice_stride,
                         int64_t out_channel_stride, int32_t channel_copies, int32_t OH, int32_t OW, cswin_stream_t stream) {
  return zoom_cubic(in, n, H, W, work, out, out_slice_stride, out_channel_stride, channel_copies, OH, OW, (cudaStream_t)stream);
}

int cswin_zoom_nearest_u8(const uint8_t* in, int32_t n, int32_t H, int32_t W, uint8_t* out, int32_t OH, int32_t OW,
                          cswin_stream_t stream) {
  return zoom_nearest_u8(in, n, H, W, out, OH, OW, (cudaStream_t)stream);
}

int cswin_layernorm_fwd(const void* x, int64_t ldx, const void* gamma, const void* beta, void* y, int64_t ldy,
                        int64_t M, int32_t C, float eps, float* mean_out, float* rstd_out, int32_t dtype,
                        cswin_stream_t stream) {
  CSWIN_REQUIRE(valid_dtype(dtype), CSWIN_ERR_INVALID, "layernorm_fwd: bad dtype %d", dtype);
  CSWIN_REQUIRE(M >= 0, CSWIN_ERR_INVALID, "layernorm_fwd: negative M");
  return layernorm_fwd(x, ldx, gamma, beta, y, ldy, M, C, eps, mean_out, rstd_out, nullptr, dtype, (cudaStream_t)stream);
}

int cswin_layernorm_stats_fwd(const void* x, int64_t ldx, const void* gamma, const void* beta, void* y, int64_t ldy,
                              int64_t M, int32_t C, float eps, float* row_stats_out, int32_t dtype, cswin_stream_t stream) {
  CSWIN_REQUIRE(valid_dtype(dtype) && M >= 0 && row_stats_out, CSWIN_ERR_INVALID, "layernorm_stats_fwd: bad arguments");
  return layernorm_fwd(x, ldx, gamma, beta, y, ldy, M, C, eps, nullptr, nullptr, row_stats_out, dtype, (cudaStream_t)stream);
}

int cswin_linear_fwd(const cswin_linear_args_t* a, int32_t dtype, cswin_stream_t stream) {
  CSWIN_REQUIRE(valid_dtype(dtype), CSWIN_ERR_INVALID, "linear_fwd: bad dtype %d", dtype);
  CSWIN_REQUIRE(a && a->a && a->w && a->out, CSWIN_ERR_INVALID, "linear_fwd: null pointer");
  CSWIN_REQUIRE(a->M >= 0 && a->N > 0 && a->K1 > 0 && a->K2 >= 0, CSWIN_ERR_INVALID, "linear_fwd: bad M/N/K");
  CSWIN_REQUIRE((a->a2 != nullptr) == (a->K2 > 0), CSWIN_ERR_INVALID, "linear_fwd: a2 and K2 must be given together");
  CSWIN_REQUIRE(a->w_layout == 0 || a->w_layout == 1, CSWIN_ERR_INVALID, "linear_fwd: w_layout must be 0 ((N,K)) or 1 ((K,N))");
  CSWIN_REQUIRE(a->lda >= a->K1 && a->ldw >= (a->w_layout ? a->N : a->K1 + a->K2) && a->ldo >= a->N, CSWIN_ERR_INVALID, "linear_fwd: leading dimension too small");
  CSWIN_REQUIRE(!a->a2 || a->lda2 >= a->K2, CSWIN_ERR_INVALID, "linear_fwd: lda2 too small");
  CSWIN_REQUIRE(!a->residual || a->ldr >= a->N, CSWIN_ERR_INVALID, "linear_fwd: ldr too small");
  CSWIN_REQUIRE((a->ln_gamma != nullptr) == (a->ln_beta != nullptr), CSWIN_ERR_INVALID, "linear_fwd: ln_gamma and ln_beta must be given together");
  CSWIN_REQUIRE(!a->ln_gamma || !a->a2, CSWIN_ERR_INVALID, "linear_fwd: LayerNorm prologue needs a single A source");
  CSWIN_REQUIRE(!a->sample_scale || a->rows_per_sample > 0, CSWIN_ERR_INVALID, "linear_fwd: rows_per_sample must be > 0");
  CSWIN_REQUIRE(a->act == 0 || a->act == 1 || a->act == 2, CSWIN_ERR_INVALID, "linear_fwd: act must be 0 (none), 1 (GELU) or 2 (x GELU'(residual))");
  CSWIN_REQUIRE(a->act != 2 || a->residual, CSWIN_ERR_INVALID, "linear_fwd: act 2 needs the pre-activation in `residual`");
  CSWIN_REQUIRE(!a->aux_out || (a->act == 1 && a->ld_aux >= a->N), CSWIN_ERR_INVALID, "linear_fwd: aux_out needs act 1 and ld_aux >= N");
  if (a->M == 0) return CSWIN_OK;
  const bool folded = a->ln_colsum || a->ln_stats || a->stats_out || a->bias_f32;
  CSWIN_REQUIRE(!folded || dtype == CSWIN_BF16, CSWIN_ERR_UNSUPPORTED, "linear_fwd: LayerNorm folding / stats_out / bias_f32 exist on the bf16 path only");
  CSWIN_REQUIRE((a->ln_colsum != nullptr) == (a->ln_stats != nullptr), CSWIN_ERR_INVALID, "linear_fwd: ln_stats and ln_colsum must be given together");
  CSWIN_REQUIRE(!a->ln_stats || (a->ln_stats_parts > 0 && a->ln_C > 0 && !a->a2 && !a->ln_gamma), CSWIN_ERR_INVALID, "linear_fwd: bad folded-LayerNorm arguments");
  if (dtype == CSWIN_BF16) {
    bool handled = false;
    int rc = linear_fwd_tc(a, (cudaStream_t)stream, &handled);
    if (rc != CSWIN_OK || handled) return rc;
    if (!a->ln_gamma) note_simt_fallback(2, "cswin_linear_fwd");       // (a LayerNorm PROLOGUE is the SIMT kernel's documented job)
  }
  CSWIN_REQUIRE(!a->aux_out && a->act != 2, CSWIN_ERR_UNSUPPORTED, "linear_fwd: the training epilogues (aux_out, act 2) exist on the bf16 tcgen05 path with 16-byte aligned rows only");
  CSWIN_REQUIRE(!folded, CSWIN_ERR_UNSUPPORTED, "linear_fwd: operands are not TMA-compatible (16-byte aligned pointers / row pitches), which the folded-LayerNorm form requires");
  return linear_fwd_simt(a, dtype, (cudaStream_t)stream);
}

int cswin_im2col_tokens(const void* x, int64_t x_bs, int64_t x_ts, void* col, int64_t ldcol, int32_t B, int32_t H,
                        int32_t W, int32_t C, int32_t KH, int32_t KW, int32_t stride, int32_t pad, int32_t dtype,
                        cswin_stream_t stream) {
  CSWIN_REQUIRE(valid_dtype(dtype), CSWIN_ERR_INVALID, "im2col_tokens: bad dtype %d", dtype);
  CSWIN_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0, CSWIN_ERR_INVALID, "im2col_tokens: bad shape");
  return im2col_tokens(x, x_bs, x_ts, col, ldcol, B, H, W, C, KH, KW, stride, pad, dtype, (cudaStream_t)stream);
}

int cswin_conv_tokens_fwd(const void* x, int64_t x_bs, int64_t x_ts, const void* w, int64_t ldw, const void* bias, void* out,
                          int64_t ldo, int32_t B, int32_t H, int32_t W, int32_t C, int32_t N, int32_t KH, int32_t KW, int32_t stride,
                          int32_t pad, int32_t dtype, cswin_stream_t stream, int32_t* handled) {
  CSWIN_REQUIRE(handled != nullptr, CSWIN_ERR_INVALID, "conv_tokens_fwd: null handled");
  *handled = 0;
  CSWIN_REQUIRE(dtype == CSWIN_BF16, CSWIN_ERR_UNSUPPORTED, "conv_tokens_fwd: exists on the bf16 / tcgen05 path only");
  CSWIN_REQUIRE(x && w && out && B >= 0 && H > 0 && W > 0 && C > 0 && N > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0 &&
                ldw >= (int64_t)KH * KW * C && ldo >= N, CSWIN_ERR_INVALID, "conv_tokens_fwd: bad arguments");
  if (B == 0) { *handled = 1; return CSWIN_OK; }
  bool h = false;
  const int rc = conv_tokens_fwd_tc(x, x_bs, x_ts, w, ldw, bias, out, ldo, B, H, W, C, N, KH, KW, stride, pad, (cudaStream_t)stream, &h);
  *handled = h ? 1 : 0;
  return rc;
}

int cswin_im2col_nchw(const void* x, int32_t x_is_f32, void* col, int64_t ldcol, int32_t B, int32_t C, int32_t H,
                      int32_t W, int32_t KH, int32_t KW, int32_t stride, int32_t pad, int32_t dtype,
                      cswin_stream_t stream) {
  CSWIN_REQUIRE(valid_dtype(dtype), CSWIN_ERR_INVALID, "im2col_nchw: bad dtype %d", dtype);
  CSWIN_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0, CSWIN_ERR_INVALID, "im2col_nchw: bad shape");
  return im2col_nchw(x, x_is_f32, col, ldcol, B, C, H, W, KH, KW, stride, pad, dtype, (cudaStream_t)stream);
}

int cswin_carafe_reassemble_fwd(const void* enc, int64_t ldenc, const void* z, int64_t ldz, const void* bias,
                                void* y, int64_t ldy, int32_t nchw_out, int32_t y_is_f32, int32_t B, int32_t H,
                                int32_t W, int32_t C, int32_t up, int32_t dtype, cswin_stream_t stream) {
  CSWIN_REQUIRE(valid_dtype(dtype), CSWIN_ERR_INVALID, "carafe_reassemble_fwd: bad dtype %d", dtype);
  CSWIN_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0, CSWIN_ERR_INVALID, "carafe_reassemble_fwd: bad shape");
  return carafe_reassemble_fwd(enc, ldenc, z, ldz, bias, y, ldy, nchw_out, y_is_f32, B, H, W, C, up, dtype, (cudaStream_t)stream);
}

int cswin_carafe_head_fwd(const void* enc, int64_t ldenc, const void* z, int64_t ldz, const void* bias, void* logits,
                          int32_t logits_is_f32, uint8_t* labels, int32_t B, int32_t H, int32_t W, int32_t C, int32_t up,
                          int32_t dtype, cswin_stream_t stream) {
  CSWIN_REQUIRE(valid_dtype(dtype), CSWIN_ERR_INVALID, "carafe_head_fwd: bad dtype %d", dtype);
  CSWIN_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0, CSWIN_ERR_INVALID, "carafe_head_fwd: bad shape");
  return carafe_head_fwd(enc, ldenc, z, ldz, bias, logits, logits_is_f32, labels, B, H, W, C, up, dtype, (cudaStream_t)stream);
}

int cswin_act_fwd(const void* z, int64_t ldz, void* out, int64_t ldo, int64_t M, int32_t N, int32_t act, int32_t dtype,
                  cswin_stream_t stream) {
  CSWIN_REQUIRE(valid_dtype(dtype) && M >= 0 && N > 0 && (act == 0 || act == 1), CSWIN_ERR_INVALID, "act_fwd: bad arguments");
  return act_fwd(z, ldz, out, ldo, M, N, act, dtype, (cudaStream_t)stream);
}

int cswin_act_bwd(const void* dout, int64_t ldd, const void* z, int64_t ldz, const float* sample_scale,
                  int32_t rows_per_sample, void* dz, int64_t ldo, int64_t M, int32_t N, int32_t act, int32_t dtype,
                  cswin_stream_t stream) {
  CSWIN_REQUIRE(valid_dtype(dtype) && M >= 0 && N > 0 && (act == 0 || act == 1), CSWIN_ERR_INVALID, "act_bwd: bad arguments");
  return act_bwd(dout, ldd, z, ldz, sample_scale, rows_per_sample, dz, ldo, M, N, act, dtype, (cudaStream_t)stream);
}

int cswin_linear_wgrad(const void* dz, int64_t ldz, const void* a, int64_t lda, float* dw, int64_t ldw, float* db, int64_t M,
                       int32_t N, int32_t K, int32_t dtype, cswin_stream_t stream) {
  CSWIN_REQUIRE(valid_dtype(dtype) && M >= 0 && N > 0 && K > 0, CSWIN_ERR_INVALID, "linear_wgrad: bad arguments");
  CSWIN_REQUIRE(ldz >= N && lda >= K && ldw >= K, CSWIN_ERR_INVALID, "linear_wgrad: leading dimension too small");
  if (M == 0) return CSWIN_OK;
  if (dtype == CSWIN_BF16) {
    bool handled = false;
    int rc = linear_wgrad_tc(dz, ldz, a, lda, dw, ldw, db, M, N, K, (cudaStream_t)stream, &handled);
    if (rc != CSWIN_OK || handled) return rc;
    note_simt_fallback(3, "cswin_linear_wgrad");
  }
  return linear_wgrad(dz, ldz, a, lda, dw, ldw, db, M, N, K, dtype, (cudaStream_t)stream);
}

int cswin_layernorm_bwd(const void* x, int64_t ldx, const void* dy, int64_t ldy, const void* gamma, const float* mean,
                        const float* rstd, void* dx, int64_t ldo, const void* dx_add, int64_t ld_add, float* dgamma,
                        float* dbeta, int64_t M, int32_t C, int32_t dtype, cswin_stream_t stream) {
  CSWIN_REQUIRE(valid_dtype(dtype) && M >= 0, CSWIN_ERR_INVALID, "layernorm_bwd: bad arguments");
  CSWIN_REQUIRE(dx_add == nullptr || ld_add >= C, CSWIN_ERR_INVALID, "layernorm_bwd: ld_add too small");
  return layernorm_bwd(x, ldx, dy, ldy, gamma, mean, rstd, dx, ldo, dx_add, dx_add ? ld_add : 0, dgamma, dbeta, M, C, dtype,
                       (cudaStream_t)stream);
}

int cswin_col2im_tokens(const void* dcol, int64_t ldcol, void* dx, int64_t x_bs, int64_t x_ts, int32_t B, int32_t H, int32_t W,
                        int32_t C, int32_t KH, int32_t KW, int32_t stride, int32_t pad, int32_t dtype, cswin_stream_t stream) {
  CSWIN_REQUIRE(valid_dtype(dtype), CSWIN_ERR_INVALID, "col2im_tokens: bad dtype %d", dtype);
  CSWIN_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0, CSWIN_ERR_INVALID, "col2im_tokens: bad shape");
  return col2im_tokens(dcol, ldcol, dx, x_bs, x_ts, B, H, W, C, KH, KW, stride, pad, dtype, (cudaStream_t)stream);
}

int cswin_carafe_reassemble_bwd(const void* enc, int64_t ldenc, const void* z, int64_t ldz, const void* dy, int32_t dy_is_f32,
                                int64_t dy_sb, int64_t dy_sy, int64_t dy_sx, int64_t dy_sc, void* denc, int64_t lddenc,
                                void* dz, int64_t lddz, float* dbias, float* kappa_ws, int32_t B, int32_t H, int32_t W,
                                int32_t C, int32_t up, int32_t dtype, cswin_stream_t stream) {
  CSWIN_REQUIRE(valid_dtype(dtype) && B >= 0 && H > 0 && W > 0, CSWIN_ERR_INVALID, "carafe_reassemble_bwd: bad arguments");
  return carafe_reassemble_bwd(enc, ldenc, z, ldz, dy, dy_is_f32, dy_sb, dy_sy, dy_sx, dy_sc, denc, lddenc, dz, lddz, dbias,
                               kappa_ws, B, H, W, C, up, dtype, (cudaStream_t)stream);
}

int cswin_carafe_head_bwd(const void* enc, int64_t ldenc, const void* z, int64_t ldz, const float* dlogits, void* denc,
                          int64_t lddenc, void* dz, int64_t lddz, int32_t zcols, float* dbias, float* kws, int32_t B, int32_t H,
                          int32_t W, int32_t C, int32_t up, int32_t dtype, cswin_stream_t stream) {
  CSWIN_REQUIRE(valid_dtype(dtype) && B >= 0 && H > 0 && W > 0, CSWIN_ERR_INVALID, "carafe_head_bwd: bad arguments");
  return carafe_head_bwd(enc, ldenc, z, ldz, dlogits, denc, lddenc, dz, lddz, zcols, dbias, kws, B, H, W, C, up, dtype,
                         (cudaStream_t)stream);
}

int cswin_seg_loss_fwd(const float* logits, const void* labels, int32_t label_bytes, float* sums, int64_t B, int32_t C, int64_t HW,
                       cswin_stream_t stream) {
  return seg_loss_fwd(logits, labels, label_bytes, sums, B, C, HW, (cudaStream_t)stream);
}

int cswin_seg_loss_bwd(const float* logits, const void* labels, int32_t label_bytes, const float* sums, const float* grad_out,
                       float* dlogits, float w_ce, float w_dice, int64_t B, int32_t C, int64_t HW, cswin_stream_t stream) {
  return seg_loss_bwd(logits, labels, label_bytes, sums, grad_out, dlogits, w_ce, w_dice, B, C, HW, (cudaStream_t)stream);
}

int cswin_sgd_momentum_step(const cswin_sgd_chunk_t* chunks, int32_t n_chunks, const float* lr, float momentum,
                            float weight_decay, cswin_stream_t stream) {
  return sgd_momentum_step(chunks, n_chunks, lr, momentum, weight_decay, (cudaStream_t)stream);
}

int cswin_mlp_fwd(const cswin_mlp_args_t* a, int32_t dtype, cswin_stream_t stream) {
  CSWIN_REQUIRE(a != nullptr, CSWIN_ERR_INVALID, "mlp_fwd: null args");
  CSWIN_REQUIRE(dtype == CSWIN_BF16, CSWIN_ERR_UNSUPPORTED, "mlp_fwd: the fused MLP exists on the bf16 / tcgen05 path only");
  CSWIN_REQUIRE(a->x && a->w1 && a->w2 && a->b1 && a->b2 && a->ln_colsum && a->ln_stats && a->out && a->M >= 0 && a->C > 0 &&
                a->hidden > 0 && a->ln_stats_parts > 0 && a->ldx >= a->C && a->ldo >= a->C && a->ldw1 >= a->C && a->ldw2 >= a->hidden,
                CSWIN_ERR_INVALID, "mlp_fwd: bad arguments");
  return mlp_fwd_tc(a, (cudaStream_t)stream);
}

int32_t cswin_mlp_stats_parts(int32_t C, int32_t hidden) { return mlp_tc_stats_parts(C, hidden); }

int cswin_qkv_lepe_attention_fwd(const cswin_qkv_attn_args_t* a, int32_t dtype, cswin_stream_t stream) {
  CSWIN_REQUIRE(a != nullptr, CSWIN_ERR_INVALID, "qkv_lepe_attention_fwd: null args");
  CSWIN_REQUIRE(dtype == CSWIN_BF16, CSWIN_ERR_UNSUPPORTED, "qkv_lepe_attention_fwd: exists on the bf16 / tcgen05 path only");
  CSWIN_REQUIRE(a->x && a->w && a->out && a->B >= 0 && a->reso > 0 && a->C > 0 && a->n_branches >= 1 && a->n_branches <= 2 &&
                a->ldw >= a->C && (a->ln_stats == nullptr || a->ln_stats_parts > 0),
                CSWIN_ERR_INVALID, "qkv_lepe_attention_fwd: bad arguments");
  return qkv_attn_fwd_tc(a, (cudaStream_t)stream);
}

int32_t cswin_qkv_lepe_attention_supported(int32_t C, int32_t reso, int32_t n_branches, const int32_t* heads, const int32_t* H_sp,
                                           const int32_t* W_sp) {
  if (heads == nullptr || H_sp == nullptr || W_sp == nullptr) return 0;
  return qkv_attn_supported(C, reso, n_branches, heads, H_sp, W_sp);
}

int cswin_stem_fwd(const void* x, int32_t x_is_f32, const void* w_packed, const float* bias, const float* gamma, const float* beta,
                   float eps, void* out, float* stats, int32_t B, int32_t H, int32_t W, int32_t dtype, cswin_stream_t stream,
                   int32_t* handled) {
  CSWIN_REQUIRE(handled != nullptr, CSWIN_ERR_INVALID, "stem_fwd: null handled");
  *handled = 0;
  CSWIN_REQUIRE(dtype == CSWIN_BF16, CSWIN_ERR_UNSUPPORTED, "stem_fwd: exists on the bf16 / tcgen05 path only");
  CSWIN_REQUIRE(x && w_packed && bias && gamma && beta && out && stats && B >= 0 && H > 0 && W > 0, CSWIN_ERR_INVALID, "stem_fwd: bad arguments");
  if (B == 0) { *handled = 1; return CSWIN_OK; }
  bool h = false;
  const int rc = stem_fwd_tc(x, x_is_f32, w_packed, bias, gamma, beta, eps, out, stats, B, H, W, (cudaStream_t)stream, &h);
  *handled = h ? 1 : 0;
  return rc;
}

int cswin_stage_plan(int32_t B, int32_t reso, int32_t C, int32_t hidden, int32_t n_branches, const int32_t* heads,
                     const int32_t* H_sp, const int32_t* W_sp, cswin_stage_plan_t* plan) {
  CSWIN_REQUIRE(heads && H_sp && W_sp && plan, CSWIN_ERR_INVALID, "stage_plan: null pointer");
  const int rc = stage_plan(B, reso, C, hidden, n_branches, heads, H_sp, W_sp, plan);
  if (rc != CSWIN_OK) set_error("stage_plan: shape outside the persistent stage kernel's envelope");
  return rc;
}

int cswin_stage_fwd(const cswin_stage_args_t* a, int32_t dtype, cswin_stream_t stream) {
  CSWIN_REQUIRE(a != nullptr, CSWIN_ERR_INVALID, "stage_fwd: null args");
  CSWIN_REQUIRE(dtype == CSWIN_BF16, CSWIN_ERR_UNSUPPORTED, "stage_fwd: exists on the bf16 / tcgen05 path only");
  CSWIN_REQUIRE(a->x && a->qkv && a->att && a->x1 && a->hid && a->blocks && a->n_blocks >= 1 && a->B >= 0 && a->reso > 0 &&
                a->n_branches >= 1 && a->n_branches <= 2, CSWIN_ERR_INVALID, "stage_fwd: bad arguments");
  if (a->B == 0) return CSWIN_OK;
  return stage_fwd_tc(a, (cudaStream_t)stream);
}

int32_t cswin_linear_stats_parts(int64_t M, int32_t N, int32_t K, int32_t act) { return linear_tc_stats_parts(M, N, K, act); }

int cswin_row_stats(const void* x, int64_t ldx, int64_t M, int32_t C, float* stats, int32_t dtype, cswin_stream_t stream) {
  CSWIN_REQUIRE(valid_dtype(dtype) && M >= 0, CSWIN_ERR_INVALID, "row_stats: bad arguments");
  return row_stats(x, ldx, M, C, stats, dtype, (cudaStream_t)stream);
}

}  // extern "C"
