// common.cuh — shared helpers for libcswin_b200.so (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/cswin_b200.h"

namespace cswin {

// ---- error plumbing (thread-local message, never throws across the ABI) ----
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
extern std::atomic<uint64_t> g_tc_launches;
extern std::atomic<int> g_gemm_smem_cap_kb;        // cswin_set_option(CSWIN_OPT_GEMM_SMEM_CAP_KB): per-CTA smem ceiling of the tcgen05 Linear (0 = none)
extern std::atomic<unsigned long long*> g_trace;   // debug: device buffer for in-kernel %globaltimer stamps (or null)

#define CSWIN_REQUIRE(cond, code, ...)            \
  do {                                            \
    if (!(cond)) {                                \
      ::cswin::set_error(__VA_ARGS__);            \
      return (code);                              \
    }                                             \
  } while (0)

#define CSWIN_CUDA_OK(expr)                                                               \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::cswin::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CSWIN_ERR_CUDA;                                                              \
    }                                                                                     \
  } while (0)

// after a <<<>>> launch: count it and surface launch-configuration errors
#define CSWIN_LAUNCH_CHECK()                      \
  do {                                            \
    ::cswin::g_launches.fetch_add(1, std::memory_order_relaxed); \
    CSWIN_CUDA_OK(cudaGetLastError());            \
  } while (0)

int sm_count();   // SM count of the current device (cached)

// ---- programmatic dependent launch (PDL) ----
// Every bf16 hot-path kernel starts with pdl_trigger() (lets the NEXT kernel in the stream begin its prologue: barrier
// init, TMEM allocation, descriptor prefetch, weight staging) and executes pdl_wait() in EVERY thread before it touches
// memory that an earlier kernel may still be writing or reading (activations in, outputs out).  Weights / biases are
// never written on this path, so they may be fetched before the wait.  CSWIN_PDL=0 disables the launch attribute.
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#endif

// ---- element access: T in {float, __nv_bfloat16}, arithmetic always fp32 ----
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- per-op launchers (defined in the .cu files, called by api.cu) ----
int lepe_attention_fwd_simt(const cswin_lepe_branch_t* br, int nb, int B, int reso, float scale, int dtype, cudaStream_t s);
int lepe_attention_fwd_tc(const cswin_lepe_branch_t* br, int nb, int B, int reso, float scale, cudaStream_t s, bool* handled);
int lepe_attention_bwd_tc(const cswin_lepe_branch_grad_t* br, int nb, int B, int reso, float scale, cudaStream_t s, bool* handled);
int lepe_attention_bwd_simt(const cswin_lepe_branch_grad_t* br, int nb, int B, int reso, float scale, int dtype, cudaStream_t s);
int zoom_cubic(const float* in, int n, int H, int W, double* work, float* out, int64_t out_ns, int64_t out_cs, int reps, int OH,
               int OW, cudaStream_t s);
int zoom_nearest_u8(const uint8_t* in, int n, int H, int W, uint8_t* out, int OH, int OW, cudaStream_t s);
int lepe_param_grad_tc(const cswin_lepe_branch_grad_t* br, int nb, int B, int reso, cudaStream_t s, bool* handled);
int layernorm_fwd(const void* x, int64_t ldx, const void* g, const void* b, void* y, int64_t ldy, int64_t M, int C,
                  float eps, float* mean, float* rstd, float* ystats, int dtype, cudaStream_t s);
int row_stats(const void* x, int64_t ldx, int64_t M, int C, float* stats, int dtype, cudaStream_t s);
int carafe_head_bwd(const void* enc, int64_t ldenc, const void* z, int64_t ldz, const float* dlogits, void* denc, int64_t lddenc,
                    void* dz, int64_t lddz, int zcols, float* dbias, float* kws, int B, int H, int W, int C, int up, int dtype,
                    cudaStream_t s);
int seg_loss_fwd(const float* logits, const void* labels, int label_bytes, float* sums, int64_t B, int C, int64_t HW, cudaStream_t s);
int seg_loss_bwd(const float* logits, const void* labels, int label_bytes, const float* sums, const float* gout, float* dlogits,
                 float w_ce, float w_dice, int64_t B, int C, int64_t HW, cudaStream_t s);
int sgd_momentum_step(const cswin_sgd_chunk_t* chunks, int n_chunks, const float* lr, float momentum, float wd, cudaStream_t s);
int mlp_fwd_tc(const cswin_mlp_args_t* a, cudaStream_t stream);
int mlp_tc_stats_parts(int C, int hidden);
int qkv_attn_fwd_tc(const cswin_qkv_attn_args_t* a, cudaStream_t stream);
int qkv_attn_supported(int C, int reso, int nb, const int* heads, const int* hs, const int* ws);
int linear_tc_stats_parts(int64_t M, int N, int K, int act);
int stage_plan(int B, int reso, int C, int hidden, int nb, const int* heads, const int* hs, const int* ws, cswin_stage_plan_t* out);
int stage_fwd_tc(const cswin_stage_args_t* a, cudaStream_t stream);
int stem_fwd_tc(const void* x, int x_is_f32, const void* w_packed, const float* bias, const float* gamma, const float* beta, float eps,
                void* out, float* stats, int B, int H, int W, cudaStream_t stream, bool* handled);
int linear_fwd_simt(const cswin_linear_args_t* a, int dtype, cudaStream_t s);
int linear_fwd_tc(const cswin_linear_args_t* a, cudaStream_t s, bool* handled);
int conv_tokens_fwd_tc(const void* x, int64_t x_bs, int64_t x_ts, const void* w, int64_t ldw, const void* bias, void* out, int64_t ldo,
                       int B, int H, int W, int C, int N, int KH, int KW, int stride, int pad, cudaStream_t stream, bool* handled);
int im2col_tokens(const void* x, int64_t x_bs, int64_t x_ts, void* col, int64_t ldcol, int B, int H, int W, int C, int KH,
                  int KW, int stride, int pad, int dtype, cudaStream_t s);
int im2col_nchw(const void* x, int x_is_f32, void* col, int64_t ldcol, int B, int C, int H, int W, int KH, int KW,
                int stride, int pad, int dtype, cudaStream_t s);
int carafe_reassemble_fwd(const void* enc, int64_t ldenc, const void* z, int64_t ldz, const void* bias, void* y,
                          int64_t ldy, int nchw_out, int y_is_f32, int B, int H, int W, int C, int up, int dtype,
                          cudaStream_t s);

int act_fwd(const void* z, int64_t ldz, void* out, int64_t ldo, int64_t M, int N, int act, int dtype, cudaStream_t s);
int act_bwd(const void* dout, int64_t ldd, const void* z, int64_t ldz, const float* sscale, int rps, void* dz, int64_t ldo,
            int64_t M, int N, int act, int dtype, cudaStream_t s);
int linear_wgrad(const void* dz, int64_t ldz, const void* a, int64_t lda, float* dw, int64_t ldw, float* db, int64_t M,
                 int N, int K, int dtype, cudaStream_t s);
int linear_wgrad_tc(const void* dz, int64_t ldz, const void* a, int64_t lda, float* dw, int64_t ldw, float* db, int64_t M,
                    int N, int K, cudaStream_t s, bool* handled);
int layernorm_bwd(const void* x, int64_t ldx, const void* dy, int64_t ldy, const void* gamma, const float* mean,
                  const float* rstd, void* dx, int64_t ldo, const void* dx_add, int64_t lda, float* dgamma, float* dbeta,
                  int64_t M, int C, int dtype, cudaStream_t s);
int col2im_tokens(const void* dcol, int64_t ldcol, void* dx, int64_t x_bs, int64_t x_ts, int B, int H, int W, int C, int KH,
                  int KW, int stride, int pad, int dtype, cudaStream_t s);
int carafe_reassemble_bwd(const void* enc, int64_t ldenc, const void* z, int64_t ldz, const void* dy, int dy_is_f32,
                          int64_t sb, int64_t sy, int64_t sx, int64_t sc, void* denc, int64_t lddenc, void* dz, int64_t lddz,
                          float* dbias, float* kws, int B, int H, int W, int C, int up, int dtype, cudaStream_t s);
int carafe_head_fwd(const void* enc, int64_t ldenc, const void* z, int64_t ldz, const void* bias, void* logits,
                    int logits_is_f32, uint8_t* labels, int B, int H, int W, int C, int up, int dtype, cudaStream_t s);

}  // namespace cswin
