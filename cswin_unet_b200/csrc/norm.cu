// norm.cu — LayerNorm over the channel dimension of token-major activations.
// Replaces nn.LayerNorm at networks/cswin_unet.py:168,179 (norm1/norm2), :218 (Merge_Block.norm), :341 (stem),
// :497 (norm), :533 (norm_up).  One warp per token row, two-pass fp32 statistics held in registers when the row
// fits (C <= 1024), HBM-bound: reads x once, writes y once.
#include "common.cuh"

namespace cswin {
namespace {

constexpr int kRowsPerCta = 8;     // 8 warps

template <typename T, int VPL>     // VPL = values per lane held in registers (C <= 32*VPL)
__global__ void __launch_bounds__(kRowsPerCta * 32) layernorm_kernel(const T* __restrict__ x, int64_t ldx,
                                                                      const T* __restrict__ g, const T* __restrict__ b,
                                                                      T* __restrict__ y, int64_t ldy, int64_t M, int C,
                                                                      float eps, float* __restrict__ mean_out,
                                                                      float* __restrict__ rstd_out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kRowsPerCta + (threadIdx.x >> 5);
  if (row >= M) return;
  const T* xr = x + row * ldx;
  float v[VPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = lane + 32 * i;
    v[i] = (c < C) ? ldf(xr + c) : 0.f;
    s += v[i];
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = lane + 32 * i;
    const float dlt = (c < C) ? v[i] - mean : 0.f;
    q = fmaf(dlt, dlt, q);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  T* yr = y + row * ldy;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = lane + 32 * i;
    if (c < C) stf(yr + c, (v[i] - mean) * rstd * ldf(g + c) + ldf(b + c));
  }
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

template <typename T>
int launch(const void* x, int64_t ldx, const void* g, const void* b, void* y, int64_t ldy, int64_t M, int C, float eps,
           float* mean, float* rstd, cudaStream_t s) {
  const unsigned grid = (unsigned)ceil_div64(M, kRowsPerCta);
#define LN_CASE(V)                                                                                               \
  layernorm_kernel<T, V><<<grid, kRowsPerCta * 32, 0, s>>>((const T*)x, ldx, (const T*)g, (const T*)b, (T*)y, ldy, M, C, \
                                                            eps, mean, rstd)
  if (C <= 64) LN_CASE(2);
  else if (C <= 128) LN_CASE(4);
  else if (C <= 256) LN_CASE(8);
  else if (C <= 512) LN_CASE(16);
  else LN_CASE(64);
#undef LN_CASE
  CSWIN_LAUNCH_CHECK();
  return CSWIN_OK;
}

}  // namespace

int layernorm_fwd(const void* x, int64_t ldx, const void* g, const void* b, void* y, int64_t ldy, int64_t M, int C,
                  float eps, float* mean, float* rstd, int dtype, cudaStream_t s) {
  CSWIN_REQUIRE(x && g && b && y, CSWIN_ERR_INVALID, "layernorm: null pointer");
  CSWIN_REQUIRE(C > 0 && C <= 2048, CSWIN_ERR_UNSUPPORTED, "layernorm: C=%d outside (0, 2048]", C);
  CSWIN_REQUIRE(ldx >= C && ldy >= C, CSWIN_ERR_INVALID, "layernorm: leading dimension smaller than C");
  if (M == 0) return CSWIN_OK;
  if (dtype == CSWIN_F32) return launch<float>(x, ldx, g, b, y, ldy, M, C, eps, mean, rstd, s);
  return launch<__nv_bfloat16>(x, ldx, g, b, y, ldy, M, C, eps, mean, rstd, s);
}

}  // namespace cswin
