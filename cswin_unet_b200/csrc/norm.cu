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

// bf16 fast path: LPR lanes per row, each lane owns NV 16-byte vectors (8 bf16) -> C = LPR * NV * 8; 32/LPR rows per warp.
// Two-pass fp32 statistics in registers, shuffle reductions inside the LPR-lane group.
template <int LPR, int NV>
__global__ void __launch_bounds__(256) layernorm_bf16_vec_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx,
                                                                  const __nv_bfloat16* __restrict__ g,
                                                                  const __nv_bfloat16* __restrict__ b,
                                                                  __nv_bfloat16* __restrict__ y, int64_t ldy, int64_t M,
                                                                  float eps, float* __restrict__ mean_out,
                                                                  float* __restrict__ rstd_out, float* __restrict__ ystats) {
  constexpr int C = LPR * NV * 8;
  constexpr int RPW = 32 / LPR;
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const int64_t row = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;
  const bool ok = row < M;
  float v[NV * 8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    uint4 u = make_uint4(0, 0, 0, 0);
    if (ok) u = *reinterpret_cast<const uint4*>(x + row * ldx + (i * LPR + sub) * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[i * 8 + 2 * e] = __uint_as_float(w[e] << 16);
      v[i * 8 + 2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
      s += v[i * 8 + 2 * e] + v[i * 8 + 2 * e + 1];
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV * 8; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.0f / C) + eps);
  float y1 = 0.f, y2 = 0.f;                             // (sum, sum^2) of the bf16 outputs, for a following folded LayerNorm
  if (ok) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c0 = (i * LPR + sub) * 8;
      const uint4 gu = *reinterpret_cast<const uint4*>(g + c0);
      const uint4 bu = *reinterpret_cast<const uint4*>(b + c0);
      const uint32_t gw[4] = {gu.x, gu.y, gu.z, gu.w}, bw[4] = {bu.x, bu.y, bu.z, bu.w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float lo = (v[i * 8 + 2 * e] - mean) * rstd * __uint_as_float(gw[e] << 16) + __uint_as_float(bw[e] << 16);
        const float hi = (v[i * 8 + 2 * e + 1] - mean) * rstd * __uint_as_float(gw[e] & 0xffff0000u) + __uint_as_float(bw[e] & 0xffff0000u);
        const __nv_bfloat162 pk = __floats2bfloat162_rn(lo, hi);
        o[e] = *reinterpret_cast<const uint32_t*>(&pk);
        const float ylo = __uint_as_float(o[e] << 16), yhi = __uint_as_float(o[e] & 0xffff0000u);
        y1 += ylo + yhi;
        y2 = fmaf(ylo, ylo, fmaf(yhi, yhi, y2));
      }
      *reinterpret_cast<uint4*>(y + row * ldy + c0) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (sub == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
  }
  if (ystats != nullptr) {                              // uniform branch: every lane takes part in the shuffles
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) { y1 += __shfl_xor_sync(0xffffffffu, y1, o); y2 += __shfl_xor_sync(0xffffffffu, y2, o); }
    if (ok && sub == 0) { ystats[row * 2] = y1; ystats[row * 2 + 1] = y2; }
  }
}

template <int LPR, int NV>
int launch_vec(const void* x, int64_t ldx, const void* g, const void* b, void* y, int64_t ldy, int64_t M, float eps,
               float* mean, float* rstd, float* ystats, cudaStream_t s) {
  constexpr int RPW = 32 / LPR;
  const int64_t rows_per_cta = 8 * RPW;
  CSWIN_CUDA_OK(launch_pdl(layernorm_bf16_vec_kernel<LPR, NV>, dim3((unsigned)ceil_div64(M, rows_per_cta)), dim3(256), 0, s,
                           (const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)g, (const __nv_bfloat16*)b, (__nv_bfloat16*)y,
                           ldy, M, eps, mean, rstd, ystats));
  CSWIN_LAUNCH_CHECK();
  return CSWIN_OK;
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

template <typename T>
__global__ void __launch_bounds__(256) row_stats_kernel(const T* __restrict__ x, int64_t ldx, int64_t M, int C,
                                                         float* __restrict__ stats) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < C; c += 32) { const float v = ldf(x + row * ldx + c); s1 += v; s2 = fmaf(v, v, s2); }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if (lane == 0) { stats[row * 2] = s1; stats[row * 2 + 1] = s2; }
}

}  // namespace

int row_stats(const void* x, int64_t ldx, int64_t M, int C, float* stats, int dtype, cudaStream_t s) {
  CSWIN_REQUIRE(x && stats && C > 0 && ldx >= C, CSWIN_ERR_INVALID, "row_stats: bad arguments");
  if (M == 0) return CSWIN_OK;
  const dim3 grid((unsigned)ceil_div64(M, 8));
  if (dtype == CSWIN_F32) CSWIN_CUDA_OK(launch_pdl(row_stats_kernel<float>, grid, dim3(256), 0, s, (const float*)x, ldx, M, C, stats));
  else CSWIN_CUDA_OK(launch_pdl(row_stats_kernel<__nv_bfloat16>, grid, dim3(256), 0, s, (const __nv_bfloat16*)x, ldx, M, C, stats));
  CSWIN_LAUNCH_CHECK();
  return CSWIN_OK;
}

int layernorm_fwd(const void* x, int64_t ldx, const void* g, const void* b, void* y, int64_t ldy, int64_t M, int C,
                  float eps, float* mean, float* rstd, float* ystats, int dtype, cudaStream_t s) {
  CSWIN_REQUIRE(x && g && b && y, CSWIN_ERR_INVALID, "layernorm: null pointer");
  CSWIN_REQUIRE(C > 0 && C <= 2048, CSWIN_ERR_UNSUPPORTED, "layernorm: C=%d outside (0, 2048]", C);
  CSWIN_REQUIRE(ldx >= C && ldy >= C, CSWIN_ERR_INVALID, "layernorm: leading dimension smaller than C");
  if (M == 0) return CSWIN_OK;
  if (dtype == CSWIN_F32) {
    CSWIN_REQUIRE(!ystats, CSWIN_ERR_UNSUPPORTED, "layernorm: output row statistics exist on the vectorised bf16 path only");
    return launch<float>(x, ldx, g, b, y, ldy, M, C, eps, mean, rstd, s);
  }
  const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(g) |
                     reinterpret_cast<uintptr_t>(b)) % 16 == 0) && (ldx * 2) % 16 == 0 && (ldy * 2) % 16 == 0;
  if (vec && C == 64) return launch_vec<8, 1>(x, ldx, g, b, y, ldy, M, eps, mean, rstd, ystats, s);
  if (vec && C == 128) return launch_vec<16, 1>(x, ldx, g, b, y, ldy, M, eps, mean, rstd, ystats, s);
  if (vec && C == 256) return launch_vec<32, 1>(x, ldx, g, b, y, ldy, M, eps, mean, rstd, ystats, s);
  if (vec && C == 512) return launch_vec<32, 2>(x, ldx, g, b, y, ldy, M, eps, mean, rstd, ystats, s);
  CSWIN_REQUIRE(!ystats, CSWIN_ERR_UNSUPPORTED, "layernorm: output row statistics need C in {64,128,256,512} and 16-byte aligned rows");
  return launch<__nv_bfloat16>(x, ldx, g, b, y, ldy, M, C, eps, mean, rstd, s);
}

}  // namespace cswin
