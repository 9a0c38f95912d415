// backward.cu — backward kernels of the hot path other than attention (fp32 arithmetic, fp32 or bf16 storage).
// The reference has no backward source: these are the autograd derivatives of networks/cswin_unet.py:160-181 (block),
// :211-220 (Merge_Block), :232-269 / :282-319 (CARAFE) — checked against torch.autograd on the CPU oracle.
//
//   act_bwd              dZ = dOut * sample_scale * act'(Z)              (GELU(erf) derivative; DropPath scale)
//   linear_wgrad         dW[N,K] += dZ^T A,  db[N] += colsum(dZ)         (fp32 accumulation into fp32 gradients)
//   (linear dgrad = cswin_linear_fwd with the transposed weight: dA = dZ W)
//   layernorm_bwd        dx, d gamma, d beta from saved mean / rstd
//   col2im_tokens        adjoint of im2col_tokens (gather form, no atomics)
//   carafe_reassemble_bwd  d enc (through the 9-tap softmax), d z, d bias
#include "common.cuh"
#include "tc_common.cuh"

namespace cswin {
namespace {

// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) act_bwd_kernel(const T* __restrict__ dout, int64_t ldd, const T* __restrict__ z,
                                                       int64_t ldz, const float* __restrict__ sscale, int rps,
                                                       T* __restrict__ dz, int64_t ldo, int64_t M, int N, int act) {
  const int64_t total = M * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / N;
    const int n = (int)(i - m * N);
    float g = ldf(dout + m * ldd + n);
    if (sscale != nullptr) g *= sscale[m / rps];
    if (act == 1) {
      const float x = ldf(z + m * ldz + n);
      const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
      const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
      g *= cdf + x * pdf;
    }
    stf(dz + m * ldo + n, g);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) act_fwd_kernel(const T* __restrict__ z, int64_t ldz, T* __restrict__ out, int64_t ldo,
                                                       int64_t M, int N, int act) {
  const int64_t total = M * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / N;
    const int n = (int)(i - m * N);
    const float x = ldf(z + m * ldz + n);
    stf(out + m * ldo + n, act == 1 ? gelu_erf(x) : x);
  }
}

// bf16, contiguous rows (ld == N), N % 8 == 0: 16-byte vectors, 8 elements per thread
// bf16 path: GELU'(x) = tc::gelu_grad_fast (clamped tanh-form fit of Phi + ex2.approx for phi; tc_common.cuh)
using tc::gelu_grad_fast;
__global__ void __launch_bounds__(256) act_bwd_bf16_vec_kernel(const uint4* __restrict__ dout, const uint4* __restrict__ z,
                                                                const float* __restrict__ sscale, int64_t elems_per_sample,
                                                                uint4* __restrict__ dz, int64_t nvec, int act) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 d = dout[i];
    const float sc = sscale != nullptr ? sscale[(i * 8) / elems_per_sample] : 1.0f;
    uint32_t dw[4] = {d.x, d.y, d.z, d.w}, zw[4] = {0, 0, 0, 0}, o[4];
    if (act == 1) { const uint4 zz = z[i]; zw[0] = zz.x; zw[1] = zz.y; zw[2] = zz.z; zw[3] = zz.w; }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float lo = __uint_as_float(dw[e] << 16) * sc, hi = __uint_as_float(dw[e] & 0xffff0000u) * sc;
      if (act == 1) { lo *= gelu_grad_fast(__uint_as_float(zw[e] << 16)); hi *= gelu_grad_fast(__uint_as_float(zw[e] & 0xffff0000u)); }
      const __nv_bfloat162 pk = __floats2bfloat162_rn(lo, hi);
      o[e] = *reinterpret_cast<const uint32_t*>(&pk);
    }
    dz[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
__global__ void __launch_bounds__(256) act_fwd_bf16_vec_kernel(const uint4* __restrict__ z, uint4* __restrict__ out, int64_t nvec) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 zz = z[i];
    const uint32_t zw[4] = {zz.x, zz.y, zz.z, zz.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat162 pk = __floats2bfloat162_rn(tc::gelu_fast(__uint_as_float(zw[e] << 16)), tc::gelu_fast(__uint_as_float(zw[e] & 0xffff0000u)));
      o[e] = *reinterpret_cast<const uint32_t*>(&pk);
    }
    out[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// dW[n,k] += sum_m dZ[m,n] A[m,k] ; 64x64 (n,k) tile per CTA, M split over gridDim.z, fp32 atomics at the end.
constexpr int WT = 64, WM = 16;
template <typename T>
__global__ void __launch_bounds__(256) linear_wgrad_kernel(const T* __restrict__ dz, int64_t ldz, const T* __restrict__ a,
                                                            int64_t lda, float* __restrict__ dw, int64_t ldw,
                                                            float* __restrict__ db, int64_t M, int N, int K, int64_t mchunk) {
  __shared__ float Zs[WM][WT + 4];
  __shared__ float As[WM][WT + 4];
  const int n0 = blockIdx.x * WT, k0 = blockIdx.y * WT;
  const int64_t mbeg = (int64_t)blockIdx.z * mchunk;
  const int64_t mend = mbeg + mchunk < M ? mbeg + mchunk : M;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  const int lr = tid >> 4;          // 0..15 row of the m-chunk
  const int lc = (tid & 15) * 4;    // 0..60 first of 4 columns
  for (int64_t m0 = mbeg; m0 < mend; m0 += WM) {
    const int64_t m = m0 + lr;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int n = n0 + lc + e, k = k0 + lc + e;
      Zs[lr][lc + e] = (m < mend && n < N) ? ldf(dz + m * ldz + n) : 0.f;
      As[lr][lc + e] = (m < mend && k < K) ? ldf(a + m * lda + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < WM; ++mm) {
      const float4 zv = *reinterpret_cast<const float4*>(&Zs[mm][ty * 4]);
      const float4 av = *reinterpret_cast<const float4*>(&As[mm][tx * 4]);
      const float z4[4] = {zv.x, zv.y, zv.z, zv.w}, a4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        bsum[i] += z4[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(z4[i], a4[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K) atomicAdd(dw + (int64_t)n * ldw + k, acc[i][j]);
    }
    if (db != nullptr && blockIdx.y == 0 && tx == 0) atomicAdd(db + n, bsum[i]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
template <typename T, int VPL>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy,
                                                             int64_t ldy, const T* __restrict__ gamma,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             T* __restrict__ dx, int64_t ldo, const T* __restrict__ dx_add,
                                                             int64_t lda, float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, int64_t M, int C) {
  __shared__ float red[8][32 * VPL * 2];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float dg[VPL], dbt[VPL], gm[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) { dg[i] = 0.f; dbt[i] = 0.f; const int c = lane + 32 * i; gm[i] = c < C ? ldf(gamma + c) : 0.f; }
  for (int64_t row = (int64_t)blockIdx.x * 8 + w; row < M; row += (int64_t)gridDim.x * 8) {
    const float mu = mean[row], rs = rstd[row];
    float xh[VPL], g[VPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = lane + 32 * i;
      xh[i] = 0.f; g[i] = 0.f;
      if (c < C) {
        const float d = ldf(dy + row * ldy + c);
        xh[i] = (ldf(x + row * ldx + c) - mu) * rs;
        g[i] = d * gm[i];
        dg[i] = fmaf(d, xh[i], dg[i]);
        dbt[i] += d;
        s1 += g[i];
        s2 = fmaf(g[i], xh[i], s2);
      }
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = lane + 32 * i;
      if (c < C) stf(dx + row * ldo + c, rs * (g[i] - s1 - xh[i] * s2) + (dx_add != nullptr ? ldf(dx_add + row * lda + c) : 0.f));
    }
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) { red[w][(lane + 32 * i) * 2] = dg[i]; red[w][(lane + 32 * i) * 2 + 1] = dbt[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) { a += red[ww][c * 2]; b += red[ww][c * 2 + 1]; }
    atomicAdd(dgamma + c, a);
    atomicAdd(dbeta + c, b);
  }
}

// bf16 rows of C = LPR * NV * 8 channels: LPR lanes per row, NV 16-byte vectors per lane, 32 / LPR rows per warp in flight.
// dgamma / dbeta are carried in registers across the rows of a warp, reduced through shared memory, one atomic per channel
// per CTA.
template <int LPR, int NV>
__global__ void __launch_bounds__(256) layernorm_bwd_bf16_vec_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx,
                                                                      const __nv_bfloat16* __restrict__ dy, int64_t ldy,
                                                                      const __nv_bfloat16* __restrict__ gamma,
                                                                      const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                      __nv_bfloat16* __restrict__ dx, int64_t ldo,
                                                                      const __nv_bfloat16* __restrict__ dx_add, int64_t lda,
                                                                      float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t M) {
  pdl_trigger();
  pdl_wait();
  constexpr int C = LPR * NV * 8, RPW = 32 / LPR;
  __shared__ float red[8][C * 2];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, sub = lane / LPR, l = lane % LPR;
  float gm[NV][8], dg[NV][8], dbt[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const uint4 u = *reinterpret_cast<const uint4*>(gamma + (v * LPR + l) * 8);
    const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      gm[v][e] = __uint_as_float((e & 1) ? (uw[e >> 1] & 0xffff0000u) : (uw[e >> 1] << 16));
      dg[v][e] = 0.f; dbt[v][e] = 0.f;
    }
  }
  const float invC = 1.0f / (float)C;
  for (int64_t base = ((int64_t)blockIdx.x * 8 + w) * RPW; base < M; base += (int64_t)gridDim.x * 8 * RPW) {
    const int64_t row = base + sub;
    const bool valid = row < M;
    const float mu = valid ? mean[row] : 0.f, rs = valid ? rstd[row] : 0.f;
    float xh[NV][8], g[NV][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      uint4 ux = make_uint4(0, 0, 0, 0), ud = make_uint4(0, 0, 0, 0);
      if (valid) {
        ux = *reinterpret_cast<const uint4*>(x + row * ldx + (v * LPR + l) * 8);
        ud = *reinterpret_cast<const uint4*>(dy + row * ldy + (v * LPR + l) * 8);
      }
      const uint32_t xw[4] = {ux.x, ux.y, ux.z, ux.w}, dw[4] = {ud.x, ud.y, ud.z, ud.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xv = __uint_as_float((e & 1) ? (xw[e >> 1] & 0xffff0000u) : (xw[e >> 1] << 16));
        const float d = __uint_as_float((e & 1) ? (dw[e >> 1] & 0xffff0000u) : (dw[e >> 1] << 16));
        xh[v][e] = (xv - mu) * rs;
        g[v][e] = d * gm[v][e];
        dg[v][e] = fmaf(d, xh[v][e], dg[v][e]);
        dbt[v][e] += d;
        s1 += g[v][e];
        s2 = fmaf(g[v][e], xh[v][e], s2);
      }
    }
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    s1 *= invC; s2 *= invC;
    if (valid) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        uint32_t o[4];
        uint4 ua = make_uint4(0, 0, 0, 0);                 // gradient that by-passed the LayerNorm through the residual add
        if (dx_add != nullptr) ua = *reinterpret_cast<const uint4*>(dx_add + row * lda + (v * LPR + l) * 8);
        const uint32_t aw[4] = {ua.x, ua.y, ua.z, ua.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(rs * (g[v][2 * e] - s1 - xh[v][2 * e] * s2) + __uint_as_float(aw[e] << 16),
                                                          rs * (g[v][2 * e + 1] - s1 - xh[v][2 * e + 1] * s2) + __uint_as_float(aw[e] & 0xffff0000u));
          o[e] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        *reinterpret_cast<uint4*>(dx + row * ldo + (v * LPR + l) * 8) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
#pragma unroll
      for (int o = LPR; o < 32; o <<= 1) {
        dg[v][e] += __shfl_xor_sync(0xffffffffu, dg[v][e], o);
        dbt[v][e] += __shfl_xor_sync(0xffffffffu, dbt[v][e], o);
      }
      if (sub == 0) { red[w][((v * LPR + l) * 8 + e) * 2] = dg[v][e]; red[w][((v * LPR + l) * 8 + e) * 2 + 1] = dbt[v][e]; }
    }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) { a += red[ww][c * 2]; b += red[ww][c * 2 + 1]; }
    atomicAdd(dgamma + c, a);
    atomicAdd(dbeta + c, b);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// dx[b, iy, ix, c] = sum over taps (ky,kx) with (iy + pad - ky) % stride == 0 ... of dcol[(b,oy,ox), (ky*KW+kx)*C + c]
template <typename T>
__global__ void __launch_bounds__(256) col2im_tokens_kernel(const T* __restrict__ dcol, int64_t ldcol, T* __restrict__ dx,
                                                             int64_t x_bs, int64_t x_ts, int B, int H, int W, int C, int KH,
                                                             int KW, int stride, int pad, int Ho, int Wo) {
  pdl_trigger();
  pdl_wait();
  const int64_t total = (int64_t)B * H * W * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int ix = (int)(r % W); r /= W;
    const int iy = (int)(r % H);
    const int b = (int)(r / H);
    float acc = 0.f;
    for (int ky = 0; ky < KH; ++ky) {
      const int ty = iy + pad - ky;
      if (ty < 0 || ty % stride) continue;
      const int oy = ty / stride;
      if (oy >= Ho) continue;
      for (int kx = 0; kx < KW; ++kx) {
        const int tx = ix + pad - kx;
        if (tx < 0 || tx % stride) continue;
        const int ox = tx / stride;
        if (ox >= Wo) continue;
        acc += ldf(dcol + ((int64_t)(b * Ho + oy) * Wo + ox) * ldcol + (int64_t)(ky * KW + kx) * C + c);
      }
    }
    stf(dx + (int64_t)b * x_bs + ((int64_t)iy * W + ix) * x_ts + c, acc);
  }
}

// bf16, 8 channels (16 bytes) per thread: fp32 accumulation over the <= ceil(KH/stride) * ceil(KW/stride) taps that hit the pixel
__global__ void __launch_bounds__(256) col2im_tokens_bf16_vec_kernel(const __nv_bfloat16* __restrict__ dcol, int64_t ldcol,
                                                                      __nv_bfloat16* __restrict__ dx, int64_t x_bs, int64_t x_ts,
                                                                      int B, int H, int W, int C, int KH, int KW, int stride,
                                                                      int pad, int Ho, int Wo) {
  pdl_trigger();
  pdl_wait();
  const int C8 = C >> 3;
  const int64_t total = (int64_t)B * H * W * C8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    int64_t r = i / C8;
    const int ix = (int)(r % W); r /= W;
    const int iy = (int)(r % H);
    const int b = (int)(r / H);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int ky = 0; ky < KH; ++ky) {
      const int ty = iy + pad - ky;
      if (ty < 0 || ty % stride) continue;
      const int oy = ty / stride;
      if (oy >= Ho) continue;
      for (int kx = 0; kx < KW; ++kx) {
        const int tx = ix + pad - kx;
        if (tx < 0 || tx % stride) continue;
        const int ox = tx / stride;
        if (ox >= Wo) continue;
        const uint4 u = *reinterpret_cast<const uint4*>(dcol + ((int64_t)(b * Ho + oy) * Wo + ox) * ldcol + (int64_t)(ky * KW + kx) * C + c8 * 8);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) { acc[2 * e] += __uint_as_float(w[e] << 16); acc[2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u); }
      }
    }
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[2 * e], acc[2 * e + 1]);
      o[e] = *reinterpret_cast<const uint32_t*>(&h2);
    }
    *reinterpret_cast<uint4*>(dx + (int64_t)b * x_bs + ((int64_t)iy * W + ix) * x_ts + c8 * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CARAFE re-assembly backward.  dy element (b, oy, ox, c) lives at dy + b*sb + oy*sy + ox*sx + c*sc (token-major or NCHW).
// Kernel A (one warp per low-res pixel): kappa (recomputed), d kappa[ae][t] = <dy[ae], z^[t]>, d enc through the softmax,
//          kappa saved to `kws` (pix, s2, 9) for kernel B, d bias partial sums.
// Kernel B (one warp per low-res pixel p'): dz[p'] = sum_t sum_ae kappa[p'-off(t)][ae][t] * dy[p'-off(t), ae]  (gather).
// Kernel A.  One warp per low-res pixel, persistent over pixels.  Lane = (sub-pixel ae, channel group g): s2 = up^2
// sub-pixels x G = 32/s2 groups of CPL = C/G channels.  Each lane reads its CPL channels of dy[ae] once (vector loads),
// forms its share of d kappa[ae][t] = <dy[ae], z^[t]> against the 9 neighbour rows of z, the G partial sums of a
// sub-pixel are combined with log2(G) shuffles, and the lane with g == 0 applies the softmax backward and writes d enc and
// kappa.  d bias is accumulated in registers across all pixels of the warp and flushed once.
template <typename T, typename TG, int CPL>
__global__ void __launch_bounds__(256) carafe_bwd_a_kernel(const T* __restrict__ enc, int64_t ldenc, const T* __restrict__ z,
                                                            int64_t ldz, const TG* __restrict__ dy, int64_t sb, int64_t sy,
                                                            int64_t sx, int64_t sc, T* __restrict__ denc, int64_t lddenc,
                                                            float* __restrict__ kws, float* __restrict__ dbias, int64_t npix,
                                                            int H, int W, int C, int up) {
  extern __shared__ float sdb[];                 // [C] per-CTA d bias
  const int lane = threadIdx.x & 31;
  const int s2 = up * up, G = 32 / s2;
  const int ae = lane / G, g = lane - ae * G;
  const int c0 = g * CPL;
  for (int c = threadIdx.x; c < C; c += blockDim.x) sdb[c] = 0.f;
  __syncthreads();
  float dbacc[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) dbacc[j] = 0.f;
  const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t pix = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pix < npix; pix += wstride) {
    const int x0 = (int)(pix % W), y0 = (int)((pix / W) % H);
    const int64_t b = pix / ((int64_t)W * H);
    const int oy = y0 * up + ae / up, ox = x0 * up + ae % up;
    float gv[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) { gv[j] = ldf(dy + b * sb + oy * sy + ox * sx + (int64_t)(c0 + j) * sc); dbacc[j] += gv[j]; }
    float dk[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      dk[t] = 0.f;
      const int yy = y0 + t / 3 - 1, xx = x0 + t % 3 - 1;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
        const T* zr = z + ((b * H + yy) * W + xx) * ldz + c0;
#pragma unroll
        for (int j = 0; j < CPL; ++j) dk[t] = fmaf(gv[j], ldf(zr + j), dk[t]);
      }
    }
    for (int o = G >> 1; o > 0; o >>= 1)
#pragma unroll
      for (int t = 0; t < 9; ++t) dk[t] += __shfl_xor_sync(0xffffffffu, dk[t], o);
    if (g == 0) {
      float k[9];
      float mx = -INFINITY;
#pragma unroll
      for (int t = 0; t < 9; ++t) { k[t] = ldf(enc + pix * ldenc + t * s2 + ae); mx = fmaxf(mx, k[t]); }
      float sum = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) { k[t] = expf(k[t] - mx); sum += k[t]; }
      const float inv = 1.0f / sum;
      float dot = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) { k[t] *= inv; dot = fmaf(k[t], dk[t], dot); }
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        stf(denc + pix * lddenc + t * s2 + ae, k[t] * (dk[t] - dot));
        kws[(pix * s2 + ae) * 9 + t] = k[t];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < CPL; ++j) atomicAdd(&sdb[c0 + j], dbacc[j]);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) if (sdb[c] != 0.f) atomicAdd(dbias + c, sdb[c]);
}

// Kernel B (gather).  One warp per low-res pixel p', lane owns V = C/32 channels; UP is a template parameter so that the
// 9 x UP^2 (kappa, dy) pairs are fully unrolled and their loads overlap.
template <typename T, typename TG, int UP, int V>
__global__ void __launch_bounds__(256) carafe_bwd_b_kernel(const TG* __restrict__ dy, int64_t sb, int64_t sy, int64_t sx,
                                                            int64_t sc, const float* __restrict__ kws, T* __restrict__ dz,
                                                            int64_t lddz, int64_t npix, int H, int W, int C) {
  constexpr int s2 = UP * UP;
  const int lane = threadIdx.x & 31;
  const int64_t pix = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pix >= npix) return;
  const int x0 = (int)(pix % W), y0 = (int)((pix / W) % H);
  const int64_t b = pix / ((int64_t)W * H);
  float acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int py = y0 - (t / 3 - 1), px = x0 - (t % 3 - 1);          // the pixel whose tap t reads (y0, x0)
    if (py < 0 || py >= H || px < 0 || px >= W) continue;
    const int64_t pp = (b * H + py) * W + px;
#pragma unroll
    for (int ae = 0; ae < s2; ++ae) {
      const float kv = kws[(pp * s2 + ae) * 9 + t];
      const TG* src = dy + b * sb + (int64_t)(py * UP + ae / UP) * sy + (int64_t)(px * UP + ae % UP) * sx + (int64_t)(lane * V) * sc;
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = fmaf(kv, ldf(src + j * sc), acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < V; ++j) stf(dz + pix * lddz + lane * V + j, acc[j]);
}

// ---------------------------------------------------------------------------------------------------------------
// Folded segmentation head backward (up = 4, NC classes, fp32 NCHW d logits).  Same work split as the forward head kernel:
// Kernel A: one warp per (batch, low-res row, 8 low-res pixels), lane = (pixel, output sub-row) owning 4 output pixels:
//           kappa recomputed from enc, d kappa = <d logits, z^>, softmax backward -> d enc; kappa parked tap-major for
//           kernel B; d bias reduced warp -> CTA -> one atomic per class per CTA.
// Kernel B: one warp per low-res pixel p', lane = (sub-pixel ae, class half): d z[p'] = sum_t sum_ae kappa[p'-off(t)][t][ae]
//           * d logits[p'-off(t), ae]  (gather; no atomics), zero in the padded columns.
template <typename T> __device__ __forceinline__ void ld4(const T* p, float (&o)[4]);
template <> __device__ __forceinline__ void ld4<float>(const float* p, float (&o)[4]) {
  const float4 v = *reinterpret_cast<const float4*>(p); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <> __device__ __forceinline__ void ld4<__nv_bfloat16>(const __nv_bfloat16* p, float (&o)[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  o[0] = __uint_as_float(u.x << 16); o[1] = __uint_as_float(u.x & 0xffff0000u);
  o[2] = __uint_as_float(u.y << 16); o[3] = __uint_as_float(u.y & 0xffff0000u);
}
template <typename T> __device__ __forceinline__ void st4(T* p, const float (&v)[4]);
template <> __device__ __forceinline__ void st4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}

template <typename T, int NC>
__global__ void __launch_bounds__(256) carafe_head_bwd_a_kernel(const T* __restrict__ enc, int64_t ldenc, const T* __restrict__ z,
                                                                 int64_t ldz, const float* __restrict__ dl, T* __restrict__ denc,
                                                                 int64_t lddenc, float* __restrict__ kws, float* __restrict__ dbias,
                                                                 int B, int H, int W) {
  __shared__ float sdb[NC];
  if (threadIdx.x < NC) sdb[threadIdx.x] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int xgroups = (W + 7) >> 3;
  const int64_t wid = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const bool wvalid = wid < (int64_t)B * H * xgroups;
  const int xg = (int)(wid % xgroups);
  const int y0 = (int)((wid / xgroups) % H);
  const int b = (int)(wid / ((int64_t)xgroups * H));
  const int x0 = xg * 8 + (lane >> 2), ay = lane & 3;
  const bool valid = wvalid && x0 < W;
  const int Ho = H * 4, Wo = W * 4, oy = y0 * 4 + ay;
  float dbl[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) dbl[c] = 0.f;
  if (valid) {
    const int64_t pix = ((int64_t)b * H + y0) * W + x0;
    float k[9][4];
#pragma unroll
    for (int t = 0; t < 9; ++t) ld4(enc + pix * ldenc + t * 16 + ay * 4, k[t]);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float mx = k[0][a];
#pragma unroll
      for (int t = 1; t < 9; ++t) mx = fmaxf(mx, k[t][a]);
      float sum = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) { k[t][a] = expf(k[t][a] - mx); sum += k[t][a]; }
      const float inv = 1.0f / sum;
#pragma unroll
      for (int t = 0; t < 9; ++t) k[t][a] *= inv;
    }
    float g[NC][4];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      ld4(dl + (((int64_t)b * NC + c) * Ho + oy) * Wo + x0 * 4, g[c]);
      dbl[c] = (g[c][0] + g[c][1]) + (g[c][2] + g[c][3]);
    }
    float dk[9][4];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
#pragma unroll
      for (int a = 0; a < 4; ++a) dk[t][a] = 0.f;
      const int yy = y0 + t / 3 - 1, xx = x0 + t % 3 - 1;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
        const T* zr = z + (((int64_t)b * H + yy) * W + xx) * ldz;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const float zc = ldf(zr + c);
#pragma unroll
          for (int a = 0; a < 4; ++a) dk[t][a] = fmaf(g[c][a], zc, dk[t][a]);
        }
      }
    }
    float dot[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int a = 0; a < 4; ++a) dot[a] = fmaf(k[t][a], dk[t][a], dot[a]);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float de[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) de[a] = k[t][a] * (dk[t][a] - dot[a]);
      st4(denc + pix * lddenc + t * 16 + ay * 4, de);
      *reinterpret_cast<float4*>(kws + (pix * 9 + t) * 16 + ay * 4) = make_float4(k[t][0], k[t][1], k[t][2], k[t][3]);
    }
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const float sdbl = warp_sum(dbl[c]);
    if (lane == 0 && sdbl != 0.f) atomicAdd(&sdb[c], sdbl);
  }
  __syncthreads();
  if (threadIdx.x < NC && sdb[threadIdx.x] != 0.f) atomicAdd(dbias + threadIdx.x, sdb[threadIdx.x]);
}

template <typename T, int NC>
__global__ void __launch_bounds__(256) carafe_head_bwd_b_kernel(const float* __restrict__ dl, const float* __restrict__ kws,
                                                                 T* __restrict__ dz, int64_t lddz, int zcols, int64_t npix, int H, int W) {
  constexpr int HC = (NC + 1) / 2;                                  // classes per lane half
  const int lane = threadIdx.x & 31, ae = lane & 15, h = lane >> 4;
  const int64_t pix = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (pix >= npix) return;
  const int x0 = (int)(pix % W), y0 = (int)((pix / W) % H);
  const int64_t b = pix / ((int64_t)W * H);
  const int Ho = H * 4, Wo = W * 4;
  float acc[HC];
#pragma unroll
  for (int j = 0; j < HC; ++j) acc[j] = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int py = y0 - (t / 3 - 1), px = x0 - (t % 3 - 1);          // the pixel whose tap t reads (y0, x0)
    if (py < 0 || py >= H || px < 0 || px >= W) continue;
    const int64_t pp = (b * H + py) * W + px;
    const float kv = kws[(pp * 9 + t) * 16 + ae];
    const float* src = dl + ((b * NC) * Ho + py * 4 + (ae >> 2)) * (int64_t)Wo + px * 4 + (ae & 3);
#pragma unroll
    for (int j = 0; j < HC; ++j) {
      const int c = h * HC + j;
      if (c < NC) acc[j] = fmaf(kv, src[(int64_t)c * Ho * Wo], acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < HC; ++j) {
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  }
  if (ae == 0) {
#pragma unroll
    for (int j = 0; j < HC; ++j) {
      const int c = h * HC + j;
      if (c < NC) stf(dz + pix * lddz + c, acc[j]);
    }
  }
  if (lane >= NC && lane < zcols) stf(dz + pix * lddz + lane, 0.f);    // padded columns
}

// ---------------------------------------------------------------------------------------------------------------
// Segmentation loss of the baseline trainer: w_ce * CrossEntropy + w_dice * DiceLoss(softmax=True)  (trainer.py:55-57,
// utils.py:9-45) on fp32 NCHW logits, two passes over the logits in total:
//   forward : per pixel softmax; sums[0] += -log p[y]; per class c: sums[1+c] += p_c [y==c] (I), sums[1+NC+c] += p_c^2 (Z),
//             sums[1+2NC+c] += [y==c] (Y)   -> loss = w_ce sums[0]/P + w_dice mean_c (1 - (2I+s)/(Z+Y+s)), s = 1e-5
//   backward: d logit_c = gout * ( w_ce/P (p_c - [y==c]) + p_c (q_c - sum_k p_k q_k) ),
//             q_c = -w_dice/NC * (2 [y==c] D_c - 2 p_c N_c) / D_c^2,  N_c = 2I_c + s, D_c = Z_c + Y_c + s.
template <int NC> __device__ __forceinline__ int load_label(const void* lab, int lbytes, int64_t i) {
  return lbytes == 8 ? (int)reinterpret_cast<const long long*>(lab)[i] : lbytes == 4 ? reinterpret_cast<const int*>(lab)[i]
                                                                                      : (int)reinterpret_cast<const uint8_t*>(lab)[i];
}

template <int NC>
__global__ void __launch_bounds__(256) seg_loss_fwd_kernel(const float* __restrict__ logits, const void* __restrict__ labels,
                                                            int lbytes, float* __restrict__ sums, int64_t B, int64_t HW) {
  float acc[1 + 3 * NC];
#pragma unroll
  for (int j = 0; j < 1 + 3 * NC; ++j) acc[j] = 0.f;
  const int64_t total = B * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / HW, px = i - b * HW;
    const float* lp = logits + b * NC * HW + px;
    float l[NC], mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < NC; ++c) { l[c] = lp[(int64_t)c * HW]; mx = fmaxf(mx, l[c]); }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) { l[c] = __expf(l[c] - mx); sum += l[c]; }
    const float inv = 1.0f / sum;
    const int y = load_label<NC>(labels, lbytes, i);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float p = l[c] * inv;
      if (c == y) { acc[0] -= __logf(fmaxf(p, 1e-38f)); acc[1 + c] += p; acc[1 + 2 * NC + c] += 1.f; }
      acc[1 + NC + c] = fmaf(p, p, acc[1 + NC + c]);
    }
  }
  __shared__ float red[8][1 + 3 * NC];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 1 + 3 * NC; ++j) {
    const float v = warp_sum(acc[j]);
    if (lane == 0) red[w][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < 1 + 3 * NC) {
    float v = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) v += red[ww][threadIdx.x];
    atomicAdd(sums + threadIdx.x, v);
  }
}

template <int NC>
__global__ void __launch_bounds__(256) seg_loss_bwd_kernel(const float* __restrict__ logits, const void* __restrict__ labels,
                                                            int lbytes, const float* __restrict__ sums, const float* __restrict__ gout,
                                                            float* __restrict__ dlogits, float w_ce, float w_dice, int64_t B, int64_t HW) {
  const int64_t total = B * HW;
  const float go = *gout;
  const float ce_scale = go * w_ce / (float)total;
  float a_y[NC], a_p[NC];                         // q_c = a_y[c] [y==c] + a_p[c] p_c
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const float N = 2.f * sums[1 + c] + 1e-5f, D = sums[1 + NC + c] + sums[1 + 2 * NC + c] + 1e-5f;
    const float k = -go * w_dice / ((float)NC * D * D);
    a_y[c] = k * 2.f * D;
    a_p[c] = -k * 2.f * N;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / HW, px = i - b * HW;
    const float* lp = logits + b * NC * HW + px;
    float p[NC], mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < NC; ++c) { p[c] = lp[(int64_t)c * HW]; mx = fmaxf(mx, p[c]); }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) { p[c] = __expf(p[c] - mx); sum += p[c]; }
    const float inv = 1.0f / sum;
    const int y = load_label<NC>(labels, lbytes, i);
    float q[NC], dot = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      p[c] *= inv;
      q[c] = fmaf(a_p[c], p[c], c == y ? a_y[c] : 0.f);
      dot = fmaf(p[c], q[c], dot);
    }
    float* dp = dlogits + b * NC * HW + px;
#pragma unroll
    for (int c = 0; c < NC; ++c) dp[(int64_t)c * HW] = fmaf(ce_scale, p[c] - (c == y ? 1.f : 0.f), p[c] * (q[c] - dot));
  }
}

unsigned grid_for(int64_t total, int per_cta) {
  return (unsigned)std::min<int64_t>(ceil_div64(total, per_cta), (int64_t)sm_count() * 32);
}

// ---------------- fused SGD(momentum, weight decay) over every parameter + bf16 shadow refresh ----------------
__global__ void __launch_bounds__(256) sgd_momentum_kernel(const cswin_sgd_chunk_t* __restrict__ chunks, const float* __restrict__ lr_p,
                                                            float momentum, float wd) {
  const cswin_sgd_chunk_t c = chunks[blockIdx.x];
  const float lr = *lr_p;
  const bool vec = ((reinterpret_cast<uintptr_t>(c.param) | reinterpret_cast<uintptr_t>(c.grad) | reinterpret_cast<uintptr_t>(c.momentum)) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(c.shadow) & 7) == 0;
  const int64_t n4 = vec ? (c.n >> 2) : 0;
  __nv_bfloat16* sh = reinterpret_cast<__nv_bfloat16*>(c.shadow);
  for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
    float4 p = reinterpret_cast<float4*>(c.param)[i];
    const float4 g = reinterpret_cast<const float4*>(c.grad)[i];
    float4 m = reinterpret_cast<float4*>(c.momentum)[i];
    m.x = fmaf(momentum, m.x, fmaf(wd, p.x, g.x)); m.y = fmaf(momentum, m.y, fmaf(wd, p.y, g.y));
    m.z = fmaf(momentum, m.z, fmaf(wd, p.z, g.z)); m.w = fmaf(momentum, m.w, fmaf(wd, p.w, g.w));
    p.x = fmaf(-lr, m.x, p.x); p.y = fmaf(-lr, m.y, p.y); p.z = fmaf(-lr, m.z, p.z); p.w = fmaf(-lr, m.w, p.w);
    reinterpret_cast<float4*>(c.momentum)[i] = m;
    reinterpret_cast<float4*>(c.param)[i] = p;
    if (sh != nullptr) {
      const __nv_bfloat162 a = __floats2bfloat162_rn(p.x, p.y), b = __floats2bfloat162_rn(p.z, p.w);
      reinterpret_cast<uint2*>(sh)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    }
  }
  for (int64_t i = n4 * 4 + threadIdx.x; i < c.n; i += blockDim.x) {
    float p = c.param[i];
    const float m = fmaf(momentum, c.momentum[i], fmaf(wd, p, c.grad[i]));
    p = fmaf(-lr, m, p);
    c.momentum[i] = m;
    c.param[i] = p;
    if (sh != nullptr) sh[i] = __float2bfloat16_rn(p);
  }
}

}  // namespace

int sgd_momentum_step(const cswin_sgd_chunk_t* chunks, int n_chunks, const float* lr, float momentum, float wd, cudaStream_t s) {
  CSWIN_REQUIRE(chunks && lr && n_chunks >= 0, CSWIN_ERR_INVALID, "sgd_momentum_step: bad arguments");
  if (n_chunks == 0) return CSWIN_OK;
  sgd_momentum_kernel<<<(unsigned)n_chunks, 256, 0, s>>>(chunks, lr, momentum, wd);
  CSWIN_LAUNCH_CHECK();
  return CSWIN_OK;
}

namespace {
}  // namespace

int act_fwd(const void* z, int64_t ldz, void* out, int64_t ldo, int64_t M, int N, int act, int dtype, cudaStream_t s) {
  CSWIN_REQUIRE(z && out, CSWIN_ERR_INVALID, "act_fwd: null pointer");
  if (M * N == 0) return CSWIN_OK;
  if (dtype == CSWIN_BF16 && act == 1 && ldz == N && ldo == N && N % 8 == 0 && (reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(out)) % 16 == 0) {
    const int64_t nvec = M * N / 8;
    CSWIN_CUDA_OK(launch_pdl(act_fwd_bf16_vec_kernel, dim3(grid_for(nvec, 256)), dim3(256), (size_t)0, s, (const uint4*)z, (uint4*)out, nvec));
    CSWIN_LAUNCH_CHECK();
    return CSWIN_OK;
  }
  const unsigned grid = grid_for(M * N, 256);
  if (dtype == CSWIN_F32) act_fwd_kernel<float><<<grid, 256, 0, s>>>((const float*)z, ldz, (float*)out, ldo, M, N, act);
  else act_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)z, ldz, (__nv_bfloat16*)out, ldo, M, N, act);
  CSWIN_LAUNCH_CHECK();
  return CSWIN_OK;
}

int act_bwd(const void* dout, int64_t ldd, const void* z, int64_t ldz, const float* sscale, int rps, void* dz, int64_t ldo,
            int64_t M, int N, int act, int dtype, cudaStream_t s) {
  CSWIN_REQUIRE(dout && dz && (act == 0 || z), CSWIN_ERR_INVALID, "act_bwd: null pointer");
  CSWIN_REQUIRE(!sscale || rps > 0, CSWIN_ERR_INVALID, "act_bwd: rows_per_sample must be > 0");
  if (M * N == 0) return CSWIN_OK;
  if (dtype == CSWIN_BF16 && ldd == N && ldo == N && (act == 0 || ldz == N) && N % 8 == 0 && (!sscale || ((int64_t)rps * N) % 8 == 0) &&
      (reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dz) | (act ? reinterpret_cast<uintptr_t>(z) : 0)) % 16 == 0) {
    const int64_t nvec = M * N / 8;
    CSWIN_CUDA_OK(launch_pdl(act_bwd_bf16_vec_kernel, dim3(grid_for(nvec, 256)), dim3(256), (size_t)0, s, (const uint4*)dout, (const uint4*)z, sscale, (int64_t)rps * N, (uint4*)dz, nvec, act));
    CSWIN_LAUNCH_CHECK();
    return CSWIN_OK;
  }
  const unsigned grid = grid_for(M * N, 256);
  if (dtype == CSWIN_F32) act_bwd_kernel<float><<<grid, 256, 0, s>>>((const float*)dout, ldd, (const float*)z, ldz, sscale, rps, (float*)dz, ldo, M, N, act);
  else act_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)dout, ldd, (const __nv_bfloat16*)z, ldz, sscale, rps, (__nv_bfloat16*)dz, ldo, M, N, act);
  CSWIN_LAUNCH_CHECK();
  return CSWIN_OK;
}

int linear_wgrad(const void* dz, int64_t ldz, const void* a, int64_t lda, float* dw, int64_t ldw, float* db, int64_t M,
                 int N, int K, int dtype, cudaStream_t s) {
  CSWIN_REQUIRE(dz && a && dw, CSWIN_ERR_INVALID, "linear_wgrad: null pointer");
  if (M == 0 || N == 0 || K == 0) return CSWIN_OK;
  const int tiles = ((N + WT - 1) / WT) * ((K + WT - 1) / WT);
  int split = (int)std::min<int64_t>((int64_t)std::max(1, 4 * sm_count() / tiles), ceil_div64(M, 256));
  if (split < 1) split = 1;
  if (split > 65535) split = 65535;
  int64_t mchunk = ceil_div64(ceil_div64(M, split), WM) * WM;
  split = (int)ceil_div64(M, mchunk);
  dim3 grid((N + WT - 1) / WT, (K + WT - 1) / WT, split);
  if (dtype == CSWIN_F32) linear_wgrad_kernel<float><<<grid, 256, 0, s>>>((const float*)dz, ldz, (const float*)a, lda, dw, ldw, db, M, N, K, mchunk);
  else linear_wgrad_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)dz, ldz, (const __nv_bfloat16*)a, lda, dw, ldw, db, M, N, K, mchunk);
  CSWIN_LAUNCH_CHECK();
  return CSWIN_OK;
}

int layernorm_bwd(const void* x, int64_t ldx, const void* dy, int64_t ldy, const void* gamma, const float* mean,
                  const float* rstd, void* dx, int64_t ldo, const void* dx_add, int64_t lda, float* dgamma, float* dbeta,
                  int64_t M, int C, int dtype, cudaStream_t s) {
  CSWIN_REQUIRE(x && dy && gamma && mean && rstd && dx && dgamma && dbeta, CSWIN_ERR_INVALID, "layernorm_bwd: null pointer");
  CSWIN_REQUIRE(C > 0 && C <= 512, CSWIN_ERR_UNSUPPORTED, "layernorm_bwd: C=%d outside (0, 512]", C);
  if (M == 0) return CSWIN_OK;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div64(M, 8), (int64_t)sm_count() * 4);
  const bool vec = dtype == CSWIN_BF16 && (ldx % 8 == 0) && (ldy % 8 == 0) && (ldo % 8 == 0) && (lda % 8 == 0) &&
                   ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) |
                     reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(dx_add)) & 15) == 0;
  if (vec && (C == 64 || C == 128 || C == 256 || C == 512)) {
    const int rpw = C == 64 ? 4 : C == 128 ? 2 : 1;
    const unsigned gv = (unsigned)std::min<int64_t>(ceil_div64(M, 8 * rpw), (int64_t)sm_count() * 2);
#define LNV(L, V) CSWIN_CUDA_OK(launch_pdl(layernorm_bwd_bf16_vec_kernel<L, V>, dim3(gv), dim3(256), (size_t)0, s, (const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)dy, ldy, (const __nv_bfloat16*)gamma, mean, rstd, (__nv_bfloat16*)dx, ldo, (const __nv_bfloat16*)dx_add, lda, dgamma, dbeta, M))
    if (C == 64) LNV(8, 1); else if (C == 128) LNV(16, 1); else if (C == 256) LNV(32, 1); else LNV(32, 2);
#undef LNV
    CSWIN_LAUNCH_CHECK();
    return CSWIN_OK;
  }
#define LNB(T, V) layernorm_bwd_kernel<T, V><<<grid, 256, 0, s>>>((const T*)x, ldx, (const T*)dy, ldy, (const T*)gamma, mean, rstd, (T*)dx, ldo, (const T*)dx_add, lda, dgamma, dbeta, M, C)
  if (dtype == CSWIN_F32) { if (C <= 128) LNB(float, 4); else LNB(float, 16); }
  else { if (C <= 128) LNB(__nv_bfloat16, 4); else LNB(__nv_bfloat16, 16); }
#undef LNB
  CSWIN_LAUNCH_CHECK();
  return CSWIN_OK;
}

int col2im_tokens(const void* dcol, int64_t ldcol, void* dx, int64_t x_bs, int64_t x_ts, int B, int H, int W, int C, int KH,
                  int KW, int stride, int pad, int dtype, cudaStream_t s) {
  CSWIN_REQUIRE(dcol && dx, CSWIN_ERR_INVALID, "col2im_tokens: null pointer");
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  const int64_t total = (int64_t)B * H * W * C;
  if (total == 0) return CSWIN_OK;
  const unsigned grid = grid_for(total, 256);
  if (dtype == CSWIN_BF16 && C % 8 == 0 && ldcol % 8 == 0 && x_bs % 8 == 0 && x_ts % 8 == 0 &&
      ((reinterpret_cast<uintptr_t>(dcol) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0) {
    const unsigned gv = grid_for(total / 8, 256);
    CSWIN_CUDA_OK(launch_pdl(col2im_tokens_bf16_vec_kernel, dim3(gv), dim3(256), (size_t)0, s, (const __nv_bfloat16*)dcol, ldcol,
                             (__nv_bfloat16*)dx, x_bs, x_ts, B, H, W, C, KH, KW, stride, pad, Ho, Wo));
    CSWIN_LAUNCH_CHECK();
    return CSWIN_OK;
  }
  if (dtype == CSWIN_F32) CSWIN_CUDA_OK(launch_pdl(col2im_tokens_kernel<float>, dim3(grid), dim3(256), (size_t)0, s, (const float*)dcol, ldcol, (float*)dx, x_bs, x_ts, B, H, W, C, KH, KW, stride, pad, Ho, Wo));
  else CSWIN_CUDA_OK(launch_pdl(col2im_tokens_kernel<__nv_bfloat16>, dim3(grid), dim3(256), (size_t)0, s, (const __nv_bfloat16*)dcol, ldcol, (__nv_bfloat16*)dx, x_bs, x_ts, B, H, W, C, KH, KW, stride, pad, Ho, Wo));
  CSWIN_LAUNCH_CHECK();
  return CSWIN_OK;
}

namespace {
template <typename T, typename TG>
int carafe_bwd_launch(const void* enc, int64_t ldenc, const void* z, int64_t ldz, const void* dy, int64_t sb, int64_t sy,
                      int64_t sx, int64_t sc, void* denc, int64_t lddenc, void* dz, int64_t lddz, float* dbias, float* kws,
                      int64_t npix, int H, int W, int C, int up, cudaStream_t s) {
  const int G = 32 / (up * up), cpl = C / G, v = C / 32;
  const unsigned ga = (unsigned)std::min<int64_t>(ceil_div64(npix, 8), (int64_t)sm_count() * 8);
  const unsigned gb = (unsigned)ceil_div64(npix, 8);
  const size_t smem = sizeof(float) * C;
#define CSWIN_CA(CPL) carafe_bwd_a_kernel<T, TG, CPL><<<ga, 256, smem, s>>>((const T*)enc, ldenc, (const T*)z, ldz, (const TG*)dy, sb, sy, sx, sc, (T*)denc, lddenc, kws, dbias, npix, H, W, C, up)
  if (cpl == 8) CSWIN_CA(8); else if (cpl == 16) CSWIN_CA(16); else if (cpl == 32) CSWIN_CA(32); else if (cpl == 64) CSWIN_CA(64);
  else { set_error("carafe_reassemble_bwd: C=%d with up=%d is not supported", C, up); return CSWIN_ERR_UNSUPPORTED; }
#undef CSWIN_CA
  CSWIN_LAUNCH_CHECK();
#define CSWIN_CB(UP, V) carafe_bwd_b_kernel<T, TG, UP, V><<<gb, 256, 0, s>>>((const TG*)dy, sb, sy, sx, sc, kws, (T*)dz, lddz, npix, H, W, C)
  if (up == 2 && v == 2) CSWIN_CB(2, 2); else if (up == 2 && v == 4) CSWIN_CB(2, 4); else if (up == 2 && v == 8) CSWIN_CB(2, 8);
  else if (up == 4 && v == 2) CSWIN_CB(4, 2); else if (up == 4 && v == 4) CSWIN_CB(4, 4); else if (up == 2 && v == 1) CSWIN_CB(2, 1);
  else if (up == 4 && v == 1) CSWIN_CB(4, 1);
  else { set_error("carafe_reassemble_bwd: C=%d with up=%d is not supported", C, up); return CSWIN_ERR_UNSUPPORTED; }
#undef CSWIN_CB
  CSWIN_LAUNCH_CHECK();
  return CSWIN_OK;
}
}  // namespace

int carafe_reassemble_bwd(const void* enc, int64_t ldenc, const void* z, int64_t ldz, const void* dy, int dy_is_f32,
                          int64_t sb, int64_t sy, int64_t sx, int64_t sc, void* denc, int64_t lddenc, void* dz, int64_t lddz,
                          float* dbias, float* kws, int B, int H, int W, int C, int up, int dtype, cudaStream_t s) {
  CSWIN_REQUIRE(enc && z && dy && denc && dz && dbias && kws, CSWIN_ERR_INVALID, "carafe_reassemble_bwd: null pointer");
  CSWIN_REQUIRE((up == 2 || up == 4) && C >= 32 && C % 32 == 0 && C <= 256, CSWIN_ERR_UNSUPPORTED,
                "carafe_reassemble_bwd: supported: up in {2,4}, C a multiple of 32 up to 256 (got up=%d C=%d)", up, C);
  const int64_t npix = (int64_t)B * H * W;
  if (npix == 0) return CSWIN_OK;
  if (dtype == CSWIN_F32) {
    CSWIN_REQUIRE(dy_is_f32, CSWIN_ERR_INVALID, "carafe_reassemble_bwd: fp32 path needs fp32 dy");
    return carafe_bwd_launch<float, float>(enc, ldenc, z, ldz, dy, sb, sy, sx, sc, denc, lddenc, dz, lddz, dbias, kws, npix, H, W, C, up, s);
  }
  if (dy_is_f32)
    return carafe_bwd_launch<__nv_bfloat16, float>(enc, ldenc, z, ldz, dy, sb, sy, sx, sc, denc, lddenc, dz, lddz, dbias, kws, npix, H, W, C, up, s);
  return carafe_bwd_launch<__nv_bfloat16, __nv_bfloat16>(enc, ldenc, z, ldz, dy, sb, sy, sx, sc, denc, lddenc, dz, lddz, dbias, kws, npix, H, W, C, up, s);
}

int carafe_head_bwd(const void* enc, int64_t ldenc, const void* z, int64_t ldz, const float* dlogits, void* denc, int64_t lddenc,
                    void* dz, int64_t lddz, int zcols, float* dbias, float* kws, int B, int H, int W, int C, int up, int dtype,
                    cudaStream_t s) {
  CSWIN_REQUIRE(enc && z && dlogits && denc && dz && dbias && kws, CSWIN_ERR_INVALID, "carafe_head_bwd: null pointer");
  CSWIN_REQUIRE(up == 4 && (C == 2 || C == 3 || C == 4 || C == 9), CSWIN_ERR_UNSUPPORTED,
                "carafe_head_bwd: supported: up = 4 and 2, 3, 4 or 9 classes (got up=%d, %d classes)", up, C);
  CSWIN_REQUIRE(ldenc >= 144 && lddenc >= 144 && ldz >= C && lddz >= zcols && zcols >= C && zcols <= 32, CSWIN_ERR_INVALID,
                "carafe_head_bwd: bad leading dimensions");
  const int es = dtype == CSWIN_F32 ? 4 : 2;
  CSWIN_REQUIRE((ldenc * es) % (4 * es) == 0 && (lddenc * es) % (4 * es) == 0 && reinterpret_cast<uintptr_t>(enc) % (4 * es) == 0 &&
                reinterpret_cast<uintptr_t>(denc) % (4 * es) == 0 && reinterpret_cast<uintptr_t>(dlogits) % 16 == 0 &&
                reinterpret_cast<uintptr_t>(kws) % 16 == 0, CSWIN_ERR_UNSUPPORTED, "carafe_head_bwd: operands must be vector-aligned");
  const int64_t npix = (int64_t)B * H * W;
  if (npix == 0) return CSWIN_OK;
  const unsigned ga = (unsigned)ceil_div64((int64_t)B * H * ((W + 7) / 8), 8), gb = (unsigned)ceil_div64(npix, 8);
#define HB(T, NC_) do { \
    carafe_head_bwd_a_kernel<T, NC_><<<ga, 256, 0, s>>>((const T*)enc, ldenc, (const T*)z, ldz, dlogits, (T*)denc, lddenc, kws, dbias, B, H, W); \
    carafe_head_bwd_b_kernel<T, NC_><<<gb, 256, 0, s>>>(dlogits, kws, (T*)dz, lddz, zcols, npix, H, W); } while (0)
#define HBT(NC_) do { if (dtype == CSWIN_F32) HB(float, NC_); else HB(__nv_bfloat16, NC_); } while (0)
  if (C == 9) HBT(9); else if (C == 4) HBT(4); else if (C == 3) HBT(3); else HBT(2);
#undef HBT
#undef HB
  CSWIN_LAUNCH_CHECK();
  return CSWIN_OK;
}

int seg_loss_fwd(const float* logits, const void* labels, int label_bytes, float* sums, int64_t B, int C, int64_t HW, cudaStream_t s) {
  CSWIN_REQUIRE(logits && labels && sums && B >= 0 && HW > 0, CSWIN_ERR_INVALID, "seg_loss_fwd: bad arguments");
  CSWIN_REQUIRE(label_bytes == 1 || label_bytes == 4 || label_bytes == 8, CSWIN_ERR_INVALID, "seg_loss: labels must be uint8, int32 or int64");
  CSWIN_REQUIRE(C == 2 || C == 3 || C == 4 || C == 9, CSWIN_ERR_UNSUPPORTED, "seg_loss: supported class counts: 2, 3, 4, 9 (got %d)", C);
  if (B == 0) return CSWIN_OK;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div64(B * HW, 256), (int64_t)sm_count() * 8);
#define SL(NC_) seg_loss_fwd_kernel<NC_><<<grid, 256, 0, s>>>(logits, labels, label_bytes, sums, B, HW)
  if (C == 9) SL(9); else if (C == 4) SL(4); else if (C == 3) SL(3); else SL(2);
#undef SL
  CSWIN_LAUNCH_CHECK();
  return CSWIN_OK;
}

int seg_loss_bwd(const float* logits, const void* labels, int label_bytes, const float* sums, const float* gout, float* dlogits,
                 float w_ce, float w_dice, int64_t B, int C, int64_t HW, cudaStream_t s) {
  CSWIN_REQUIRE(logits && labels && sums && gout && dlogits && B >= 0 && HW > 0, CSWIN_ERR_INVALID, "seg_loss_bwd: bad arguments");
  CSWIN_REQUIRE(label_bytes == 1 || label_bytes == 4 || label_bytes == 8, CSWIN_ERR_INVALID, "seg_loss: labels must be uint8, int32 or int64");
  CSWIN_REQUIRE(C == 2 || C == 3 || C == 4 || C == 9, CSWIN_ERR_UNSUPPORTED, "seg_loss: supported class counts: 2, 3, 4, 9 (got %d)", C);
  if (B == 0) return CSWIN_OK;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div64(B * HW, 256), (int64_t)sm_count() * 16);
#define SL(NC_) seg_loss_bwd_kernel<NC_><<<grid, 256, 0, s>>>(logits, labels, label_bytes, sums, gout, dlogits, w_ce, w_dice, B, HW)
  if (C == 9) SL(9); else if (C == 4) SL(4); else if (C == 3) SL(3); else SL(2);
#undef SL
  CSWIN_LAUNCH_CHECK();
  return CSWIN_OK;
}

}  // namespace cswin
