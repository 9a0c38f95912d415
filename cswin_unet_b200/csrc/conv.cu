// conv.cu — gathers that turn the convolutions of the hot path into Linear calls, and the CARAFE reassembly.
//
//  im2col_tokens : Merge_Block.conv (3x3 s2 p1, networks/cswin_unet.py:216) and CARAFE.encoder (3x3 s1 p1, :241)
//                  straight from the token-major (B, L, C) activation — the reference first materialises an NCHW
//                  copy (:214, :235); here channels stay innermost so every access is a contiguous C-vector.
//  im2col_nchw   : stem conv (7x7 s4 p2, :339) from the NCHW network input.
//  carafe_reassemble : pixel_shuffle + softmax over the 9 taps + 3x3 neighbourhood re-assembly + pixel_shuffle
//                  (:242-263 / :292-313), run on the output of the 1x1 `out` conv taken at LOW resolution (see
//                  cswin_b200.h).  All three are HBM-bound streaming kernels.
#include <cstdlib>

#include "common.cuh"

namespace cswin {
namespace {

template <typename T>
__global__ void __launch_bounds__(256) im2col_tokens_kernel(const T* __restrict__ x, int64_t x_bs, int64_t x_ts,
                                                             T* __restrict__ col, int64_t ldcol, int B, int H, int W,
                                                             int C, int KH, int KW, int stride, int pad, int Ho, int Wo, int vec) {
  // one thread per (output pixel, tap, 8-channel chunk); C % 8 == 0 is guaranteed by the launcher
  pdl_trigger();
  pdl_wait();
  const int cv = C >> 3;
  const int64_t total = (int64_t)B * Ho * Wo * KH * KW * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    int64_t r = i / cv;
    const int tap = (int)(r % (KH * KW)); r /= (KH * KW);
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const int ky = tap / KW, kx = tap - ky * KW;
    const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
    T* dst = col + ((int64_t)(b * Ho + oy) * Wo + ox) * ldcol + (int64_t)tap * C + c8 * 8;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      const T* src = x + (int64_t)b * x_bs + ((int64_t)iy * W + ix) * x_ts + c8 * 8;
      if (vec) {
#pragma unroll
        for (int j = 0; j < (int)(8 * sizeof(T) / 16); ++j) reinterpret_cast<uint4*>(dst)[j] = reinterpret_cast<const uint4*>(src)[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = src[j];
      }
    } else {
      if (vec) {
#pragma unroll
        for (int j = 0; j < (int)(8 * sizeof(T) / 16); ++j) reinterpret_cast<uint4*>(dst)[j] = make_uint4(0, 0, 0, 0);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) stf(dst + j, 0.f);
      }
    }
  }
}

// One CTA per (batch, output row, 32-pixel segment): the KH input rows needed by the segment are staged in shared
// memory with coalesced loads (fp32 or bf16 image), then each of the 32 col rows (ldcol elements, zero padded) is
// written contiguously.  Column order (c*KH + ky)*KW + kx == flattening of conv.weight (N, C, KH, KW).
constexpr int kI2cSeg = 32;
// CK/CS > 0: compile-time kernel size / stride (the stem is 7x7 stride 4), so the index divisions become mul-shifts.
template <typename TI, typename T, int CK, int CS>
__global__ void __launch_bounds__(256) im2col_nchw_kernel(const TI* __restrict__ x, T* __restrict__ col, int64_t ldcol_,
                                                           int B, int C, int H, int W, int KH_, int KW_, int stride_,
                                                           int pad, int Ho, int Wo, int vec8) {
  const int KH = CK > 0 ? CK : KH_, KW = CK > 0 ? CK : KW_, stride = CS > 0 ? CS : stride_;
  const int ldcol = (int)ldcol_;
  extern __shared__ float patch[];                     // [C][KH][span]
  pdl_trigger();
  pdl_wait();
  const int segs = (Wo + kI2cSeg - 1) / kI2cSeg;
  const int seg = blockIdx.x % segs;
  const int oy = (blockIdx.x / segs) % Ho;
  const int b = blockIdx.x / (segs * Ho);
  const int ox0 = seg * kI2cSeg;
  const int span = (kI2cSeg - 1) * stride + KW;
  const int ix0 = ox0 * stride - pad, iy0 = oy * stride - pad;
  for (int i = threadIdx.x; i < C * KH * span; i += blockDim.x) {
    const int dx = i % span;
    const int ky = (i / span) % KH;
    const int c = i / (span * KH);
    const int iy = iy0 + ky, ix = ix0 + dx;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = ldf(x + (((int64_t)b * C + c) * H + iy) * W + ix);
    patch[i] = v;
  }
  __syncthreads();
  const int K = C * KH * KW;
  const int npix = min(kI2cSeg, Wo - ox0);
  T* crow = col + ((int64_t)(b * Ho + oy) * Wo + ox0) * ldcol;
  if (sizeof(T) == 2 && vec8) {
    // bf16, rows of ldcol = 8 n elements, 16-byte aligned: a per-column offset table (built once per CTA) turns every output
    // element into two shared-memory loads, and each thread emits whole 16-byte stores
    uint16_t* offs = reinterpret_cast<uint16_t*>(patch + ((C * KH * span + 3) & ~3));
    for (int k = threadIdx.x; k < ldcol; k += blockDim.x) {
      const int kx = k % KW, ky = (k / KW) % KH, c = k / (KW * KH);
      offs[k] = (uint16_t)(k < K ? (c * KH + ky) * span + kx : 0);
    }
    __syncthreads();
    const int nvec = ldcol >> 3;
    for (int i = threadIdx.x; i < npix * nvec; i += blockDim.x) {
      const int p = i / nvec, v = i - p * nvec;
      const uint4 o = reinterpret_cast<const uint4*>(offs)[v];
      const float* pp = patch + p * stride;
      const int k0 = v * 8;
      const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float lo = (k0 + 2 * e < K) ? pp[ow[e] & 0xffffu] : 0.f;
        const float hi = (k0 + 2 * e + 1 < K) ? pp[ow[e] >> 16] : 0.f;
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(lo, hi);
        w[e] = *reinterpret_cast<const uint32_t*>(&h2);
      }
      *reinterpret_cast<uint4*>(crow + (int64_t)p * ldcol + k0) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    return;
  }
  for (int i = threadIdx.x; i < npix * ldcol; i += blockDim.x) {
    const int p = i / ldcol;
    const int k = i - p * ldcol;
    float v = 0.f;
    if (k < K) {
      const int kx = k % KW;
      const int ky = (k / KW) % KH;
      const int c = k / (KW * KH);
      v = patch[(c * KH + ky) * span + p * stride + kx];
    }
    stf(crow + i, v);
  }
}

// One CTA per low-resolution pixel: its 9 neighbour rows of z and its 9 s^2 encoder logits are staged once in
// shared memory and feed the s^2 output pixels it owns.
template <typename T, typename TO>
__global__ void __launch_bounds__(128) carafe_reassemble_kernel(const T* __restrict__ enc, int64_t ldenc,
                                                                 const T* __restrict__ z, int64_t ldz,
                                                                 const T* __restrict__ bias, TO* __restrict__ y,
                                                                 int64_t ldy, int nchw_out, int B, int H, int W, int C,
                                                                 int up) {
  extern __shared__ float sm[];
  const int s2 = up * up;
  float* Zs = sm;                 // [9][C]
  float* Kp = Zs + 9 * C;         // [s2][9] softmaxed taps
  const int pix = blockIdx.x;     // (b, y, x)
  const int x0 = pix % W;
  const int y0 = (pix / W) % H;
  const int b = pix / (W * H);
  const int tid = threadIdx.x;

  for (int i = tid; i < 9 * C; i += blockDim.x) {
    const int t = i / C, c = i - t * C;
    const int yy = y0 + t / 3 - 1, xx = x0 + t % 3 - 1;
    float v = 0.f;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = ldf(z + ((int64_t)(b * H + yy) * W + xx) * ldz + c);
    Zs[i] = v;
  }
  if (tid < s2) {
    const T* e = enc + (int64_t)pix * ldenc + tid;        // channel t*s2 + (a*up+e)
    float l[9];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < 9; ++t) { l[t] = ldf(e + t * s2); mx = fmaxf(mx, l[t]); }
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) { l[t] = expf(l[t] - mx); sum += l[t]; }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int t = 0; t < 9; ++t) Kp[tid * 9 + t] = l[t] * inv;
  }
  __syncthreads();

  const int Ho = H * up, Wo = W * up;
  for (int i = tid; i < s2 * C; i += blockDim.x) {
    const int ae = i / C, c = i - ae * C;
    const int a = ae / up, e = ae - a * up;
    float acc = ldf(bias + c);
#pragma unroll
    for (int t = 0; t < 9; ++t) acc = fmaf(Kp[ae * 9 + t], Zs[t * C + c], acc);
    const int oy = y0 * up + a, ox = x0 * up + e;
    if (nchw_out) stf(y + (((int64_t)b * C + c) * Ho + oy) * Wo + ox, acc);
    else          stf(y + ((int64_t)(b * Ho + oy) * Wo + ox) * ldy + c, acc);
  }
}

// bf16, token-major output: one WARP per low-resolution pixel, lane owns V = C/32 consecutive channels (V in {2,4,8}):
// the 9 neighbour rows of z are read once as 4..16-byte vectors into registers, the softmaxed taps of the s^2 sub-pixels
// are computed by lanes < s^2 and broadcast with shuffles, and every output row (C channels) is one coalesced store.
template <int V>
__global__ void __launch_bounds__(256) carafe_reassemble_warp_kernel(const __nv_bfloat16* __restrict__ enc, int64_t ldenc,
                                                                      const __nv_bfloat16* __restrict__ z, int64_t ldz,
                                                                      const __nv_bfloat16* __restrict__ bias,
                                                                      __nv_bfloat16* __restrict__ y, int64_t ldy, int64_t npix,
                                                                      int H, int W, int up) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t pix = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pix >= npix) return;
  const int s2 = up * up;
  const int p32 = (int)pix;                             // (the launcher keeps npix < 2^31: 32-bit divisions, a 64-bit one is ~100 instructions)
  const int prow = p32 / W;
  const int x0 = p32 - prow * W;
  const int64_t b = prow / H;
  const int y0 = prow - (int)b * H;
  float zr[9][V];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int yy = y0 + t / 3 - 1, xx = x0 + t % 3 - 1;
    const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
    const __nv_bfloat16* src = z + ((b * H + yy) * W + xx) * ldz + lane * V;
#pragma unroll
    for (int e = 0; e < V; e += 2) {
      uint32_t u = 0;
      if (in) u = *reinterpret_cast<const uint32_t*>(src + e);
      zr[t][e] = __uint_as_float(u << 16);
      zr[t][e + 1] = __uint_as_float(u & 0xffff0000u);
    }
  }
  float kt[9];
  {
    const int ae = lane < s2 ? lane : 0;
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < 9; ++t) { kt[t] = __bfloat162float(enc[pix * ldenc + t * s2 + ae]); mx = fmaxf(mx, kt[t]); }
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) { kt[t] = expf(kt[t] - mx); sum += kt[t]; }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int t = 0; t < 9; ++t) kt[t] *= inv;
  }
  float bv[V];
#pragma unroll
  for (int e = 0; e < V; ++e) bv[e] = __bfloat162float(bias[lane * V + e]);
  const int Wo = W * up;
  for (int ae = 0; ae < s2; ++ae) {
    float acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = bv[e];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float k = __shfl_sync(0xffffffffu, kt[t], ae);
#pragma unroll
      for (int e = 0; e < V; ++e) acc[e] = fmaf(k, zr[t][e], acc[e]);
    }
    const int oy = y0 * up + ae / up, ox = x0 * up + ae % up;
    __nv_bfloat16* dst = y + ((b * H * up + oy) * Wo + ox) * ldy + lane * V;
#pragma unroll
    for (int e = 0; e < V; e += 2) {
      const __nv_bfloat162 pk = __floats2bfloat162_rn(acc[e], acc[e + 1]);
      *reinterpret_cast<__nv_bfloat162*>(dst + e) = pk;
    }
  }
}

// Same re-assembly with LPP = C / 8 lanes per low-resolution pixel (8 channels = one 16-byte load per lane and tap) and 32 / LPP pixels
// per warp: for C = 64 the warp-per-pixel kernel above moves 4 bytes per lane and spends most of its instructions on per-pixel index
// math and the 9-tap softmax; here those are shared by 4 (C = 64) or 2 (C = 128) pixels per warp.  Needs up^2 <= LPP (the lanes
// gl < up^2 of a group own one sub-pixel's softmax each).  Same arithmetic, same order of accumulation as the kernel above.
template <int LPP>
__global__ void __launch_bounds__(256) carafe_reassemble_grp_kernel(const __nv_bfloat16* __restrict__ enc, int64_t ldenc,
                                                                     const __nv_bfloat16* __restrict__ z, int64_t ldz,
                                                                     const __nv_bfloat16* __restrict__ bias,
                                                                     __nv_bfloat16* __restrict__ y, int64_t ldy, int64_t npix,
                                                                     int H, int W, int up) {
  pdl_trigger();
  pdl_wait();
  constexpr int PPW = 32 / LPP;
  const int lane = threadIdx.x & 31, grp = lane / LPP, gl = lane - grp * LPP;
  const int64_t pix_raw = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * PPW + grp;
  const bool live = pix_raw < npix;
  const int p32 = (int)(live ? pix_raw : npix - 1);     // (npix < 2^31, checked by the launcher); idle groups shadow the last pixel
  const int s2 = up * up;
  const int prow = p32 / W;
  const int x0 = p32 - prow * W;
  const int b = prow / H;
  const int y0 = prow - b * H;
  float zr[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int yy = y0 + t / 3 - 1, xx = x0 + t % 3 - 1;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (yy >= 0 && yy < H && xx >= 0 && xx < W)
      u = *reinterpret_cast<const uint4*>(z + (((int64_t)b * H + yy) * W + xx) * ldz + gl * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) { zr[t][2 * e] = __uint_as_float(w[e] << 16); zr[t][2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u); }
  }
  float kt[9];
  {
    const int ae = gl < s2 ? gl : 0;
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < 9; ++t) { kt[t] = __bfloat162float(enc[(int64_t)p32 * ldenc + t * s2 + ae]); mx = fmaxf(mx, kt[t]); }
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) { kt[t] = expf(kt[t] - mx); sum += kt[t]; }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int t = 0; t < 9; ++t) kt[t] *= inv;
  }
  float bv[8];
  {
    const uint4 u = *reinterpret_cast<const uint4*>(bias + gl * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) { bv[2 * e] = __uint_as_float(w[e] << 16); bv[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u); }
  }
  const int Wo = W * up;
  for (int ae = 0; ae < s2; ++ae) {
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = bv[e];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float k = __shfl_sync(0xffffffffu, kt[t], grp * LPP + ae);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(k, zr[t][e], acc[e]);
    }
    if (live) {
      const int oy = y0 * up + ae / up, ox = x0 * up + ae % up;
      __nv_bfloat16* dst = y + (((int64_t)b * H * up + oy) * Wo + ox) * ldy + gl * 8;
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 pk = __floats2bfloat162_rn(acc[2 * e], acc[2 * e + 1]);
        o[e] = *reinterpret_cast<const uint32_t*>(&pk);
      }
      *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// Segmentation head: CARAFE4 re-assembly of the folded (out o output) 1x1 map, one thread per OUTPUT pixel computing all
// C <= 16 classes: NCHW logits are written coalesced along x, and the arg-max label map (what test_single_volume keeps,
// utils.py:73-75 — softmax is monotone so argmax(softmax(l)) == argmax(l)) can be emitted directly as uint8.
template <typename T, typename TO>
__global__ void __launch_bounds__(256) carafe_head_kernel(const T* __restrict__ enc, int64_t ldenc, const T* __restrict__ z,
                                                           int64_t ldz, const T* __restrict__ bias, TO* __restrict__ logits,
                                                           uint8_t* __restrict__ labels, int B, int H, int W, int C, int up,
                                                           int zvec) {
  pdl_trigger();
  pdl_wait();
  const int Ho = H * up, Wo = W * up, s2 = up * up;
  const int64_t total = (int64_t)B * Ho * Wo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo);
    const int oy = (int)((i / Wo) % Ho);
    const int b = (int)(i / ((int64_t)Wo * Ho));
    const int x0 = ox / up, y0 = oy / up;
    const int ae = (oy - y0 * up) * up + (ox - x0 * up);
    const int64_t pix = ((int64_t)b * H + y0) * W + x0;
    float k[9];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < 9; ++t) { k[t] = ldf(enc + pix * ldenc + t * s2 + ae); mx = fmaxf(mx, k[t]); }
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) { k[t] = expf(k[t] - mx); sum += k[t]; }
    const float inv = 1.0f / sum;
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = (c < C) ? ldf(bias + c) : -INFINITY;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = y0 + t / 3 - 1, xx = x0 + t % 3 - 1;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
        const T* zr = z + (((int64_t)b * H + yy) * W + xx) * ldz;
        const float kt = k[t] * inv;
        if (sizeof(T) == 2 && zvec) {                       // rows padded to 16 bf16 (32 B): two 16-byte loads
          const uint4 u0 = reinterpret_cast<const uint4*>(zr)[0], u1 = reinterpret_cast<const uint4*>(zr)[1];
          const uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
          for (int c = 0; c < 16; ++c)
            if (c < C) acc[c] = fmaf(kt, __uint_as_float((c & 1) ? (w[c >> 1] & 0xffff0000u) : (w[c >> 1] << 16)), acc[c]);
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c) if (c < C) acc[c] = fmaf(kt, ldf(zr + c), acc[c]);
        }
      }
    }
    if (logits != nullptr) {
#pragma unroll
      for (int c = 0; c < 16; ++c) if (c < C) stf(logits + (((int64_t)b * C + c) * Ho + oy) * Wo + ox, acc[c]);
    }
    if (labels != nullptr) {
      int best = 0; float bv = acc[0];
#pragma unroll
      for (int c = 1; c < 16; ++c) if (c < C && acc[c] > bv) { bv = acc[c]; best = c; }     // first maximum, like torch.argmax
      labels[i] = (uint8_t)best;
    }
  }
}

// Segmentation head, up = 4, bf16, NC classes, z rows padded to 16: one warp per (batch, low-res row, 8 low-res pixels); lane =
// (pixel xl = lane / 4, output sub-row ay = lane % 4) owns the 4 output pixels (4 x0 .. 4 x0 + 3) of output row 4 y0 + ay.  The 4
// lanes of a pixel read one contiguous 32-byte sector of encoder logits per tap and share (broadcast) the 9 neighbour z rows;
// labels leave as one 32-bit store per lane (8 lanes = 32 contiguous bytes of a label row), logits as 8 / 16-byte stores.
template <int NC, typename TO>
__global__ void __launch_bounds__(256) carafe_head_up4_kernel(const __nv_bfloat16* __restrict__ enc, int64_t ldenc,
                                                               const __nv_bfloat16* __restrict__ z, int64_t ldz,
                                                               const __nv_bfloat16* __restrict__ bias, TO* __restrict__ logits,
                                                               uint8_t* __restrict__ labels, int B, int H, int W) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int xgroups = (W + 7) >> 3;
  const int64_t wid = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wid >= (int64_t)B * H * xgroups) return;
  const int w32 = (int)wid;                             // (< 2^31 warps: 32-bit divisions)
  const int wrow = w32 / xgroups;
  const int xg = w32 - wrow * xgroups;
  const int b = wrow / H;
  const int y0 = wrow - b * H;
  const int x0 = xg * 8 + (lane >> 2), ay = lane & 3;
  if (x0 >= W) return;
  const int64_t pix = ((int64_t)b * H + y0) * W + x0;
  float k[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const uint2 u = *reinterpret_cast<const uint2*>(enc + pix * ldenc + t * 16 + ay * 4);
    k[t][0] = __uint_as_float(u.x << 16); k[t][1] = __uint_as_float(u.x & 0xffff0000u);
    k[t][2] = __uint_as_float(u.y << 16); k[t][3] = __uint_as_float(u.y & 0xffff0000u);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    float mx = k[0][a];
#pragma unroll
    for (int t = 1; t < 9; ++t) mx = fmaxf(mx, k[t][a]);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) { k[t][a] = __expf(k[t][a] - mx); sum += k[t][a]; }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int t = 0; t < 9; ++t) k[t][a] *= inv;
  }
  float acc[4][NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const float bv = __bfloat162float(bias[c]);
#pragma unroll
    for (int a = 0; a < 4; ++a) acc[a][c] = bv;
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int yy = y0 + t / 3 - 1, xx = x0 + t % 3 - 1;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      const uint4* zr = reinterpret_cast<const uint4*>(z + (((int64_t)b * H + yy) * W + xx) * ldz);
      const uint4 u0 = zr[0], u1 = zr[1];
      const uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const float zc = __uint_as_float((c & 1) ? (w[c >> 1] & 0xffff0000u) : (w[c >> 1] << 16));
#pragma unroll
        for (int a = 0; a < 4; ++a) acc[a][c] = fmaf(k[t][a], zc, acc[a][c]);
      }
    }
  }
  const int Ho = H * 4, Wo = W * 4, oy = y0 * 4 + ay;
  if (logits != nullptr) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      TO* dst = logits + (((int64_t)b * NC + c) * Ho + oy) * Wo + x0 * 4;
      if (sizeof(TO) == 4) {
        *reinterpret_cast<float4*>(dst) = make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
      } else {
        const __nv_bfloat162 p0 = __floats2bfloat162_rn(acc[0][c], acc[1][c]), p1 = __floats2bfloat162_rn(acc[2][c], acc[3][c]);
        *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
      }
    }
  }
  if (labels != nullptr) {
    uint32_t packed = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      int best = 0; float bv = acc[a][0];
#pragma unroll
      for (int c = 1; c < NC; ++c) if (acc[a][c] > bv) { bv = acc[a][c]; best = c; }          // first maximum, like torch.argmax
      packed |= (uint32_t)best << (8 * a);
    }
    *reinterpret_cast<uint32_t*>(labels + ((int64_t)b * Ho + oy) * Wo + x0 * 4) = packed;
  }
}

}  // namespace

int carafe_head_fwd(const void* enc, int64_t ldenc, const void* z, int64_t ldz, const void* bias, void* logits,
                    int logits_is_f32, uint8_t* labels, int B, int H, int W, int C, int up, int dtype, cudaStream_t s) {
  CSWIN_REQUIRE(enc && z && bias && (logits || labels), CSWIN_ERR_INVALID, "carafe_head: null pointer");
  CSWIN_REQUIRE(C >= 1 && C <= 16, CSWIN_ERR_UNSUPPORTED, "carafe_head: %d classes (supported: 1..16)", C);
  CSWIN_REQUIRE(up >= 1 && ldenc >= 9 * up * up && ldz >= C, CSWIN_ERR_INVALID, "carafe_head: bad up / leading dimension");
  const int64_t total = (int64_t)B * H * up * W * up;
  if (total == 0) return CSWIN_OK;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div64(total, 256), (int64_t)sm_count() * 32);
  const int zvec = dtype == CSWIN_BF16 && ldz >= 16 && (ldz * 2) % 16 == 0 && reinterpret_cast<uintptr_t>(z) % 16 == 0;
  if (zvec && up == 4 && C == 9 && total < 0x7fffffff && (ldenc * 2) % 8 == 0 && reinterpret_cast<uintptr_t>(enc) % 8 == 0 &&
      reinterpret_cast<uintptr_t>(logits) % 16 == 0 && reinterpret_cast<uintptr_t>(labels) % 4 == 0) {
    const unsigned g4 = (unsigned)ceil_div64((int64_t)B * H * ((W + 7) / 8), 8);
    const __nv_bfloat16 *e_ = (const __nv_bfloat16*)enc, *z_ = (const __nv_bfloat16*)z, *b_ = (const __nv_bfloat16*)bias;
    if (logits_is_f32) CSWIN_CUDA_OK(launch_pdl(carafe_head_up4_kernel<9, float>, dim3(g4), dim3(256), (size_t)0, s, e_, ldenc, z_, ldz, b_, (float*)logits, labels, B, H, W));
    else CSWIN_CUDA_OK(launch_pdl(carafe_head_up4_kernel<9, __nv_bfloat16>, dim3(g4), dim3(256), (size_t)0, s, e_, ldenc, z_, ldz, b_, (__nv_bfloat16*)logits, labels, B, H, W));
    CSWIN_LAUNCH_CHECK();
    return CSWIN_OK;
  }
  if (dtype == CSWIN_F32) {
    CSWIN_REQUIRE(!logits || logits_is_f32, CSWIN_ERR_INVALID, "carafe_head: fp32 path writes fp32 logits");
    CSWIN_CUDA_OK(launch_pdl(carafe_head_kernel<float, float>, dim3(grid), dim3(256), (size_t)(0), s, (const float*)enc, ldenc, (const float*)z, ldz, (const float*)bias, (float*)logits, labels, B, H, W, C, up, zvec));
  } else if (logits_is_f32) {
    CSWIN_CUDA_OK(launch_pdl(carafe_head_kernel<__nv_bfloat16, float>, dim3(grid), dim3(256), (size_t)(0), s, (const __nv_bfloat16*)enc, ldenc, (const __nv_bfloat16*)z, ldz, (const __nv_bfloat16*)bias, (float*)logits, labels, B, H, W, C, up, zvec));
  } else {
    CSWIN_CUDA_OK(launch_pdl(carafe_head_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(grid), dim3(256), (size_t)(0), s, (const __nv_bfloat16*)enc, ldenc, (const __nv_bfloat16*)z, ldz, (const __nv_bfloat16*)bias, (__nv_bfloat16*)logits, labels, B, H, W, C, up, zvec));
  }
  CSWIN_LAUNCH_CHECK();
  return CSWIN_OK;
}

int im2col_tokens(const void* x, int64_t x_bs, int64_t x_ts, void* col, int64_t ldcol, int B, int H, int W, int C, int KH,
                  int KW, int stride, int pad, int dtype, cudaStream_t s) {
  CSWIN_REQUIRE(x && col, CSWIN_ERR_INVALID, "im2col_tokens: null pointer");
  CSWIN_REQUIRE(C % 8 == 0, CSWIN_ERR_UNSUPPORTED, "im2col_tokens: C=%d must be a multiple of 8", C);
  CSWIN_REQUIRE(ldcol >= (int64_t)KH * KW * C, CSWIN_ERR_INVALID, "im2col_tokens: ldcol too small");
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  const int64_t total = (int64_t)B * Ho * Wo * KH * KW * (C / 8);
  if (total == 0) return CSWIN_OK;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div64(total, 256), (int64_t)sm_count() * 16);
  const int es = dtype == CSWIN_F32 ? 4 : 2;
  const int vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(col)) % 16 == 0) && (x_bs * es) % 16 == 0 &&
                  (x_ts * es) % 16 == 0 && (ldcol * es) % 16 == 0;
  if (dtype == CSWIN_F32)
    CSWIN_CUDA_OK(launch_pdl(im2col_tokens_kernel<float>, dim3(grid), dim3(256), (size_t)(0), s, (const float*)x, x_bs, x_ts, (float*)col, ldcol, B, H, W, C, KH, KW, stride, pad, Ho, Wo, vec));
  else
    CSWIN_CUDA_OK(launch_pdl(im2col_tokens_kernel<__nv_bfloat16>, dim3(grid), dim3(256), (size_t)(0), s, (const __nv_bfloat16*)x, x_bs, x_ts, (__nv_bfloat16*)col, ldcol, B, H, W, C, KH, KW, stride, pad, Ho, Wo, vec));
  CSWIN_LAUNCH_CHECK();
  return CSWIN_OK;
}

int im2col_nchw(const void* x, int x_is_f32, void* col, int64_t ldcol, int B, int C, int H, int W, int KH, int KW,
                int stride, int pad, int dtype, cudaStream_t s) {
  CSWIN_REQUIRE(x && col, CSWIN_ERR_INVALID, "im2col_nchw: null pointer");
  CSWIN_REQUIRE(ldcol >= (int64_t)C * KH * KW && ldcol < (1 << 20), CSWIN_ERR_INVALID, "im2col_nchw: bad ldcol");
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  if ((int64_t)B * Ho * Wo == 0) return CSWIN_OK;
  const int span = (kI2cSeg - 1) * stride + KW;
  const int vec8 = dtype == CSWIN_BF16 && ldcol % 8 == 0 && reinterpret_cast<uintptr_t>(col) % 16 == 0 && (size_t)C * KH * span < 65536;
  const size_t smem = sizeof(float) * (((size_t)C * KH * span + 3) & ~(size_t)3) + (vec8 ? 2 * (size_t)ldcol : 0);
  CSWIN_REQUIRE(smem <= 48 * 1024, CSWIN_ERR_UNSUPPORTED, "im2col_nchw: C*KH*span=%zu floats exceed 48 KB of shared memory", smem / 4);
  const int64_t ctas = (int64_t)B * Ho * ((Wo + kI2cSeg - 1) / kI2cSeg);
  CSWIN_REQUIRE(ctas < (1ll << 31), CSWIN_ERR_UNSUPPORTED, "im2col_nchw: grid too large");
  const unsigned grid = (unsigned)ctas;
  if (dtype == CSWIN_F32) {
    CSWIN_REQUIRE(x_is_f32, CSWIN_ERR_INVALID, "im2col_nchw: fp32 path needs an fp32 image");
    if (KH == 7 && KW == 7 && stride == 4) CSWIN_CUDA_OK(launch_pdl(im2col_nchw_kernel<float, float, 7, 4>, dim3(grid), dim3(256), (size_t)(smem), s, (const float*)x, (float*)col, ldcol, B, C, H, W, KH, KW, stride, pad, Ho, Wo, vec8));
    else CSWIN_CUDA_OK(launch_pdl(im2col_nchw_kernel<float, float, 0, 0>, dim3(grid), dim3(256), (size_t)(smem), s, (const float*)x, (float*)col, ldcol, B, C, H, W, KH, KW, stride, pad, Ho, Wo, vec8));
  } else if (x_is_f32) {
    if (KH == 7 && KW == 7 && stride == 4) CSWIN_CUDA_OK(launch_pdl(im2col_nchw_kernel<float, __nv_bfloat16, 7, 4>, dim3(grid), dim3(256), (size_t)(smem), s, (const float*)x, (__nv_bfloat16*)col, ldcol, B, C, H, W, KH, KW, stride, pad, Ho, Wo, vec8));
    else CSWIN_CUDA_OK(launch_pdl(im2col_nchw_kernel<float, __nv_bfloat16, 0, 0>, dim3(grid), dim3(256), (size_t)(smem), s, (const float*)x, (__nv_bfloat16*)col, ldcol, B, C, H, W, KH, KW, stride, pad, Ho, Wo, vec8));
  } else {
    if (KH == 7 && KW == 7 && stride == 4) CSWIN_CUDA_OK(launch_pdl(im2col_nchw_kernel<__nv_bfloat16, __nv_bfloat16, 7, 4>, dim3(grid), dim3(256), (size_t)(smem), s, (const __nv_bfloat16*)x, (__nv_bfloat16*)col, ldcol, B, C, H, W, KH, KW, stride, pad, Ho, Wo, vec8));
    else CSWIN_CUDA_OK(launch_pdl(im2col_nchw_kernel<__nv_bfloat16, __nv_bfloat16, 0, 0>, dim3(grid), dim3(256), (size_t)(smem), s, (const __nv_bfloat16*)x, (__nv_bfloat16*)col, ldcol, B, C, H, W, KH, KW, stride, pad, Ho, Wo, vec8));
  }
  CSWIN_LAUNCH_CHECK();
  return CSWIN_OK;
}

int carafe_reassemble_fwd(const void* enc, int64_t ldenc, const void* z, int64_t ldz, const void* bias, void* y,
                          int64_t ldy, int nchw_out, int y_is_f32, int B, int H, int W, int C, int up, int dtype,
                          cudaStream_t s) {
  CSWIN_REQUIRE(enc && z && bias && y, CSWIN_ERR_INVALID, "carafe_reassemble: null pointer");
  CSWIN_REQUIRE(up >= 1 && up * up <= 128, CSWIN_ERR_UNSUPPORTED, "carafe_reassemble: up=%d not supported", up);
  CSWIN_REQUIRE(ldenc >= 9 * up * up && ldz >= C, CSWIN_ERR_INVALID, "carafe_reassemble: leading dimension too small");
  const int64_t pixels = (int64_t)B * H * W;
  if (pixels == 0) return CSWIN_OK;
  const size_t smem = sizeof(float) * ((size_t)9 * C + (size_t)up * up * 9);
  CSWIN_REQUIRE(smem <= 48 * 1024, CSWIN_ERR_UNSUPPORTED, "carafe_reassemble: C=%d too large", C);
  if (dtype == CSWIN_BF16 && !nchw_out && !y_is_f32 && up * up <= 32 && (C == 64 || C == 128 || C == 256) && pixels < 0x7fffffff &&
      ldz % 2 == 0 && ldy % 2 == 0 && reinterpret_cast<uintptr_t>(z) % 4 == 0 && reinterpret_cast<uintptr_t>(y) % 4 == 0) {
    const unsigned g2 = (unsigned)ceil_div64(pixels, 8);
    const __nv_bfloat16 *e_ = (const __nv_bfloat16*)enc, *z_ = (const __nv_bfloat16*)z, *b_ = (const __nv_bfloat16*)bias;
    // several pixels per warp with 16-byte loads where the rows allow it (CSWIN_CARAFE_GROUPED=0: the warp-per-pixel kernel, A/B switch)
    static const bool grouped = [] { const char* e = getenv("CSWIN_CARAFE_GROUPED"); return !(e && e[0] == '0'); }();
    const bool al16 = (ldz * 2) % 16 == 0 && (ldy * 2) % 16 == 0 && reinterpret_cast<uintptr_t>(z) % 16 == 0 &&
                      reinterpret_cast<uintptr_t>(y) % 16 == 0 && reinterpret_cast<uintptr_t>(bias) % 16 == 0;
    if (grouped && al16 && (C == 64 || C == 128) && up * up <= C / 8) {
      const int ppw = 32 / (C / 8);
      const unsigned gg = (unsigned)ceil_div64(pixels, 8 * ppw);
      if (C == 64) CSWIN_CUDA_OK(launch_pdl(carafe_reassemble_grp_kernel<8>, dim3(gg), dim3(256), (size_t)(0), s, e_, ldenc, z_, ldz, b_, (__nv_bfloat16*)y, ldy, pixels, H, W, up));
      else CSWIN_CUDA_OK(launch_pdl(carafe_reassemble_grp_kernel<16>, dim3(gg), dim3(256), (size_t)(0), s, e_, ldenc, z_, ldz, b_, (__nv_bfloat16*)y, ldy, pixels, H, W, up));
      CSWIN_LAUNCH_CHECK();
      return CSWIN_OK;
    }
    if (C == 64) CSWIN_CUDA_OK(launch_pdl(carafe_reassemble_warp_kernel<2>, dim3(g2), dim3(256), (size_t)(0), s, e_, ldenc, z_, ldz, b_, (__nv_bfloat16*)y, ldy, pixels, H, W, up));
    else if (C == 128) CSWIN_CUDA_OK(launch_pdl(carafe_reassemble_warp_kernel<4>, dim3(g2), dim3(256), (size_t)(0), s, e_, ldenc, z_, ldz, b_, (__nv_bfloat16*)y, ldy, pixels, H, W, up));
    else CSWIN_CUDA_OK(launch_pdl(carafe_reassemble_warp_kernel<8>, dim3(g2), dim3(256), (size_t)(0), s, e_, ldenc, z_, ldz, b_, (__nv_bfloat16*)y, ldy, pixels, H, W, up));
    CSWIN_LAUNCH_CHECK();
    return CSWIN_OK;
  }
  const unsigned grid = (unsigned)pixels;
  if (dtype == CSWIN_F32) {
    CSWIN_REQUIRE(y_is_f32, CSWIN_ERR_INVALID, "carafe_reassemble: fp32 path writes fp32");
    carafe_reassemble_kernel<float, float><<<grid, 128, smem, s>>>((const float*)enc, ldenc, (const float*)z, ldz, (const float*)bias, (float*)y, ldy, nchw_out, B, H, W, C, up);
  } else if (y_is_f32) {
    carafe_reassemble_kernel<__nv_bfloat16, float><<<grid, 128, smem, s>>>((const __nv_bfloat16*)enc, ldenc, (const __nv_bfloat16*)z, ldz, (const __nv_bfloat16*)bias, (float*)y, ldy, nchw_out, B, H, W, C, up);
  } else {
    carafe_reassemble_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 128, smem, s>>>((const __nv_bfloat16*)enc, ldenc, (const __nv_bfloat16*)z, ldz, (const __nv_bfloat16*)bias, (__nv_bfloat16*)y, ldy, nchw_out, B, H, W, C, up);
  }
  CSWIN_LAUNCH_CHECK();
  return CSWIN_OK;
}

}  // namespace cswin
