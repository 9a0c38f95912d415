// resample.cu — the two scipy.ndimage.zoom calls of the reference's evaluation loop on the GPU:
//   utils.py:69   slice = zoom(slice, (P / x, P / y), order=3)      float32 (x, y) -> (P, P), cubic B-spline
//   utils.py:77   pred  = zoom(out,   (x / P, y / P), order=0)      label map (P, P) -> (x, y), nearest
// with scipy's defaults mode='constant', cval=0, prefilter=True, grid_mode=False, reproduced operation by operation in
// float64 (oracle/zoom_oracle.py is the CPU restatement, pinned against scipy itself):
//   coordinate of output o along an axis: cc = o * ((in - 1) / (out - 1)); cc outside [0, in - 1] (as float64 — with
//   512 -> 224 the last one is 511.00000000000006) gives cval = 0: scipy's last row / column are zero and so are ours;
//   order 3: recursive prefilter (pole sqrt(3) - 2, gain 6, mirror initialisation) along H, then along W, 4 x 4 taps with
//   mirrored edge indices; order 0: index floor(cc + 0.5).
// Both are bandwidth kernels: the prefilter is one thread per line (the recursion is sequential), the interpolation one thread
// per output pixel.  On the CPU the two calls cost ~15 ms per 512^2 slice, i.e. 300 x the network's share of a slice.
#include <algorithm>

#include "common.cuh"

namespace cswin {
namespace {

// One thread per line: c = prefilter(src line) in float64.  src may be the float32 image (first axis) or the float64 work
// buffer itself (second axis, in place).  Element i of line l lives at base(l) + i * estride.
template <typename TIn>
__global__ void __launch_bounds__(128) spline_prefilter_kernel(const TIn* src, double* dst, int64_t n_lines,   // (src may alias dst)
                                                               int lines_per_img, int64_t img_stride, int64_t lstride, int64_t estride,
                                                               int n) {
  const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_lines) return;
  const int64_t base = (l / lines_per_img) * img_stride + (l % lines_per_img) * lstride;
  const TIn* s = src + base;
  double* c = dst + base;
  if (n < 2) { if (n == 1) c[0] = (double)s[0]; return; }
  const double z = sqrt(3.0) - 2.0;                       // the cubic B-spline pole, formed as scipy forms it
  const double gain = __dmul_rn(1.0 - z, 1.0 - 1.0 / z);  // (1 - z)(1 - 1/z) = 6 up to rounding, as scipy forms it
  const double z_n_1 = pow(z, (double)(n - 1));
  // causal initialisation with mirror boundaries (scipy _init_causal_mirror)
  double c0 = __dadd_rn(__dmul_rn((double)s[0], gain), __dmul_rn(z_n_1, __dmul_rn((double)s[(int64_t)(n - 1) * estride], gain)));
  double z_i = z;
  for (int i = 1; i < n - 1; ++i) {
    const double a = __dmul_rn((double)s[(int64_t)i * estride], gain), b = __dmul_rn((double)s[(int64_t)(n - 1 - i) * estride], gain);
    c0 = __dadd_rn(c0, __dmul_rn(z_i, __dadd_rn(a, __dmul_rn(z_n_1, b))));
    z_i = __dmul_rn(z_i, z);
  }
  double prev = __ddiv_rn(c0, __dsub_rn(1.0, __dmul_rn(z_n_1, z_n_1)));
  c[0] = prev;
  for (int i = 1; i < n; ++i) {                           // causal pass
    prev = __dadd_rn(__dmul_rn((double)s[(int64_t)i * estride], gain), __dmul_rn(z, prev));
    c[(int64_t)i * estride] = prev;
  }
  // anticausal initialisation (scipy _init_anticausal_mirror) and pass
  double nxt = __dmul_rn(__dadd_rn(__dmul_rn(z, c[(int64_t)(n - 2) * estride]), c[(int64_t)(n - 1) * estride]),
                         __ddiv_rn(z, __dsub_rn(__dmul_rn(z, z), 1.0)));
  c[(int64_t)(n - 1) * estride] = nxt;
  for (int i = n - 2; i >= 0; --i) {
    nxt = __dmul_rn(z, __dsub_rn(nxt, c[(int64_t)i * estride]));
    c[(int64_t)i * estride] = nxt;
  }
}

struct Taps { int idx[4]; double w[4]; bool inside; };

__device__ __forceinline__ Taps cubic_taps(int o, int n_in, double zf) {
  Taps t;
  const double cc = __dmul_rn((double)o, zf);
  t.inside = cc >= 0.0 && cc <= (double)(n_in - 1);
  const double fl = floor(cc);
  const double x = __dsub_rn(cc, fl), x2 = __dmul_rn(x, x), x3 = __dmul_rn(x2, x);
  const double u = __dsub_rn(1.0, x);
  t.w[0] = __ddiv_rn(__dmul_rn(__dmul_rn(u, u), u), 6.0);
  t.w[1] = __ddiv_rn(__dadd_rn(__dsub_rn(__dmul_rn(3.0, x3), __dmul_rn(6.0, x2)), 4.0), 6.0);
  t.w[2] = __ddiv_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(-3.0, x3), __dmul_rn(3.0, x2)), __dmul_rn(3.0, x)), 1.0), 6.0);
  t.w[3] = __ddiv_rn(x3, 6.0);
  const int st = (int)fl - 1;
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    int i = st + l;
    if (i < 0) i = -i;
    else if (i >= n_in) i = 2 * n_in - 2 - i;
    t.idx[l] = min(max(i, 0), n_in - 1);
  }
  return t;
}

// out[(s, rep, oy, ox)] = sum_{a,b} wy[a] wx[b] c[s, iy[a], ix[b]]  (0 outside), written `reps` times (1 -> 3 channel repeat
// of vision_transformer.py:40-41 for free)
__global__ void __launch_bounds__(256) zoom_cubic_kernel(const double* __restrict__ c, int n, int H, int W, float* __restrict__ out,
                                                         int64_t out_ns, int64_t out_cs, int reps, int OH, int OW, double zfy, double zfx) {
  const int64_t total = (int64_t)n * OH * OW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % OW), oy = (int)((i / OW) % OH), s = (int)(i / ((int64_t)OW * OH));
    const Taps ty = cubic_taps(oy, H, zfy), tx = cubic_taps(ox, W, zfx);
    double acc = 0.0;
    if (ty.inside && tx.inside) {
      const double* cs = c + (int64_t)s * H * W;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
          acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(ty.w[a], tx.w[b]), cs[(int64_t)ty.idx[a] * W + tx.idx[b]]));
    }
    const float v = (float)acc;
    for (int r = 0; r < reps; ++r) out[(int64_t)s * out_ns + (int64_t)r * out_cs + (int64_t)oy * OW + ox] = v;
  }
}

__global__ void __launch_bounds__(256) zoom_nearest_u8_kernel(const uint8_t* __restrict__ in, int n, int H, int W, uint8_t* __restrict__ out,
                                                              int OH, int OW, double zfy, double zfx) {
  const int64_t total = (int64_t)n * OH * OW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % OW), oy = (int)((i / OW) % OH), s = (int)(i / ((int64_t)OW * OH));
    const double cy = __dmul_rn((double)oy, zfy), cx = __dmul_rn((double)ox, zfx);
    uint8_t v = 0;
    if (cy >= 0.0 && cy <= (double)(H - 1) && cx >= 0.0 && cx <= (double)(W - 1)) {
      const int iy = min(max((int)floor(__dadd_rn(cy, 0.5)), 0), H - 1), ix = min(max((int)floor(__dadd_rn(cx, 0.5)), 0), W - 1);
      v = in[((int64_t)s * H + iy) * W + ix];
    }
    out[i] = v;
  }
}

unsigned grid_for(int64_t total, int per_cta) {
  return (unsigned)std::min<int64_t>((total + per_cta - 1) / per_cta, (int64_t)sm_count() * 16);
}

double zoom_factor(int n_in, int n_out) { return n_out > 1 ? (double)(n_in - 1) / (double)(n_out - 1) : 1.0; }

}  // namespace

int zoom_cubic(const float* in, int n, int H, int W, double* work, float* out, int64_t out_ns, int64_t out_cs, int reps, int OH,
               int OW, cudaStream_t s) {
  CSWIN_REQUIRE(in && work && out, CSWIN_ERR_INVALID, "zoom_cubic: null pointer");
  CSWIN_REQUIRE(n >= 0 && H > 0 && W > 0 && OH > 0 && OW > 0 && reps >= 1, CSWIN_ERR_INVALID, "zoom_cubic: bad shape");
  if (n == 0) return CSWIN_OK;
  const int64_t img = (int64_t)H * W;
  {   // along H (axis 0 of a slice): one line per (slice, column); neighbouring threads = neighbouring columns (coalesced)
    const int64_t lines = (int64_t)n * W;
    spline_prefilter_kernel<float><<<(unsigned)((lines + 127) / 128), 128, 0, s>>>(in, work, lines, W, img, 1, W, H);
    CSWIN_LAUNCH_CHECK();
  }
  {   // along W, in place on the float64 coefficients
    const int64_t lines = (int64_t)n * H;
    spline_prefilter_kernel<double><<<(unsigned)((lines + 127) / 128), 128, 0, s>>>(work, work, lines, H, img, W, 1, W);
    CSWIN_LAUNCH_CHECK();
  }
  const int64_t total = (int64_t)n * OH * OW;
  zoom_cubic_kernel<<<grid_for(total, 256), 256, 0, s>>>(work, n, H, W, out, out_ns, out_cs, reps, OH, OW, zoom_factor(H, OH), zoom_factor(W, OW));
  CSWIN_LAUNCH_CHECK();
  return CSWIN_OK;
}

int zoom_nearest_u8(const uint8_t* in, int n, int H, int W, uint8_t* out, int OH, int OW, cudaStream_t s) {
  CSWIN_REQUIRE(in && out, CSWIN_ERR_INVALID, "zoom_nearest_u8: null pointer");
  CSWIN_REQUIRE(n >= 0 && H > 0 && W > 0 && OH > 0 && OW > 0, CSWIN_ERR_INVALID, "zoom_nearest_u8: bad shape");
  if (n == 0) return CSWIN_OK;
  const int64_t total = (int64_t)n * OH * OW;
  zoom_nearest_u8_kernel<<<grid_for(total, 256), 256, 0, s>>>(in, n, H, W, out, OH, OW, zoom_factor(H, OH), zoom_factor(W, OW));
  CSWIN_LAUNCH_CHECK();
  return CSWIN_OK;
}

}  // namespace cswin
