// resample.cu — the two scipy.ndimage.zoom calls of the reference's evaluation loop on the GPU:
//   utils.py:69   slice = zoom(slice, (P / x, P / y), order=3)      float32 (x, y) -> (P, P), cubic B-spline
//   utils.py:77   pred  = zoom(out,   (x / P, y / P), order=0)      label map (P, P) -> (x, y), nearest
// with scipy's defaults mode='constant', cval=0, prefilter=True, grid_mode=False, reproduced operation by operation in
// float64 (oracle/zoom_oracle.py is the CPU restatement, pinned against scipy itself):
//   coordinate of output o along an axis: cc = o * ((in - 1) / (out - 1)); cc outside [0, in - 1] (as float64 — with
//   512 -> 224 the last one is 511.00000000000006) gives cval = 0: scipy's last row / column are zero and so are ours;
//   order 3: recursive prefilter (pole sqrt(3) - 2, gain 6, mirror initialisation) along H, then along W, 4 x 4 taps with
//   mirrored edge indices; order 0: index floor(cc + 0.5).
// Both are bandwidth kernels: the prefilter is one thread per line (the recursion is sequential; loads are batched / staged through
// shared memory so that they stay off the dependent chain), the interpolation one thread per output pixel.  On the CPU the two calls cost ~15 ms per 512^2 slice, i.e. 300 x the network's share of a slice.
#include <algorithm>

#include "common.cuh"

namespace cswin {
namespace {

// c = prefilter(src line) in float64, one thread per line; the recursion along a line is sequential and its operation order is
// scipy's (bit-compatible results), so the parallelism is across lines and the job of the kernels is to keep the loads off the
// dependent chain: the first version issued one (uncoalesced, for rows) load per recursion step and was latency-bound at ~0.6 us
// per step (322 + 259 us for 8 slices of 512^2, profiles/r02_ncu_rest_kernels_summary.txt).
constexpr int kPfChunk = 8;

struct PfConst { double z, gain, z_n_1, anti; };
__device__ __forceinline__ PfConst pf_const(int n) {
  PfConst k;
  k.z = sqrt(3.0) - 2.0;                                   // the cubic B-spline pole, formed as scipy forms it
  k.gain = __dmul_rn(1.0 - k.z, 1.0 - 1.0 / k.z);          // (1 - z)(1 - 1/z) = 6 up to rounding, as scipy forms it
  k.z_n_1 = pow(k.z, (double)(n - 1));
  k.anti = __ddiv_rn(k.z, __dsub_rn(__dmul_rn(k.z, k.z), 1.0));
  return k;
}

// Lines whose neighbours are adjacent in memory (lane l <-> line l, element i at base + i * estride): along H of a slice.  Every
// access of a warp is one contiguous segment; elements are fetched kPfChunk at a time before the dependent steps that use them.
template <typename TIn>
__global__ void __launch_bounds__(32) spline_prefilter_cols_kernel(const TIn* src, double* dst, int64_t n_lines, int lines_per_img,
                                                                   int64_t img_stride, int64_t estride, int n) {
  const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_lines) return;
  const int64_t base = (l / lines_per_img) * img_stride + (l % lines_per_img);
  const TIn* s = src + base;
  double* c = dst + base;
  if (n < 2) { if (n == 1) c[0] = (double)s[0]; return; }
  const PfConst k = pf_const(n);
  // causal initialisation with mirror boundaries (scipy _init_causal_mirror)
  double c0 = __dadd_rn(__dmul_rn((double)s[0], k.gain), __dmul_rn(k.z_n_1, __dmul_rn((double)s[(int64_t)(n - 1) * estride], k.gain)));
  double z_i = k.z;
  for (int i0 = 1; i0 < n - 1; i0 += kPfChunk) {
    double a[kPfChunk], b[kPfChunk];
#pragma unroll
    for (int j = 0; j < kPfChunk; ++j) {
      const int i = i0 + j;
      a[j] = i < n - 1 ? (double)s[(int64_t)i * estride] : 0.0;
      b[j] = i < n - 1 ? (double)s[(int64_t)(n - 1 - i) * estride] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < kPfChunk; ++j) {
      if (i0 + j < n - 1) {
        c0 = __dadd_rn(c0, __dmul_rn(z_i, __dadd_rn(__dmul_rn(a[j], k.gain), __dmul_rn(k.z_n_1, __dmul_rn(b[j], k.gain)))));
        z_i = __dmul_rn(z_i, k.z);
      }
    }
  }
  double prev = __ddiv_rn(c0, __dsub_rn(1.0, __dmul_rn(k.z_n_1, k.z_n_1)));
  c[0] = prev;
  for (int i0 = 1; i0 < n; i0 += kPfChunk) {              // causal pass
    double a[kPfChunk];
#pragma unroll
    for (int j = 0; j < kPfChunk; ++j) a[j] = i0 + j < n ? (double)s[(int64_t)(i0 + j) * estride] : 0.0;
#pragma unroll
    for (int j = 0; j < kPfChunk; ++j) {
      if (i0 + j < n) {
        prev = __dadd_rn(__dmul_rn(a[j], k.gain), __dmul_rn(k.z, prev));
        c[(int64_t)(i0 + j) * estride] = prev;
      }
    }
  }
  // anticausal initialisation (scipy _init_anticausal_mirror) and pass; c[n-1] == prev
  double nxt = __dmul_rn(__dadd_rn(__dmul_rn(k.z, c[(int64_t)(n - 2) * estride]), prev), k.anti);
  c[(int64_t)(n - 1) * estride] = nxt;
  for (int i0 = n - 2; i0 >= 0; i0 -= kPfChunk) {
    double a[kPfChunk];
#pragma unroll
    for (int j = 0; j < kPfChunk; ++j) a[j] = i0 - j >= 0 ? c[(int64_t)(i0 - j) * estride] : 0.0;
#pragma unroll
    for (int j = 0; j < kPfChunk; ++j) {
      if (i0 - j >= 0) {
        nxt = __dmul_rn(k.z, __dsub_rn(nxt, a[j]));
        c[(int64_t)(i0 - j) * estride] = nxt;
      }
    }
  }
}

// Lines that are contiguous in memory (a row of a slice), in place on the float64 coefficients: one warp owns 32 consecutive rows and
// moves them through shared memory in 32-column tiles (coalesced 256-byte row segments in, the same out), lane l runs the recursion
// of row l on its tile row.  Tile rows are padded to 33 doubles: the 16 lanes of a half-warp hit 16 different bank pairs.
constexpr int kPfT = 32;
__global__ void __launch_bounds__(32) spline_prefilter_rows_kernel(double* c, int64_t n_lines, int64_t lstride, int n) {
  __shared__ double ta[kPfT][kPfT + 1], tb[kPfT][kPfT + 1];
  const int lane = threadIdx.x;
  const int64_t l0 = (int64_t)blockIdx.x * kPfT;           // first line of this warp (lines of consecutive slices are contiguous)
  const int rows = (int)min((int64_t)kPfT, n_lines - l0);
  double* base = c + l0 * lstride;
  if (n < 2) return;                                        // (a single element is its own coefficient: nothing to do in place)
  const PfConst k = pf_const(n);
  const bool live = lane < rows;
  auto load_tile = [&](double (*t)[kPfT + 1], int col0) {   // tile[r][j] = line r, element col0 + j
    for (int r = 0; r < rows; ++r) { const int col = col0 + lane; t[r][lane] = col < n && col >= 0 ? base[(int64_t)r * lstride + col] : 0.0; }
    __syncwarp();
  };
  auto store_tile = [&](double (*t)[kPfT + 1], int col0) {
    __syncwarp();
    for (int r = 0; r < rows; ++r) { const int col = col0 + lane; if (col < n && col >= 0) base[(int64_t)r * lstride + col] = t[r][lane]; }
    __syncwarp();
  };
  // ---- causal initialisation: c0 = s[0] g + z^(n-1) s[n-1] g + sum_{i=1}^{n-2} z^i (s[i] g + z^(n-1) s[n-1-i] g), i ascending ----
  double c0 = 0.0, z_i = k.z;
  for (int col0 = 0; col0 < n; col0 += kPfT) {
    // element i = col0 + j pairs with element n-1-i = (n-1-col0) - j: the mirror tile starts at n-1-col0-(kPfT-1) and is read backwards
    load_tile(ta, col0);
    load_tile(tb, n - 1 - col0 - (kPfT - 1));
    if (live) {
      for (int j = 0; j < kPfT; ++j) {
        const int i = col0 + j;
        if (i >= n - 1) break;
        const double a = ta[lane][j], b = tb[lane][kPfT - 1 - j];
        if (i == 0) c0 = __dadd_rn(__dmul_rn(a, k.gain), __dmul_rn(k.z_n_1, __dmul_rn(b, k.gain)));
        else {
          c0 = __dadd_rn(c0, __dmul_rn(z_i, __dadd_rn(__dmul_rn(a, k.gain), __dmul_rn(k.z_n_1, __dmul_rn(b, k.gain)))));
          z_i = __dmul_rn(z_i, k.z);
        }
      }
    }
    __syncwarp();
  }
  double prev = __ddiv_rn(c0, __dsub_rn(1.0, __dmul_rn(k.z_n_1, k.z_n_1)));
  // ---- causal pass ----
  for (int col0 = 0; col0 < n; col0 += kPfT) {
    load_tile(ta, col0);
    if (live) {
      for (int j = 0; j < kPfT && col0 + j < n; ++j) {
        if (col0 + j > 0) prev = __dadd_rn(__dmul_rn(ta[lane][j], k.gain), __dmul_rn(k.z, prev));
        ta[lane][j] = prev;
      }
    }
    store_tile(ta, col0);
  }
  // ---- anticausal initialisation and pass (tiles from the last one backwards; the last tile may be ragged) ----
  double nxt = 0.0;
  const int last0 = ((n - 1) / kPfT) * kPfT;
  for (int col0 = last0; col0 >= 0; col0 -= kPfT) {
    load_tile(ta, col0);
    if (col0 == last0 && (n - 1) - last0 == 0) load_tile(tb, col0 - kPfT);     // c[n-2] lives in the previous tile
    if (live) {
      for (int j = min(kPfT - 1, n - 1 - col0); j >= 0; --j) {
        const int i = col0 + j;
        if (i == n - 1) {
          const double cm2 = j > 0 ? ta[lane][j - 1] : tb[lane][kPfT - 1];
          nxt = __dmul_rn(__dadd_rn(__dmul_rn(k.z, cm2), ta[lane][j]), k.anti);
        } else {
          nxt = __dmul_rn(k.z, __dsub_rn(nxt, ta[lane][j]));
        }
        ta[lane][j] = nxt;
      }
    }
    store_tile(ta, col0);
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
    spline_prefilter_cols_kernel<float><<<(unsigned)((lines + 31) / 32), 32, 0, s>>>(in, work, lines, W, img, W, H);
    CSWIN_LAUNCH_CHECK();
  }
  {   // along W, in place on the float64 coefficients
    const int64_t lines = (int64_t)n * H;
    spline_prefilter_rows_kernel<<<(unsigned)((lines + 31) / 32), 32, 0, s>>>(work, lines, W, W);   // rows of consecutive slices are contiguous
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
