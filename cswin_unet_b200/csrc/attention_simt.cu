// attention_simt.cu — fp32-arithmetic LePE stripe attention (forward + backward), CUDA cores only.
//
// This is the exact-arithmetic path (CSWIN_F32: <=1e-4 parity with the fp32 reference) and the general-shape
// path (any head_dim <= 128, any window up to 512 tokens).  The bf16 tensor-core path lives in attention_tc.cu.
//
// One CTA (4 warps) per (batch, window, head) problem.  K and V of the window are staged once in shared memory
// as fp32 [N][d|1] (odd row pitch => conflict-free column walks); each warp then owns query rows n = w, w+4, ...:
// lanes span the kv index for scores / softmax and span the head channels for P.V, the LePE depthwise 3x3
// conv (taken from the same V tile, zero padded at the WINDOW border) and the coalesced store into the (B,L,C)
// concat layout.  Nothing but q,k,v is read from and out written to global memory.
//
// Reference semantics: networks/cswin_unet.py:82-109 (+59-80, 184-202); math in SURVEY.md Appendix A.
#include "common.cuh"

namespace cswin {
namespace {

template <typename T>
struct BranchDev {
  const T* q; const T* k; const T* v; T* out; const T* cw; const T* cb; float* lse;
  int64_t q_bs, q_ts, k_bs, k_ts, v_bs, v_ts, o_bs, o_ts;
  int C_b, heads, hs, ws, nww, nwin, d, prob_begin;
};
template <typename T>
struct AttnParams {
  BranchDev<T> br[2];
  int nb, B, reso;
  float scale;
};

constexpr int kWarps = 4;

template <typename T, int MAXI>
__global__ void __launch_bounds__(kWarps * 32) lepe_attn_fwd_simt_kernel(const AttnParams<T> P) {
  extern __shared__ float smem[];
  const int p = blockIdx.x;
  const int bi = (P.nb > 1 && p >= P.br[1].prob_begin) ? 1 : 0;
  const BranchDev<T>& br = P.br[bi];
  int local = p - br.prob_begin;
  const int head = local % br.heads; local /= br.heads;
  const int win = local % br.nwin;
  const int b = local / br.nwin;
  const int ih = win / br.nww, iw = win % br.nww;
  const int hs = br.hs, ws = br.ws, d = br.d, N = hs * ws, dp = d | 1;
  const int W = P.reso;
  const int ch0 = head * d;

  float* Ks = smem;                    // [N][dp]
  float* Vs = Ks + N * dp;             // [N][dp]
  float* Qs = Vs + N * dp;             // [kWarps][d]
  float* Ps = Qs + kWarps * d;         // [kWarps][N]
  float* Wc = Ps + kWarps * N;         // [d][9]
  float* Bc = Wc + d * 9;              // [d]

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const T* kb = br.k + (int64_t)b * br.k_bs + ch0;
  const T* vb = br.v + (int64_t)b * br.v_bs + ch0;
  for (int i = tid; i < N * d; i += kWarps * 32) {
    const int n = i / d, j = i - n * d;
    const int r = n / ws, c = n - r * ws;
    const int64_t tok = (int64_t)(ih * hs + r) * W + (iw * ws + c);
    Ks[n * dp + j] = ldf(kb + tok * br.k_ts + j);
    Vs[n * dp + j] = ldf(vb + tok * br.v_ts + j);
  }
  for (int i = tid; i < d * 9; i += kWarps * 32) Wc[i] = ldf(br.cw + (int64_t)ch0 * 9 + i);
  for (int i = tid; i < d; i += kWarps * 32) Bc[i] = ldf(br.cb + ch0 + i);
  __syncthreads();

  const T* qb = br.q + (int64_t)b * br.q_bs + ch0;
  T* ob = br.out + (int64_t)b * br.o_bs + ch0;
  float* qs = Qs + w * d;
  float* ps = Ps + w * N;
  for (int n = w; n < N; n += kWarps) {
    const int r = n / ws, c = n - r * ws;
    const int64_t tok = (int64_t)(ih * hs + r) * W + (iw * ws + c);
    for (int j = lane; j < d; j += 32) qs[j] = ldf(qb + tok * br.q_ts + j) * P.scale;   // q*scale first (:98)
    __syncwarp();
    float sv[MAXI];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      const int m = lane + 32 * i;
      float acc = -INFINITY;
      if (m < N) {
        acc = 0.f;
        const float* kr = Ks + m * dp;
        for (int j = 0; j < d; ++j) acc = fmaf(qs[j], kr[j], acc);
      }
      sv[i] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      const int m = lane + 32 * i;
      if (m < N) {
        const float e = expf(sv[i] - mx);
        ps[m] = e;
        sum += e;
      }
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.0f / sum;
    for (int j = lane; j < d; j += 32) {
      float o = 0.f;
      for (int m = 0; m < N; ++m) o = fmaf(ps[m], Vs[m * dp + j], o);
      float lp = Bc[j];
#pragma unroll
      for (int dr = -1; dr <= 1; ++dr) {
#pragma unroll
        for (int dc = -1; dc <= 1; ++dc) {
          const int rr = r + dr, cc = c + dc;
          if (rr >= 0 && rr < hs && cc >= 0 && cc < ws)
            lp = fmaf(Wc[j * 9 + (dr + 1) * 3 + (dc + 1)], Vs[(rr * ws + cc) * dp + j], lp);
        }
      }
      stf(ob + tok * br.o_ts + j, o * inv + lp);
    }
    if (br.lse != nullptr && lane == 0)
      br.lse[((int64_t)b * W * W + tok) * br.heads + head] = mx + logf(sum);
    __syncwarp();
  }
}

template <typename T>
int fill_params(AttnParams<T>& P, const cswin_lepe_branch_t* brs, int nb, int B, int reso, float scale, int* total,
                int* max_n, int* max_d) {
  P.nb = nb; P.B = B; P.reso = reso; P.scale = scale;
  int begin = 0; *max_n = 0; *max_d = 0;
  for (int i = 0; i < nb; ++i) {
    const cswin_lepe_branch_t& s = brs[i];
    CSWIN_REQUIRE(s.q && s.k && s.v && s.out && s.conv_w && s.conv_b, CSWIN_ERR_INVALID, "lepe_attention: null pointer in branch %d", i);
    CSWIN_REQUIRE(s.heads > 0 && s.C_b > 0 && s.C_b % s.heads == 0, CSWIN_ERR_INVALID, "lepe_attention: C_b %d not divisible by heads %d", s.C_b, s.heads);
    CSWIN_REQUIRE(s.H_sp > 0 && s.W_sp > 0 && reso % s.H_sp == 0 && reso % s.W_sp == 0, CSWIN_ERR_INVALID,
                  "lepe_attention: resolution %d not divisible by stripe %dx%d", reso, s.H_sp, s.W_sp);
    BranchDev<T>& d = P.br[i];
    d.q = (const T*)s.q; d.k = (const T*)s.k; d.v = (const T*)s.v; d.out = (T*)s.out;
    d.cw = (const T*)s.conv_w; d.cb = (const T*)s.conv_b; d.lse = s.lse;
    d.q_bs = s.q_bs; d.q_ts = s.q_ts; d.k_bs = s.k_bs; d.k_ts = s.k_ts; d.v_bs = s.v_bs; d.v_ts = s.v_ts;
    d.o_bs = s.o_bs; d.o_ts = s.o_ts;
    d.C_b = s.C_b; d.heads = s.heads; d.hs = s.H_sp; d.ws = s.W_sp; d.nww = reso / s.W_sp;
    d.nwin = (reso / s.H_sp) * (reso / s.W_sp); d.d = s.C_b / s.heads; d.prob_begin = begin;
    begin += B * d.nwin * d.heads;
    if (s.H_sp * s.W_sp > *max_n) *max_n = s.H_sp * s.W_sp;
    if (d.d > *max_d) *max_d = d.d;
  }
  *total = begin;
  return CSWIN_OK;
}

template <typename T>
int launch_fwd(const cswin_lepe_branch_t* brs, int nb, int B, int reso, float scale, cudaStream_t stream) {
  AttnParams<T> P;
  int total, max_n, max_d;
  int rc = fill_params(P, brs, nb, B, reso, scale, &total, &max_n, &max_d);
  if (rc) return rc;
  CSWIN_REQUIRE(max_d <= 128, CSWIN_ERR_UNSUPPORTED, "lepe_attention: head_dim %d > 128 not supported", max_d);
  CSWIN_REQUIRE(max_n <= 512, CSWIN_ERR_UNSUPPORTED, "lepe_attention: window of %d tokens > 512 not supported", max_n);
  const int dp = max_d | 1;
  const size_t smem = sizeof(float) * ((size_t)2 * max_n * dp + kWarps * max_d + (size_t)kWarps * max_n + max_d * 10);
  CSWIN_REQUIRE(smem <= 227 * 1024, CSWIN_ERR_UNSUPPORTED, "lepe_attention: window %d x head_dim %d needs %zu B smem", max_n, max_d, smem);
  if (total == 0) return CSWIN_OK;
  auto kern = (max_n <= 256) ? lepe_attn_fwd_simt_kernel<T, 8> : lepe_attn_fwd_simt_kernel<T, 16>;
  // opt in to large dynamic shared memory ONCE per kernel (not per launch: the call is not stream-capture safe)
  static std::atomic<int> opted[2] = {{0}, {0}};
  if (smem > 48 * 1024 && !opted[max_n <= 256].exchange(1))
    CSWIN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  kern<<<total, kWarps * 32, smem, stream>>>(P);
  CSWIN_LAUNCH_CHECK();
  return CSWIN_OK;
}


// ------------------------------------------------------------------------------------------------------------------
// backward (autograd of networks/cswin_unet.py:82-109; formulas: SURVEY.md Appendix A).  One CTA (8 warps) per
// (batch, window, head).  Q, K, V and the upstream gradient G of the window are staged in shared memory as fp32.
//   pass A (one warp per query row n):  P_n = softmax(scale q_n K^T) recomputed, dP = G_n V^T, delta_n = <P_n, dP>,
//           dS = P o (dP - delta), dQ_n = scale dS K;  lse_n and delta_n are kept in shared memory;
//   pass B (one warp per key row m):    column m of P and dS recomputed from lse / delta,
//           dK_m = scale dS^T Q, dV_m = P^T G + depthwise-conv-transpose(G) with WINDOW-local zero padding;
//   LePE parameters: d w[j,t] = sum_n G[n,j] V^[n+t, j], d b[j] = sum_n G[n,j], reduced per CTA then fp32 atomics.
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
struct BranchGradDev {
  BranchDev<T> f;
  const T* g; int64_t g_bs, g_ts;
  T* dq; T* dk; T* dv;
  int64_t dq_bs, dq_ts, dk_bs, dk_ts, dv_bs, dv_ts;
  float* dcw; float* dcb;
};
template <typename T>
struct AttnGradParams {
  BranchGradDev<T> br[2];
  int nb, B, reso;
  float scale;
};

constexpr int kBwdWarps = 8;

template <typename T, int MAXI>
__global__ void __launch_bounds__(kBwdWarps * 32) lepe_attn_bwd_simt_kernel(const AttnGradParams<T> P) {
  extern __shared__ float smem[];
  const int p = blockIdx.x;
  const int bi = (P.nb > 1 && p >= P.br[1].f.prob_begin) ? 1 : 0;
  const BranchGradDev<T>& bg = P.br[bi];
  const BranchDev<T>& br = bg.f;
  int local = p - br.prob_begin;
  const int head = local % br.heads; local /= br.heads;
  const int win = local % br.nwin;
  const int b = local / br.nwin;
  const int ih = win / br.nww, iw = win % br.nww;
  const int hs = br.hs, ws = br.ws, d = br.d, N = hs * ws, dp = d | 1;
  const int W = P.reso;
  const int ch0 = head * d;

  float* Qs = smem;                       // [N][dp]
  float* Ks = Qs + N * dp;
  float* Vs = Ks + N * dp;
  float* Gs = Vs + N * dp;
  float* Lse = Gs + N * dp;               // [N]
  float* Dl = Lse + N;                    // [N]
  float* Ab = Dl + N;                     // [kBwdWarps][N]  dS row / column
  float* Bb = Ab + kBwdWarps * N;         // [kBwdWarps][N]  P column
  float* Wc = Bb + kBwdWarps * N;         // [d][9]
  float* Part = Wc + d * 9;               // [kBwdWarps][d][10] partial d w / d b

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  auto tok_of = [&](int n) -> int64_t { const int r = n / ws, c = n - r * ws; return (int64_t)(ih * hs + r) * W + (iw * ws + c); };
  for (int i = tid; i < N * d; i += kBwdWarps * 32) {
    const int n = i / d, j = i - n * d;
    const int64_t tok = tok_of(n);
    Qs[n * dp + j] = ldf(br.q + (int64_t)b * br.q_bs + tok * br.q_ts + ch0 + j);
    Ks[n * dp + j] = ldf(br.k + (int64_t)b * br.k_bs + tok * br.k_ts + ch0 + j);
    Vs[n * dp + j] = ldf(br.v + (int64_t)b * br.v_bs + tok * br.v_ts + ch0 + j);
    Gs[n * dp + j] = ldf(bg.g + (int64_t)b * bg.g_bs + tok * bg.g_ts + ch0 + j);
  }
  for (int i = tid; i < d * 9; i += kBwdWarps * 32) Wc[i] = ldf(br.cw + (int64_t)ch0 * 9 + i);
  __syncthreads();

  float* ab = Ab + w * N;
  float* bb = Bb + w * N;
  // ---- pass A: rows ----
  for (int n = w; n < N; n += kBwdWarps) {
    float sv[MAXI], dv_[MAXI];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      const int m = lane + 32 * i;
      float acc = -INFINITY, accd = 0.f;
      if (m < N) {
        acc = 0.f;
        for (int j = 0; j < d; ++j) {
          acc = fmaf(Qs[n * dp + j], Ks[m * dp + j], acc);
          accd = fmaf(Gs[n * dp + j], Vs[m * dp + j], accd);
        }
        acc *= P.scale;
      }
      sv[i] = acc; dv_[i] = accd;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXI; ++i) if (lane + 32 * i < N) sum += expf(sv[i] - mx);
    sum = warp_sum(sum);
    const float lse = mx + logf(sum);
    float delta = 0.f;
#pragma unroll
    for (int i = 0; i < MAXI; ++i) if (lane + 32 * i < N) { sv[i] = expf(sv[i] - lse); delta = fmaf(sv[i], dv_[i], delta); }
    delta = warp_sum(delta);
#pragma unroll
    for (int i = 0; i < MAXI; ++i) { const int m = lane + 32 * i; if (m < N) ab[m] = sv[i] * (dv_[i] - delta); }
    if (lane == 0) { Lse[n] = lse; Dl[n] = delta; }
    __syncwarp();
    const int64_t tok = tok_of(n);
    for (int j = lane; j < d; j += 32) {
      float acc = 0.f;
      for (int m = 0; m < N; ++m) acc = fmaf(ab[m], Ks[m * dp + j], acc);
      stf(bg.dq + (int64_t)b * bg.dq_bs + tok * bg.dq_ts + ch0 + j, acc * P.scale);
    }
    __syncwarp();
  }
  __syncthreads();
  // ---- pass B: columns ----
  for (int m = w; m < N; m += kBwdWarps) {
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      const int n = lane + 32 * i;
      if (n < N) {
        float acc = 0.f, accd = 0.f;
        for (int j = 0; j < d; ++j) {
          acc = fmaf(Qs[n * dp + j], Ks[m * dp + j], acc);
          accd = fmaf(Gs[n * dp + j], Vs[m * dp + j], accd);
        }
        const float pr = expf(acc * P.scale - Lse[n]);
        bb[n] = pr;
        ab[n] = pr * (accd - Dl[n]);
      }
    }
    __syncwarp();
    const int rm = m / ws, cm = m - rm * ws;
    const int64_t tok = tok_of(m);
    for (int j = lane; j < d; j += 32) {
      float dk = 0.f, dvv = 0.f;
      for (int n = 0; n < N; ++n) {
        dk = fmaf(ab[n], Qs[n * dp + j], dk);
        dvv = fmaf(bb[n], Gs[n * dp + j], dvv);
      }
      // conv-transpose: output position (rm - dr, cm - dc) read V at (rm, cm) through tap (dr, dc)
#pragma unroll
      for (int dr = -1; dr <= 1; ++dr)
#pragma unroll
        for (int dc = -1; dc <= 1; ++dc) {
          const int rr = rm - dr, cc = cm - dc;
          if (rr >= 0 && rr < hs && cc >= 0 && cc < ws)
            dvv = fmaf(Wc[j * 9 + (dr + 1) * 3 + (dc + 1)], Gs[(rr * ws + cc) * dp + j], dvv);
        }
      stf(bg.dk + (int64_t)b * bg.dk_bs + tok * bg.dk_ts + ch0 + j, dk * P.scale);
      stf(bg.dv + (int64_t)b * bg.dv_bs + tok * bg.dv_ts + ch0 + j, dvv);
    }
    __syncwarp();
  }
  // ---- LePE parameter gradients (skipped when the caller computes them with cswin_lepe_param_grad) ----
  if (bg.dcw == nullptr) return;                         // CTA-uniform
  for (int jj = 0; jj < d; jj += 32) {
    const int j = jj + lane;
    float acc[10];
#pragma unroll
    for (int t = 0; t < 10; ++t) acc[t] = 0.f;
    if (j < d) {
      for (int n = w; n < N; n += kBwdWarps) {
        const int r = n / ws, c = n - r * ws;
        const float g = Gs[n * dp + j];
        acc[9] += g;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int rr = r + t / 3 - 1, cc = c + t % 3 - 1;
          if (rr >= 0 && rr < hs && cc >= 0 && cc < ws) acc[t] = fmaf(g, Vs[(rr * ws + cc) * dp + j], acc[t]);
        }
      }
#pragma unroll
      for (int t = 0; t < 10; ++t) Part[(w * d + j) * 10 + t] = acc[t];
    }
  }
  __syncthreads();
  for (int i = tid; i < d * 10; i += kBwdWarps * 32) {
    float s = 0.f;
    for (int ww = 0; ww < kBwdWarps; ++ww) s += Part[ww * d * 10 + i];
    const int j = i / 10, t = i - j * 10;
    if (t < 9) atomicAdd(bg.dcw + (int64_t)(ch0 + j) * 9 + t, s);
    else atomicAdd(bg.dcb + ch0 + j, s);
  }
}

template <typename T>
int launch_bwd(const cswin_lepe_branch_grad_t* gs, int nb, int B, int reso, float scale, cudaStream_t stream) {
  cswin_lepe_branch_t fwd[2] = {};
  for (int i = 0; i < nb; ++i) {
    fwd[i] = gs[i].fwd;
    if (fwd[i].out == nullptr) fwd[i].out = const_cast<void*>(gs[i].dout);          // `out` itself is not needed
  }
  AttnParams<T> F;
  int total, max_n, max_d;
  int rc = fill_params(F, fwd, nb, B, reso, scale, &total, &max_n, &max_d);
  if (rc) return rc;
  AttnGradParams<T> P;
  P.nb = nb; P.B = B; P.reso = reso; P.scale = scale;
  for (int i = 0; i < nb; ++i) {
    const cswin_lepe_branch_grad_t& s = gs[i];
    CSWIN_REQUIRE(s.dout && s.dq && s.dk && s.dv && ((s.dconv_w != nullptr) == (s.dconv_b != nullptr)), CSWIN_ERR_INVALID, "lepe_attention_bwd: null pointer in branch %d", i);
    BranchGradDev<T>& d = P.br[i];
    d.f = F.br[i];
    d.g = (const T*)s.dout; d.g_bs = s.do_bs; d.g_ts = s.do_ts;
    d.dq = (T*)s.dq; d.dk = (T*)s.dk; d.dv = (T*)s.dv;
    d.dq_bs = s.dq_bs; d.dq_ts = s.dq_ts; d.dk_bs = s.dk_bs; d.dk_ts = s.dk_ts; d.dv_bs = s.dv_bs; d.dv_ts = s.dv_ts;
    d.dcw = s.dconv_w; d.dcb = s.dconv_b;
  }
  if (nb == 1) P.br[1] = P.br[0];
  CSWIN_REQUIRE(max_d <= 128, CSWIN_ERR_UNSUPPORTED, "lepe_attention_bwd: head_dim %d > 128 not supported", max_d);
  const int dp = max_d | 1;
  const size_t smem = sizeof(float) * ((size_t)4 * max_n * dp + 2 * max_n + (size_t)2 * kBwdWarps * max_n + max_d * 9 +
                                       (size_t)kBwdWarps * max_d * 10);
  CSWIN_REQUIRE(max_n <= 512 && smem <= 227 * 1024, CSWIN_ERR_UNSUPPORTED,
                "lepe_attention_bwd: window %d x head_dim %d needs %zu B smem (limit 227 KB)", max_n, max_d, smem);
  if (total == 0) return CSWIN_OK;
  auto kern = (max_n <= 256) ? lepe_attn_bwd_simt_kernel<T, 8> : lepe_attn_bwd_simt_kernel<T, 16>;
  static std::atomic<int> opted[2] = {{0}, {0}};
  if (smem > 48 * 1024 && !opted[max_n <= 256].exchange(1))
    CSWIN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  kern<<<total, kBwdWarps * 32, smem, stream>>>(P);
  CSWIN_LAUNCH_CHECK();
  return CSWIN_OK;
}

}  // namespace

int lepe_attention_fwd_simt(const cswin_lepe_branch_t* br, int nb, int B, int reso, float scale, int dtype, cudaStream_t s) {
  if (dtype == CSWIN_F32) return launch_fwd<float>(br, nb, B, reso, scale, s);
  return launch_fwd<__nv_bfloat16>(br, nb, B, reso, scale, s);
}

}  // namespace cswin

namespace cswin {
int lepe_attention_bwd_simt(const cswin_lepe_branch_grad_t* br, int nb, int B, int reso, float scale, int dtype, cudaStream_t s) {
  if (dtype == CSWIN_F32) return launch_bwd<float>(br, nb, B, reso, scale, s);
  return launch_bwd<__nv_bfloat16>(br, nb, B, reso, scale, s);
}
}  // namespace cswin
