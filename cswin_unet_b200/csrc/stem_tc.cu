// stem_tc.cu — the CSWin-UNet stem in ONE launch (bf16 compute, tcgen05 / TMEM / TMA, sm_100a):
//     Conv2d(3, 64, kernel 7, stride 4, padding 2)  ->  'b c h w -> b (h w) c'  ->  LayerNorm(64)
// (networks/cswin_unet.py:338-342, `stage1_conv_embed`), replacing im2col_nchw + Linear + LayerNorm of the composed path
// (22 + 15 + 10 us per forward at batch 24, and a 23 MB column matrix through L2): an implicit GEMM whose A tile is gathered from
// the NCHW image inside the CTA.
//
// CTA = R = floor(128 / Wo) full output rows of one image (112 of the 128 MMA rows at Wo = 56) x all 64 output channels, 256 threads:
//   1. those R output rows need 4 R + 3 input rows x 3 channels: staged as bf16 in shared memory with coalesced loads (fp32 or bf16
//      image; a lane keeps 32 loads in flight), zero padded left / right / top / bottom (the conv's padding = 2) — every input pixel
//      is fetched once per CTA and used by up to four output pixels;
//   2. two threads per output pixel gather its 147 taps (k = (c*7 + ky)*7 + kx, the flattening of conv.weight) from that stage
//      into the UMMA K-major / 128-byte-swizzled A tile [3 k-blocks][128 rows][64] (k = 147..191 are zero);
//   3. one elected thread issues 12 tcgen05.mma (M128 x N64 x K16) against the weight tile W (64 x 192 bf16, TMA) -> 64 fp32
//      columns in TMEM;
//   4. epilogue, one thread per pixel: + bias, LayerNorm over the pixel's 64 channels entirely in registers (two-pass fp32
//      statistics), bf16 store of the row (128 contiguous bytes) and its (sum, sum^2) for the folded LayerNorm of the first block.
#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"

namespace cswin {
namespace {

using namespace tc;

constexpr int kC = 3, kK = 7, kS = 4, kPad = 2, kN = 64;
constexpr int kTaps = kC * kK * kK;                    // 147
constexpr int kKp = 192;                               // K padded to 3 blocks of 64
constexpr int kRows = 128;
constexpr int kThreads = 256;
constexpr int kStageIt = 8;                             // loads per lane and staged row in one batch (256 elements of a row)

struct alignas(64) StemParams {
  CUtensorMap map_w;                                   // (64, 192) bf16, box {64, 64}, 128-byte swizzle
  const void* x; int x_is_f32;
  const float* bias; const float* gamma; const float* beta;
  __nv_bfloat16* out; float* stats;
  int B, H, W, Ho, Wo, span, R, tiles_img;             // span = staged row length = (Wo - 1) * 4 + 7 rounded up to even; R = output
                                                       // rows per tile; tiles_img = ceil(Ho / R)
  int64_t M;
  float eps;
};

__global__ void __launch_bounds__(kThreads, 2) stem_tc_kernel(const __grid_constant__ StemParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* As = smem;                                  // [3][128][128 B]
  uint8_t* Ws = As + 3 * kRows * 128;                  // [3][64][128 B]
  __nv_bfloat16* In = reinterpret_cast<__nv_bfloat16*>(Ws + 3 * kN * 128);      // [3 channels][4 R + 3 input rows][span]
  const int span = P.span;
  const int irows = 4 * P.R + 3;                       // staged input rows per channel
  float* sPar = reinterpret_cast<float*>(In + kC * irows * span + 8);              // bias, gamma, beta [3][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPar + 3 * kN);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  const uint32_t bar_w = smem_u32(&bars[0]), bar_acc = smem_u32(&bars[1]);
  if (warp == 0 && elect_one()) {
    mbar_init(bar_w, 1); mbar_init(bar_acc, 1);
    fence_barrier_init();
    fence_proxy_async();
    tma_prefetch_desc(&P.map_w);
    mbar_expect_tx(bar_w, 3 * kN * 128);
    for (int kb = 0; kb < 3; ++kb) tma_load_2d(smem_u32(Ws + kb * kN * 128), &P.map_w, bar_w, kb * 64, 0);   // weights: no dependency
  }
  if (warp == 1) { tmem_alloc(smem_u32(tmem_slot), 64); tmem_relinquish(); }
  if (tid < 3 * kN) sPar[tid] = tid < kN ? P.bias[tid] : tid < 2 * kN ? P.gamma[tid - kN] : P.beta[tid - 2 * kN];
  pdl_wait();                                          // the image (previous kernel / copy) and `out` are safe from here

  const int Wo = P.Wo, Ho = P.Ho, R = P.R;
  const int b = blockIdx.x / P.tiles_img, oy0 = (blockIdx.x - b * P.tiles_img) * R;
  const int nvalid = min(R, Ho - oy0) * Wo;             // live rows of this tile (tile row r = (oy0 + r / Wo, r % Wo))
  const int64_t m0 = ((int64_t)b * Ho + oy0) * Wo;      // token index of tile row 0
  // ---- 1. stage the input rows: In[c][j][xs] = x[b, c, 4 oy0 - 2 + j, xs - 2] (0 outside); one warp per staged row, lanes along
  //         x (coalesced); a lane fetches its elements of FOUR rows (32 independent loads in flight) before it converts and stores
  //         them: the loads are L2 / HBM latency, issued one at a time they made this the longest phase of the kernel ----
  {
    const int nrows = kC * irows;
    for (int x0 = 0; x0 < span; x0 += 32 * kStageIt) {
      for (int row0 = warp * 4; row0 < nrows; row0 += (kThreads / 32) * 4) {
        float v[4][kStageIt];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const int row = row0 + rr;
          const int c = row / irows, j = row - c * irows;
          const int iy = oy0 * kS - kPad + j;
          const bool rok = row < nrows && iy >= 0 && iy < P.H;
          const int64_t base = (((int64_t)b * kC + c) * P.H + iy) * P.W - kPad;
#pragma unroll
          for (int q = 0; q < kStageIt; ++q) {
            const int xs = x0 + lane + 32 * q;
            const bool ok = rok && xs >= kPad && xs < P.W + kPad;
            v[rr][q] = !ok ? 0.f : P.x_is_f32 ? reinterpret_cast<const float*>(P.x)[base + xs]
                                              : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(P.x)[base + xs]);
          }
        }
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          if (row0 + rr < nrows) {
            __nv_bfloat16* dst = In + (size_t)(row0 + rr) * span;
#pragma unroll
            for (int q = 0; q < kStageIt; ++q) { const int xs = x0 + lane + 32 * q; if (xs < span) dst[xs] = __float2bfloat16_rn(v[rr][q]); }
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- 2. gather the A tile: thread (row r, half h) builds 12 of the row's 24 16-byte chunks; with the chunk index a compile-time
  //         constant every tap is one 2-byte shared-memory load at a fixed offset from its (c, ky) row ----
  {
    const int r = tid & 127, h = tid >> 7;
    const bool live = r < nvalid;
    const int orow = live ? r / Wo : 0, ox = live ? r - orow * Wo : 0;
    // tap (c, ky, kx) of this pixel = In[c][4 orow + ky][4 ox + kx]: row (c * irows + 4 orow + ky) of the stage
    const uint16_t* s16 = reinterpret_cast<const uint16_t*>(In + (size_t)(orow * kS) * span + ox * kS);
    const int cstep = (irows - kK) * span;               // extra offset per channel: s16[(c * 7 + ky) * span + c * cstep + kx]
    const uint32_t arow = smem_u32(As) + r * 128;
    auto gather = [&](auto hc) {
      constexpr int H = decltype(hc)::value;
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        constexpr int dummy = 0; (void)dummy;
        const int ch = H * 12 + j;
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k0 = ch * 8 + 2 * e, k1 = k0 + 1;
          uint32_t lo = 0, hi = 0;
          if (k0 < kTaps) lo = s16[(k0 / kK) * span + (k0 / (kK * kK)) * cstep + (k0 % kK)];
          if (k1 < kTaps) hi = s16[(k1 / kK) * span + (k1 / (kK * kK)) * cstep + (k1 % kK)];
          w[e] = live ? (lo | (hi << 16)) : 0u;
        }
        const int kb = ch >> 3, cc = ch & 7;
        const uint32_t addr = arow + kb * (kRows * 128) + (((cc ^ (r & 7)) & 7) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
      }
    };
    if (h == 0) gather(std::integral_constant<int, 0>{}); else gather(std::integral_constant<int, 1>{});
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // ---- 3. MMA ----
  if (warp == 0 && elect_one()) {
    mbar_wait(bar_w, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(kRows, kN, 0, 0);
#pragma unroll
    for (int kb = 0; kb < 3; ++kb) {
      const uint64_t ad = make_smem_desc(smem_u32(As + kb * kRows * 128), 16, 1024, kLayoutSw128);
      const uint64_t wd = make_smem_desc(smem_u32(Ws + kb * kN * 128), 16, 1024, kLayoutSw128);
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_ss(tmem_base, ad + 2 * k, wd + 2 * k, idesc, (kb | k) != 0);
    }
    tc_commit(bar_acc);
  }
  // ---- 4. epilogue: warps 0..3, thread = pixel (TMEM lane), the 64 channels of the pixel in registers ----
  if (warp < 4) {
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t v0[32], v1[32];
    tmem_ld32(trow, v0);
    tmem_ld32(trow + 32, v1);
    tmem_wait_ld();
    const int er = warp * 32 + lane;                     // tile row of this thread
    const int64_t m = m0 + er;
    float y[64];
    float s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) { y[j] = __uint_as_float(v0[j]) + sPar[j]; y[32 + j] = __uint_as_float(v1[j]) + sPar[32 + j]; }
#pragma unroll
    for (int j = 0; j < 64; ++j) s1 += y[j];
    const float mean = s1 * (1.0f / kN);
    float s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) { const float d = y[j] - mean; s2 = fmaf(d, d, s2); }
    const float rstd = rsqrtf(s2 * (1.0f / kN) + P.eps);
    float t1 = 0.f, t2 = 0.f;
    uint32_t w[32];
#pragma unroll
    for (int j = 0; j < 64; j += 2) {
      const float a = fmaf((y[j] - mean) * rstd, sPar[kN + j], sPar[2 * kN + j]);
      const float b = fmaf((y[j + 1] - mean) * rstd, sPar[kN + j + 1], sPar[2 * kN + j + 1]);
      const uint32_t p = pack_bf16x2(a, b);
      w[j >> 1] = p;
      const float ra = bf16_lo(p), rb = bf16_hi(p);    // statistics of the ROUNDED values: what the next block's GEMM will read
      t1 += ra + rb;
      t2 = fmaf(ra, ra, fmaf(rb, rb, t2));
    }
    if (er < nvalid) {
      uint4* dst = reinterpret_cast<uint4*>(P.out + m * kN);
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
      reinterpret_cast<float2*>(P.stats)[m] = make_float2(t1, t2);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 64);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int stem_fwd_tc(const void* x, int x_is_f32, const void* w_packed, const float* bias, const float* gamma, const float* beta, float eps,
                void* out, float* stats, int B, int H, int W, cudaStream_t stream, bool* handled) {
  *handled = false;
  const int Ho = (H + 2 * kPad - kK) / kS + 1, Wo = (W + 2 * kPad - kK) / kS + 1;
  if (B <= 0 || Ho <= 0 || Wo <= 0 || !aligned16(w_packed) || !aligned16(out) || (reinterpret_cast<uintptr_t>(stats) & 7)) return CSWIN_OK;
  if (tc::encode_tiled_fn() == nullptr) return CSWIN_OK;
  StemParams P;
  {
    const uint64_t dims[2] = {(uint64_t)kKp, (uint64_t)kN}, str[1] = {(uint64_t)kKp * 2};
    const uint32_t box[2] = {64, 64};
    if (!tc::make_tensor_map_bf16(&P.map_w, w_packed, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return CSWIN_ERR_CUDA;
  }
  P.x = x; P.x_is_f32 = x_is_f32; P.bias = bias; P.gamma = gamma; P.beta = beta; P.out = (__nv_bfloat16*)out; P.stats = stats;
  P.B = B; P.H = H; P.W = W; P.Ho = Ho; P.Wo = Wo; P.eps = eps;
  P.span = ((Wo - 1) * kS + kK + 1) & ~1;
  P.M = (int64_t)B * Ho * Wo;
  if (Wo > kRows) return CSWIN_OK;                      // (an output row must fit in one tile: wider images take the composed path)
  P.R = kRows / Wo; P.tiles_img = (Ho + P.R - 1) / P.R;
  const size_t smem = 1024 + 3 * kRows * 128 + 3 * kN * 128 + (size_t)(kC * (4 * P.R + 3) * P.span + 8) * 2 + 3 * kN * 4 + 64;
  if (smem > 200 * 1024) return CSWIN_OK;               // (above 113 KB: one CTA per SM)
  static std::atomic<int> configured{0};
  if (!configured.load(std::memory_order_acquire)) {
    CSWIN_CUDA_OK(cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured.store(1, std::memory_order_release);
  }
  const int64_t ctas = (int64_t)B * P.tiles_img;
  if (ctas >= (1ll << 31)) return CSWIN_OK;
  CSWIN_CUDA_OK(launch_pdl(stem_tc_kernel, dim3((unsigned)ctas), dim3(kThreads), smem, stream, P));
  CSWIN_LAUNCH_CHECK();
  g_tc_launches.fetch_add(1, std::memory_order_relaxed);
  *handled = true;
  return CSWIN_OK;
}

}  // namespace cswin
