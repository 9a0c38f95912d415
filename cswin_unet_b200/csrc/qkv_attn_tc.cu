// qkv_attn_tc.cu — [LayerNorm (folded) -> qkv Linear -> LePE stripe attention of both branches] as ONE tcgen05 kernel (sm_100a).
//
// Replaces, per CSWinBlock.forward (networks/cswin_unet.py:160-181): `img = norm1(x)` :168, `qkv = self.qkv(img)` :169, the
// reshape / permute view of :169, both `self.attns[i](qkv[..., slice])` calls :172-176 with everything LePEAttention.forward does
// (:82-109: im2cswin :59-65, get_lepe :67-80, img2windows :184-191, windows2img :194-202) and the `torch.cat` of :174.
// The (B, L, 3C) qkv tensor — the largest activation of a block after the MLP hidden — never exists in global memory, and the
// attention problems start from q / k / v tiles that are already in shared memory.
//
// Work unit (CTA) = (image b, branch, head group, window tile):
//   * window tile: 128 token rows = 128 TMEM lanes holding ONE stripe window (64 < N <= 128 tokens) or TWO windows (N <= 64,
//     rows 0..63 / 64..127), gathered by TMA straight from the (B, H, W, C) activation as 4-D boxes (64 ch, W_sp, H_sp, 1) in
//     WINDOW token order — img2windows is the tensor map;
//   * head group: HG = min(2, heads of the branch) heads, i.e. NQ = 32 HG columns each of q, k and v.
// Phase 1 (GEMM):   acc[128 x 3 NQ] (TMEM, fp32) = X_tile[128 x C] . W'[rows of q|k|v of these heads, C]^T, K in 64-wide blocks
//                   through a TMA ring (128-byte swizzle), single-thread tcgen05.mma, exactly like gemm_tc.cu.
// Phase 2 (epilogue): thread = (row, half of the columns): folded LayerNorm  rstd * (acc - mean * colsum) + bias'  (row statistics
//                   from the producer's side channel, see cswin_linear_args_t), bf16, written as the Q / K / V tiles of each head
//                   in the 64-byte-swizzled UMMA layout the attention MMAs consume (padding rows are written as zeros).
// Phase 3 (attention, per head): S = Q K^T into TMEM, fp32 softmax from TMEM (scale folded into ex2), bf16 P back into TMEM,
//                   O = P V (A from TMEM, V MN-major from smem), LePE 3x3 depthwise conv of V from the same smem tile with
//                   WINDOW-local zero padding, 32-byte stores into the (B, L, C) concat layout — the code of attention_tc.cu,
//                   with the S MMAs of both heads issued up front and head 1's softmax overlapping head 0's P.V.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace cswin {
namespace {

using namespace tc;

constexpr int kThreads = 256;
constexpr int kTileRows = 128;
constexpr int kXStage = kTileRows * 128;     // one 64-channel K block of the token tile: 128 rows x 128 B (SW128)
constexpr int kHeadTile = kTileRows * 64;    // Q / K / V tile of one head: 128 rows x 64 B (32 bf16 channels, SW64)
constexpr int kMaxStages = 4;
constexpr int kMaxCols = 192;                // 3 * NQ <= 192
enum { kFull = 0, kEmpty = kMaxStages, kAcc = 2 * kMaxStages, kS = kAcc + 1, kO = kS + 2, kNumBars = kO + 2 };

struct QaBranch {
  const __nv_bfloat16* cw; const __nv_bfloat16* cb;
  int heads, hs, ws, nww, nwin, N, slots, wtiles, nhg, cta_begin, ch0;
};
struct alignas(64) QaParams {
  CUtensorMap map_x[2];
  CUtensorMap map_w;
  QaBranch br[2];
  const float* bias; const float* cs; const float* ln_stats; int ln_parts; float ln_invC, ln_eps;
  __nv_bfloat16* out; int64_t o_bs, o_ts;
  int nb, reso, C, nkb, NQ, HG, stages, tmem_cols, region;
  float scale, scale_log2e;
  unsigned long long* trace;
};

__device__ __forceinline__ uint32_t sw64_chunk_addr(uint32_t base, int row, int chunk) {
  return base + row * 64 + (((chunk ^ (row >> 1)) & 3) << 4);              // Swizzle<2,4,3> (64-byte swizzle)
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kThreads, 2) qkv_lepe_attn_tc_kernel(const __grid_constant__ QaParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int NQ = P.NQ, HG = P.HG, S = P.stages, NW = 3 * NQ;
  const uint32_t w_bytes = (uint32_t)NW * 128u, stage_bytes = kXStage + w_bytes;
  uint8_t* Ring = smem;                                  // [S][ X 16 KB | W NW x 128 B ]
  const uint32_t tiles_u32 = smem_u32(smem);             // [HG][q, k, v][8 KB]: aliases the ring once the GEMM is complete
  float* sBias = reinterpret_cast<float*>(smem + P.region);
  float* sCs = sBias + kMaxCols;
  float* Wt = sCs + kMaxCols;                            // [2][9][32]
  float* Bc = Wt + 2 * 9 * 32;                           // [2][32]
  float* Xch = Bc + 2 * 32;                              // [2 heads][2 halves][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Xch + 2 * 2 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & 127, half = tid >> 7;
  if (tid == 0) trace_stamp(P.trace, 0);
  const int bi = (P.nb > 1 && (int)blockIdx.x >= P.br[1].cta_begin) ? 1 : 0;
  const QaBranch& br = P.br[bi];
  int local = (int)blockIdx.x - br.cta_begin;
  const int hg = local % br.nhg; local /= br.nhg;
  const int wt = local % br.wtiles;
  const int b = local / br.wtiles;
  const int N = br.N, hs = br.hs, ws = br.ws, slots = br.slots;
  const int slot_rows = kTileRows / slots;
  const int win0 = wt * slots;
  const int np = min(slots, br.nwin - win0);             // windows in this tile
  const int kext = slots == 2 ? 128 : ((N + 15) & ~15);  // key extent of the attention MMAs
  const int slot = row / slot_rows;                      // warp-uniform
  const int n = row - slot * slot_rows;                  // token index inside the window
  const bool valid = slot < np && n < N;
  const int win = win0 + min(slot, np - 1);
  const int ih = win / br.nww, iw = win - ih * br.nww;
  const int r = n / ws, c = n - r * ws;
  const int64_t tok = (int64_t)(ih * hs + r) * P.reso + (iw * ws + c);
  const int head0 = hg * HG;                             // first head (inside the branch) of this CTA
  const int wrow0 = br.ch0 + head0 * 32;                 // first q row of W' (k rows: + C, v rows: + 2 C)

  auto bar = [&](int i) { return smem_u32(&bars[i]); };
  const uint32_t stage_tx = (uint32_t)(np * N * 128) + w_bytes;
  auto load_w = [&](int kb, int s) {
    const uint32_t dst = smem_u32(Ring + (size_t)s * stage_bytes + kXStage);
#pragma unroll
    for (int which = 0; which < 3; ++which)
      tma_load_2d(dst + which * NQ * 128, &P.map_w, bar(kFull + s), kb * 64, which * P.C + wrow0);
  };

  const int npre = P.nkb < S ? P.nkb : S;
  if (warp == 0 && elect_one()) {
    for (int s = 0; s < S; ++s) { mbar_init(bar(kFull + s), 1); mbar_init(bar(kEmpty + s), 1); }
    mbar_init(bar(kAcc), 1);
    for (int h = 0; h < 2; ++h) { mbar_init(bar(kS + h), 1); mbar_init(bar(kO + h), 1); }
    fence_barrier_init();
    fence_proxy_async();
    tma_prefetch_desc(&P.map_w);
    for (int kb = 0; kb < npre; ++kb) {                  // weights do not depend on the previous kernel: before the PDL wait
      mbar_expect_tx(bar(kFull + kb), stage_tx);
      load_w(kb, kb);
    }
    tma_prefetch_desc(&P.map_x[bi]);
  }
  if (warp == 1) { tmem_alloc(smem_u32(tmem_slot), (uint32_t)P.tmem_cols); tmem_relinquish(); }
  for (int j = tid; j < NW; j += kThreads) {             // per-column constants of this CTA's q | k | v columns
    const int which = j / NQ;
    const int g = which * P.C + wrow0 + (j - which * NQ);
    sBias[j] = P.bias != nullptr ? P.bias[g] : 0.f;
    sCs[j] = P.cs != nullptr ? P.cs[g] : 0.f;
  }
  if (tid < HG * 36) {                                   // LePE weights: 288 contiguous bf16 per head -> Wt[hh][tap][ch] fp32
    const int hh = tid / 36, i = tid - hh * 36;
    const uint4 raw = *reinterpret_cast<const uint4*>(br.cw + (size_t)(head0 + hh) * 288 + i * 8);
    const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int idx = i * 8 + e;                         // = ch * 9 + tap
      const int ch = idx / 9, t = idx - ch * 9;
      Wt[(hh * 9 + t) * 32 + ch] = (e & 1) ? bf16_hi(w4[e >> 1]) : bf16_lo(w4[e >> 1]);
    }
  } else if (tid >= 128 && tid < 128 + HG * 32) {
    const int hh = (tid - 128) >> 5, ch = tid & 31;
    Bc[hh * 32 + ch] = __bfloat162float(br.cb[(head0 + hh) * 32 + ch]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // the dependent kernel may start its prologue only now: a dependent CTA that allocated TMEM first would wait for this
  // grid (griddepcontrol.wait) while this CTA waits for TMEM
  pdl_trigger();
  if (tid == 0) trace_stamp(P.trace, 1);
  pdl_wait();                                            // x, ln_stats (previous kernel's output) and `out` are safe from here

  // ---------------- phase 1: the qkv GEMM of this tile ----------------
  if (warp == 0) {
    if (elect_one()) {                                   // TMA producer
      int s = 0; uint32_t ph = 1;
      for (int kb = 0; kb < P.nkb; ++kb) {
        if (kb >= S) {
          mbar_wait(bar(kEmpty + s), ph);
          mbar_expect_tx(bar(kFull + s), stage_tx);
          load_w(kb, s);
        }
        const uint32_t dst = smem_u32(Ring + (size_t)s * stage_bytes);
        for (int sl = 0; sl < np; ++sl) {
          const int w = win0 + sl, wih = w / br.nww, wiw = w - wih * br.nww;
          tma_load_4d(dst + sl * slot_rows * 128, &P.map_x[bi], bar(kFull + s), kb * 64, wiw * ws, wih * hs, b);
        }
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {                                   // MMA issuer
      const uint32_t idesc = make_idesc_bf16(128, NW, 0, 0);
      int s = 0; uint32_t ph = 0;
      for (int kb = 0; kb < P.nkb; ++kb) {
        mbar_wait(bar(kFull + s), ph);
        tc_fence_after();
        const uint32_t xa = smem_u32(Ring + (size_t)s * stage_bytes);
        const uint64_t ad = make_smem_desc(xa, 16, 1024, kLayoutSw128);
        const uint64_t wd = make_smem_desc(xa + kXStage, 16, 1024, kLayoutSw128);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_ss(tmem_base, ad + 2 * k, wd + 2 * k, idesc, (kb | k) != 0);
        tc_commit(bar(kEmpty + s));
        if (++s == S) { s = 0; ph ^= 1; }
      }
      tc_commit(bar(kAcc));
    }
  }
  // folded LayerNorm: statistics of my token's row from the producer's partial sums (overlaps the GEMM)
  float ln_mean = 0.f, ln_rstd = 1.f;
  if (P.ln_stats != nullptr && valid) {
    const int64_t m = (int64_t)b * P.reso * P.reso + tok;
    float s1 = 0.f, s2 = 0.f;
    for (int p = 0; p < P.ln_parts; ++p) { s1 += P.ln_stats[(m * P.ln_parts + p) * 2]; s2 += P.ln_stats[(m * P.ln_parts + p) * 2 + 1]; }
    ln_mean = s1 * P.ln_invC;
    ln_rstd = rsqrtf(fmaxf(fmaf(-ln_mean, ln_mean, s2 * P.ln_invC), 0.f) + P.ln_eps);
  }
  __syncwarp();
  mbar_wait(bar(kAcc), 0);
  if (tid == 0) trace_stamp(P.trace, 2);                 // accumulator ready
  tc_fence_after();

  // ---------------- phase 2: accumulator -> bf16 Q / K / V tiles in shared memory ----------------
  const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  {
    const int ncol_h = NW >> 1;                          // 48 or 96 columns per thread, in 16-column pieces
    const float nmean = -ln_mean;
    for (int ch = 0; ch < ncol_h / 16; ++ch) {
      const int col0 = half * ncol_h + ch * 16;
      uint32_t v[16];
      tmem_ld16(trow + col0, v);
      tmem_wait_ld();
      const int which = col0 / NQ, rem = col0 - which * NQ;
      const int hh = rem >> 5, c16 = (rem >> 4) & 1;
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float y0 = fmaf(ln_rstd, fmaf(nmean, sCs[col0 + 2 * j], __uint_as_float(v[2 * j])), sBias[col0 + 2 * j]);
        const float y1 = fmaf(ln_rstd, fmaf(nmean, sCs[col0 + 2 * j + 1], __uint_as_float(v[2 * j + 1])), sBias[col0 + 2 * j + 1]);
        pk[j] = valid ? pack_bf16x2(y0, y1) : 0u;        // padding rows: zeros (0 * stale-NaN would poison the MMAs)
      }
      const uint32_t base = tiles_u32 + (uint32_t)((hh * 3 + which) * kHeadTile);
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sw64_chunk_addr(base, row, 2 * c16)), "r"(pk[0]), "r"(pk[1]),
                   "r"(pk[2]), "r"(pk[3]) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sw64_chunk_addr(base, row, 2 * c16 + 1)), "r"(pk[4]), "r"(pk[5]),
                   "r"(pk[6]), "r"(pk[7]) : "memory");
    }
  }
  fence_proxy_async();                                   // generic-proxy tile writes -> visible to the tensor core
  tc_fence_before();
  __syncthreads();
  if (tid == 0) trace_stamp(P.trace, 3);                 // q, k, v tiles published

  // ---------------- phase 3: attention ----------------
  if (warp == 0 && elect_one()) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, kext, 0, 0);
    for (int hh = 0; hh < HG; ++hh) {
      const uint64_t qd = make_smem_desc(tiles_u32 + (hh * 3 + 0) * kHeadTile, 16, 8 * 64, kLayoutSw64);
      const uint64_t kd = make_smem_desc(tiles_u32 + (hh * 3 + 1) * kHeadTile, 16, 8 * 64, kLayoutSw64);
      mma_ss(tmem_base + hh * 128, qd, kd, idesc, false);
      mma_ss(tmem_base + hh * 128, qd + 2, kd + 2, idesc, true);
      tc_commit(bar(kS + hh));
    }
  }
  const int hcols = slot_rows >> 1;                      // my share of the row's keys: 32 (two windows) or 64 (one window)
  const int kbeg = half * hcols;                         // first slot-local key of my half
  const int cbeg = slot * slot_rows + kbeg;              // first S column of my half
  const int nch = hcols >> 5;                            // 32-column chunks: 1 or 2
  float inv_sum[2] = {1.f, 1.f};
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    if (hh < HG) {
      float* Xmax = Xch + hh * 256;
      mbar_wait(bar(kS + hh), 0);
      tc_fence_after();
      const uint32_t srow = trow + hh * 128;
      float mx = -INFINITY;
      for (int cc = 0; cc < nch; ++cc) {
        if (kbeg + 32 * cc >= kext) break;
        uint32_t v[32];
        tmem_ld32(srow + cbeg + 32 * cc, v);
        tmem_wait_ld();
        const int lim = N - (kbeg + 32 * cc);
        if (lim >= 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j < lim) mx = fmaxf(mx, __uint_as_float(v[j]));
        }
      }
      Xmax[half * 128 + row] = mx;
      __syncthreads();
      mx = fmaxf(mx, Xmax[(half ^ 1) * 128 + row]);
      const float mxs = mx * P.scale_log2e;
      float sum = 0.f;
      float2 sum2 = make_float2(0.f, 0.f);
      const float2 sl2 = make_float2(P.scale_log2e, P.scale_log2e), nmxs2 = make_float2(-mxs, -mxs);
      uint32_t pk[2][16];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        if (cc < nch && kbeg + 32 * cc < kext) {
          uint32_t v[32];
          tmem_ld32(srow + cbeg + 32 * cc, v);
          tmem_wait_ld();
          const int lim = N - (kbeg + 32 * cc);
          if (lim >= 32) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float2 t = ffma2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sl2, nmxs2);
              const float2 e = make_float2(ex2_approx(t.x), ex2_approx(t.y));
              sum2 = fadd2(sum2, e);
              pk[cc][j >> 1] = pack_bf16x2(e.x, e.y);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float e0 = (j < lim) ? ex2_approx(fmaf(__uint_as_float(v[j]), P.scale_log2e, -mxs)) : 0.f;
              const float e1 = (j + 1 < lim) ? ex2_approx(fmaf(__uint_as_float(v[j + 1]), P.scale_log2e, -mxs)) : 0.f;
              sum += e0 + e1;
              pk[cc][j >> 1] = pack_bf16x2(e0, e1);
            }
          }
        }
      }
      // P (bf16, columns [0,64) of this head's region) aliases S columns the partner thread of this row may still be reading
      __syncthreads();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc)
        if (cc < nch && kbeg + 32 * cc < kext) tmem_st16(srow + ((cbeg + 32 * cc) >> 1), pk[cc]);
      if (slots == 2) {                                  // keys of the other window: P = 0
        uint32_t z[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) z[j] = 0u;
        tmem_st16(srow + ((1 - slot) * 32) + half * 16, z);
      }
      sum += sum2.x + sum2.y;
      Xmax[half * 128 + row] = sum;                      // (the max exchange is complete: reuse as the sum exchange)
      tmem_wait_st();
      tc_fence_before();
      __syncthreads();
      if (warp == 0 && elect_one()) {
        tc_fence_after();
        const uint64_t vd = make_smem_desc(tiles_u32 + (hh * 3 + 2) * kHeadTile, 8 * 64, 8 * 64, kLayoutSw64);
        const uint32_t idesc = make_idesc_bf16(128, 32, 0, 1);       // B = V is MN-major
        for (int k = 0; k < kext / 16; ++k)
          mma_ts(tmem_base + hh * 128 + 64, tmem_base + hh * 128 + 8 * k, vd + (uint64_t)k * ((16 * 64) >> 4), idesc, k > 0);
        tc_commit(bar(kO + hh));
      }
      sum += Xmax[(half ^ 1) * 128 + row];
      inv_sum[hh] = 1.0f / sum;
    }
  }
  if (tid == 0) trace_stamp(P.trace, 4);                 // both P.V issued

  // ---- LePE for my token, channels [16 half, 16 half + 16) of each head, then O / rowsum + LePE -> out ----
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    if (hh < HG) {
      float2 lp[8];
      {
        const float2* bc = reinterpret_cast<const float2*>(Bc + hh * 32 + half * 16);
#pragma unroll
        for (int j = 0; j < 8; ++j) lp[j] = bc[j];
      }
      if (valid) {
        const uint32_t vbase = tiles_u32 + (hh * 3 + 2) * kHeadTile;
        const float* wt_ = Wt + hh * 9 * 32 + half * 16;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int rr = r + t / 3 - 1, cc = c + t % 3 - 1;
          if (rr >= 0 && rr < hs && cc >= 0 && cc < ws) {
            const int vr = slot * slot_rows + rr * ws + cc;
#pragma unroll
            for (int q2 = 0; q2 < 2; ++q2) {
              uint4 vv;
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(vv.x), "=r"(vv.y), "=r"(vv.z), "=r"(vv.w)
                           : "r"(sw64_chunk_addr(vbase, vr, half * 2 + q2)));
              const float4 w0 = *reinterpret_cast<const float4*>(wt_ + t * 32 + q2 * 8);
              const float4 w1 = *reinterpret_cast<const float4*>(wt_ + t * 32 + q2 * 8 + 4);
              lp[q2 * 4 + 0] = ffma2(make_float2(w0.x, w0.y), make_float2(bf16_lo(vv.x), bf16_hi(vv.x)), lp[q2 * 4 + 0]);
              lp[q2 * 4 + 1] = ffma2(make_float2(w0.z, w0.w), make_float2(bf16_lo(vv.y), bf16_hi(vv.y)), lp[q2 * 4 + 1]);
              lp[q2 * 4 + 2] = ffma2(make_float2(w1.x, w1.y), make_float2(bf16_lo(vv.z), bf16_hi(vv.z)), lp[q2 * 4 + 2]);
              lp[q2 * 4 + 3] = ffma2(make_float2(w1.z, w1.w), make_float2(bf16_lo(vv.w), bf16_hi(vv.w)), lp[q2 * 4 + 3]);
            }
          }
        }
      }
      mbar_wait(bar(kO + hh), 0);
      tc_fence_after();
      uint32_t o[16];
      tmem_ld16(trow + hh * 128 + 64 + half * 16, o);
      tmem_wait_ld();
      if (valid) {
        const float inv = inv_sum[hh];
        __nv_bfloat16* dst = P.out + (int64_t)b * P.o_bs + tok * P.o_ts + br.ch0 + (head0 + hh) * 32 + half * 16;
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 y = ffma2(make_float2(__uint_as_float(o[2 * j]), __uint_as_float(o[2 * j + 1])), make_float2(inv, inv), lp[j]);
          w[j] = pack_bf16x2(y.x, y.y);
        }
        *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(dst + 8) = make_uint4(w[4], w[5], w[6], w[7]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) trace_stamp(P.trace, 7);
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// (branch geometry) -> is it inside the kernel's envelope?
bool branch_ok(int reso, int heads, int hs, int ws) {
  if (heads <= 0 || hs <= 0 || ws <= 0 || reso % hs || reso % ws) return false;
  if (hs * ws > 128 || hs > 256 || ws > 256) return false;
  const int hg = heads < 2 ? heads : 2;
  return heads % hg == 0;
}

}  // namespace

int qkv_attn_supported(int C, int reso, int nb, const int* heads, const int* hs, const int* ws) {
  if (C % 64 || C < 64 || C > 64 * kMaxStages || nb < 1 || nb > 2 || reso <= 0) return 0;
  int ctot = 0;
  for (int i = 0; i < nb; ++i) {
    if (!branch_ok(reso, heads[i], hs[i], ws[i])) return 0;
    if (i > 0 && (heads[i] < 2) != (heads[0] < 2)) return 0;        // one head-group width per launch
    ctot += heads[i] * 32;
  }
  if (ctot != C) return 0;                                            // head_dim 32, branches tile the channels
  return tc::encode_tiled_fn() != nullptr ? 1 : 0;
}

int qkv_attn_fwd_tc(const cswin_qkv_attn_args_t* a, cudaStream_t stream) {
  int heads[2], hs[2], ws[2];
  for (int i = 0; i < 2; ++i) { heads[i] = a->br[i].heads; hs[i] = a->br[i].H_sp; ws[i] = a->br[i].W_sp; }
  CSWIN_REQUIRE(qkv_attn_supported(a->C, a->reso, a->n_branches, heads, hs, ws), CSWIN_ERR_UNSUPPORTED,
                "qkv_lepe_attention_fwd: outside the fused kernel's envelope (C in {64,128,192,256}, head_dim 32, windows <= 128 tokens); "
                "use cswin_linear_fwd + cswin_lepe_attention_fwd");
  CSWIN_REQUIRE(aligned16(a->x) && aligned16(a->w) && aligned16(a->out) && (a->x_ts * 2) % 16 == 0 && (a->x_bs * 2) % 16 == 0 &&
                (a->ldw * 2) % 16 == 0 && (a->o_ts * 2) % 16 == 0 && (a->o_bs * 2) % 16 == 0 && a->x_ts > 0 && a->o_ts > 0,
                CSWIN_ERR_UNSUPPORTED, "qkv_lepe_attention_fwd: operands are not TMA-compatible (16-byte aligned pointers / pitches)");
  CSWIN_REQUIRE((a->ln_stats == nullptr) == (a->ln_colsum == nullptr), CSWIN_ERR_INVALID,
                "qkv_lepe_attention_fwd: ln_stats and ln_colsum go together");
  if (a->B == 0) return CSWIN_OK;

  QaParams P;
  memset(&P, 0, sizeof(P));
  P.nb = a->n_branches; P.reso = a->reso; P.C = a->C; P.nkb = a->C / 64;
  P.HG = heads[0] < 2 ? 1 : 2;
  P.NQ = 32 * P.HG;
  P.tmem_cols = P.HG == 1 ? 128 : 256;
  P.bias = a->bias_f32; P.cs = a->ln_colsum; P.ln_stats = a->ln_stats; P.ln_parts = a->ln_stats_parts;
  P.ln_invC = 1.0f / (float)a->C; P.ln_eps = a->ln_eps;
  P.out = (__nv_bfloat16*)a->out; P.o_bs = a->o_bs; P.o_ts = a->o_ts;
  P.scale = a->scale; P.scale_log2e = a->scale * 1.4426950408889634f;
  P.trace = g_trace.load(std::memory_order_relaxed);
  const int stage_bytes = kXStage + 3 * P.NQ * 128;
  static const int forced_stages = [] { const char* e = getenv("CSWIN_QA_STAGES"); return e ? atoi(e) : 0; }();
  P.stages = P.nkb;
  if ((size_t)P.stages * stage_bytes > 100 * 1024) P.stages = 2;      // two CTAs per SM (C = 256: 2 x 40 KB instead of 160 KB)
  if (forced_stages >= 1 && forced_stages <= kMaxStages) P.stages = forced_stages < P.nkb ? forced_stages : P.nkb;
  const int tiles_bytes = P.HG * 3 * kHeadTile;
  P.region = P.stages * stage_bytes > tiles_bytes ? P.stages * stage_bytes : tiles_bytes;

  int ctas = 0, ch0 = 0;
  for (int i = 0; i < a->n_branches; ++i) {
    QaBranch& d = P.br[i];
    d.cw = (const __nv_bfloat16*)a->br[i].conv_w; d.cb = (const __nv_bfloat16*)a->br[i].conv_b;
    CSWIN_REQUIRE(d.cw && d.cb && aligned16(d.cw), CSWIN_ERR_INVALID, "qkv_lepe_attention_fwd: conv_w / conv_b missing or unaligned");
    d.heads = heads[i]; d.hs = hs[i]; d.ws = ws[i]; d.nww = a->reso / ws[i];
    d.nwin = (a->reso / hs[i]) * (a->reso / ws[i]); d.N = hs[i] * ws[i];
    d.slots = d.N <= 64 ? 2 : 1;
    d.wtiles = (d.nwin + d.slots - 1) / d.slots;
    d.nhg = d.heads / P.HG;
    d.cta_begin = ctas; d.ch0 = ch0;
    ctas += a->B * d.wtiles * d.nhg;
    ch0 += d.heads * 32;
    const uint64_t dims[4] = {(uint64_t)a->C, (uint64_t)a->reso, (uint64_t)a->reso, (uint64_t)a->B};
    const uint64_t str[3] = {(uint64_t)a->x_ts * 2, (uint64_t)a->x_ts * 2 * a->reso, (uint64_t)a->x_bs * 2};
    const uint32_t box[4] = {64, (uint32_t)d.ws, (uint32_t)d.hs, 1};
    if (!tc::make_tensor_map_bf16(&P.map_x[i], a->x, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return CSWIN_ERR_CUDA;
  }
  if (a->n_branches == 1) { P.br[1] = P.br[0]; P.map_x[1] = P.map_x[0]; }
  {
    const uint64_t dims[2] = {(uint64_t)a->C, (uint64_t)(3 * a->C)}, str[1] = {(uint64_t)a->ldw * 2};
    const uint32_t box[2] = {64, (uint32_t)P.NQ};
    if (!tc::make_tensor_map_bf16(&P.map_w, a->w, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return CSWIN_ERR_CUDA;
  }
  const size_t smem = 1024 + (size_t)P.region + (2 * kMaxCols + 2 * 9 * 32 + 2 * 32 + 2 * 2 * 128) * 4 + kNumBars * 8 + 64;
  static std::atomic<int> configured{0};
  if (!configured.load(std::memory_order_acquire)) {
    CSWIN_CUDA_OK(cudaFuncSetAttribute(qkv_lepe_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured.store(1, std::memory_order_release);
  }
  CSWIN_CUDA_OK(launch_pdl(qkv_lepe_attn_tc_kernel, dim3((unsigned)ctas), dim3(kThreads), smem, stream, P));
  CSWIN_LAUNCH_CHECK();
  g_tc_launches.fetch_add(1, std::memory_order_relaxed);
  return CSWIN_OK;
}

}  // namespace cswin
