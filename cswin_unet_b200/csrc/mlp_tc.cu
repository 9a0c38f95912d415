// mlp_tc.cu — the whole MLP half of a CSWinBlock in ONE tcgen05 kernel (sm_100a, bf16):
//
//   out = x + GELU( LayerNorm(x) W1^T + b1 ) W2^T + b2                       (cswin_unet.py:179, Mlp :22-26)
//
// The (M, 4C) hidden activation never exists in global memory.  Per CTA: one 128-row tile of x (resident in shared
// memory for the whole kernel) and a contiguous range of hidden units, walked in 64-wide sub-chunks j:
//
//   MMA1(j): acc1[j&1] (TMEM, 128 x 64 fp32)  = X (128 x C)  .  W1[h_j : h_j+64, :]^T               K = C
//   EPI(j) : H[j&1]   (smem, 128 x 64 bf16)  = GELU( rstd * (acc1 - mean * colsum) + b1' )          folded LayerNorm
//   MMA2(j): acc2     (TMEM, 128 x C fp32)  += H[j&1] (128 x 64) . W2[:, h_j : h_j+64]^T            K = 64
//
// software-pipelined so that MMA1(j+1) runs on the tensor core while the 8 epilogue warps do EPI(j).  LayerNorm is applied
// algebraically (see cswin_linear_args_t): W1 arrives as W1 o gamma, the row statistics come from the (sum, sum^2) side
// channel the producing Linear wrote, so the GEMM consumes the raw rows.  W1 / W2 sub-chunks (128*C bytes each, the same
// size) stream through one TMA ring in consumption order W1_0, W1_1, W2_0, W1_2, W2_1, ...
//
// Hidden split: for C = 256 the 37 row tiles of the 224^2 / batch-24 workload would leave 3/4 of the GPU idle, so the hidden
// dimension is split across a thread-block cluster of 2 (2 x 37 = 74 CTAs; clusters of 4 cannot all be co-resident on the
// 148-SM part).  Each CTA parks its partial acc2 in shared memory as fp32, the pair exchanges through distributed shared
// memory, and each CTA finishes (bias, residual, bf16 store, row statistics for the next block's folded LayerNorm) half of
// the columns.  Without a split the same code path runs with a cluster of 1.
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread tcgen05.mma issuer, warps 2..9 = epilogue.
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace cswin {
namespace {

using namespace tc;

// -DCSWIN_MLP_PROFILE: per-role wait-cycle accounting into the debug trace slots 8..15 (tools/trace_kernel.py mlp)
#ifdef CSWIN_MLP_PROFILE
#define PWAIT(acc, stmt) do { const long long c0_ = clock64(); stmt; acc += clock64() - c0_; } while (0)
#else
#define PWAIT(acc, stmt) do { stmt; } while (0)
#endif

constexpr int BM = 128, HC = 64;
constexpr int kThreads = 320;
constexpr int kMaxSlots = 8;
constexpr int kMaxHidPerCta = 512;

struct alignas(64) MlpTcParams {
  CUtensorMap map_x, map_w1, map_w2;
  const float* b1; const float* cs1; const float* b2;
  const float* ln_stats; int ln_parts; float ln_invC, ln_eps;
  const __nv_bfloat16* x; int64_t ldx;
  __nv_bfloat16* out; int64_t ldo;
  float* stats_out;
  int64_t M;
  int C, hid_per_cta, nsub, slots, spl, tmem_cols;
  unsigned long long* trace;
};

// barrier indices
enum { kFull = 0, kEmpty = kMaxSlots, kXFull = 2 * kMaxSlots, kAcc1Full, kAcc1Free = kAcc1Full + 2, kHFull = kAcc1Free + 2,
       kHFree = kHFull + 2, kAcc2Full = kHFree + 2, kNumBars };

// t-th weight chunk in consumption order: W1_0, W1_1, W2_0, W1_2, W2_1, ..., W1_{n-1}, W2_{n-2}, W2_{n-1}
__device__ __forceinline__ void chunk_of(int t, int nsub, bool& is_w2, int& j) {
  if (t == 0) { is_w2 = false; j = 0; }
  else if (t == 2 * nsub - 1) { is_w2 = true; j = nsub - 1; }
  else if (t & 1) { is_w2 = false; j = (t + 1) >> 1; }
  else { is_w2 = true; j = (t >> 1) - 1; }
}

__global__ void __launch_bounds__(kThreads, 2) mlp_tc_kernel(const __grid_constant__ MlpTcParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int C = P.C, KB1 = C >> 6, NS = P.slots, nsub = P.nsub;
  const uint32_t x_bytes = (uint32_t)BM * C * 2, slot_bytes = 128u * C;
  uint8_t* Xs = smem;                                        // [C/64][128][64] bf16, SW128
  uint8_t* Ring = Xs + x_bytes;                              // [NS][128*C bytes]
  uint8_t* Hs = Ring + (size_t)NS * slot_bytes;              // [2][128][64] bf16, SW128 K-major (A operand of MMA2)
  float* sB1 = reinterpret_cast<float*>(Hs + 2 * 16384);     // [hid_per_cta]
  float* sCs = sB1 + kMaxHidPerCta;                          // [hid_per_cta]
  float* sB2 = sCs + kMaxHidPerCta;                          // [C]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB2 + 256);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
  float* Stage = reinterpret_cast<float*>(smem);             // [128][C] fp32 partial acc2, aliases X + ring once all MMAs are done

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  if (tid == 0) trace_stamp(P.trace, 0);
  const uint32_t rank = P.spl > 1 ? cluster_ctarank() : 0u;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int h0 = (int)rank * P.hid_per_cta;                  // first hidden unit of this CTA

  auto bar = [&](int i) { return smem_u32(&bars[i]); };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(bar(kFull + s), 1); mbar_init(bar(kEmpty + s), 1); }
    mbar_init(bar(kXFull), 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(kAcc1Full + b), 1); mbar_init(bar(kAcc1Free + b), 8);
      mbar_init(bar(kHFull + b), 8);    mbar_init(bar(kHFree + b), 1);
    }
    mbar_init(bar(kAcc2Full), 1);
    fence_barrier_init();
    tma_prefetch_desc(&P.map_x); tma_prefetch_desc(&P.map_w1); tma_prefetch_desc(&P.map_w2);
  }
  if (warp == 1) { tmem_alloc(smem_u32(tmem_slot), (uint32_t)P.tmem_cols); tmem_relinquish(); }
  if (warp >= 2) {                                           // per-column constants (weights: no dependence on the producer kernel)
    for (int j = tid - 64; j < P.hid_per_cta; j += 256) { sB1[j] = P.b1[h0 + j]; sCs[j] = P.cs1[h0 + j]; }
    for (int j = tid - 64; j < C; j += 256) sB2[j] = P.b2[j];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) trace_stamp(P.trace, 1);

  const int nchunks = 2 * nsub;
  auto issue_chunk = [&](int t, int s) {                     // t-th weight sub-chunk -> ring slot s (= t % NS)
    bool is_w2; int j;
    chunk_of(t, nsub, is_w2, j);
    mbar_expect_tx(bar(kFull + s), slot_bytes);
    const uint32_t dst = smem_u32(Ring + (size_t)s * slot_bytes);
    if (!is_w2) for (int kb = 0; kb < KB1; ++kb) tma_load_2d(dst + kb * 8192, &P.map_w1, bar(kFull + s), kb * 64, h0 + j * HC);
    else tma_load_2d(dst, &P.map_w2, bar(kFull + s), h0 + j * HC, 0);
  };
  const int npre = nchunks < NS ? nchunks : NS;
  if (warp == 0 && elect_one()) for (int t = 0; t < npre; ++t) issue_chunk(t, t);     // weights: before the PDL wait
  pdl_wait();                                                // x, ln_stats and out are safe from here

  if (warp == 0) {
    if (elect_one()) {                                       // ---- TMA producer ----
      mbar_expect_tx(bar(kXFull), x_bytes);
      for (int kb = 0; kb < KB1; ++kb) tma_load_2d(smem_u32(Xs + (size_t)kb * 16384), &P.map_x, bar(kXFull), kb * 64, (int)m0);
      int ps = 0; uint32_t pph = 0;                          // npre == NS whenever this loop runs: slot 0, first release
      for (int t = npre; t < nchunks; ++t) {
        mbar_wait(bar(kEmpty + ps), pph);
        issue_chunk(t, ps);
        if (++ps == NS) { ps = 0; pph ^= 1; }
      }
      trace_stamp(P.trace, 2);
    }
  } else if (warp == 1) {
    if (elect_one()) {                                       // ---- MMA issuer (elect.sync keeps the warp-uniform datapath) ----
      const uint32_t idesc1 = make_idesc_bf16(BM, HC, 0, 0), idesc2 = make_idesc_bf16(BM, C, 0, 0);
      const uint32_t acc2 = tmem_base + 128;
      int cs = 0; uint32_t cph = 0;                          // ring slot / parity of the next chunk to consume
#ifdef CSWIN_MLP_PROFILE
      long long w_ring = 0, w_free = 0, w_h = 0, c_start = clock64(), i_m1 = 0, i_m2 = 0, i_cm = 0;
#endif
      auto mma1 = [&](int j) {
        const int s = cs;
        PWAIT(w_ring, mbar_wait(bar(kFull + s), cph));
        tc_fence_after();
        const uint32_t wbase = smem_u32(Ring + (size_t)s * slot_bytes);
        PWAIT(i_m1, for (int kb = 0; kb < KB1; ++kb) {
          const uint64_t ad = make_smem_desc(smem_u32(Xs + (size_t)kb * 16384), 16, 1024, kLayoutSw128);
          const uint64_t wd = make_smem_desc(wbase + kb * 8192, 16, 1024, kLayoutSw128);
          _Pragma("unroll")
          for (int k = 0; k < 4; ++k) mma_ss(tmem_base + (j & 1) * HC, ad + 2 * k, wd + 2 * k, idesc1, (kb | k) != 0);
        });
        PWAIT(i_cm, tc_commit(bar(kEmpty + s)); tc_commit(bar(kAcc1Full + (j & 1))));
        if (++cs == NS) { cs = 0; cph ^= 1; }
      };
      auto mma2 = [&](int j) {
        const int s = cs;
        PWAIT(w_ring, mbar_wait(bar(kFull + s), cph));
        PWAIT(w_h, mbar_wait(bar(kHFull + (j & 1)), (j >> 1) & 1));
        tc_fence_after();
        const uint64_t ad = make_smem_desc(smem_u32(Hs + (j & 1) * 16384), 16, 1024, kLayoutSw128);
        const uint64_t wd = make_smem_desc(smem_u32(Ring + (size_t)s * slot_bytes), 16, 1024, kLayoutSw128);
        PWAIT(i_m2, _Pragma("unroll") for (int k = 0; k < 4; ++k) mma_ss(acc2, ad + 2 * k, wd + 2 * k, idesc2, (j | k) != 0));
        PWAIT(i_cm, tc_commit(bar(kEmpty + s)); tc_commit(bar(kHFree + (j & 1))));
        if (++cs == NS) { cs = 0; cph ^= 1; }
      };
      mbar_wait(bar(kXFull), 0);
      trace_stamp(P.trace, 3);
      mma1(0);
      for (int j = 0; j < nsub; ++j) {
        if (j + 1 < nsub) {
          if (j + 1 >= 2) { PWAIT(w_free, mbar_wait(bar(kAcc1Free + ((j + 1) & 1)), (((j + 1) >> 1) - 1) & 1)); tc_fence_after(); }
          mma1(j + 1);
        }
        mma2(j);
      }
      tc_commit(bar(kAcc2Full));
      trace_stamp(P.trace, 4);
#ifdef CSWIN_MLP_PROFILE
      if (P.trace != nullptr && blockIdx.x + blockIdx.y * gridDim.x < 1024) {
        unsigned long long* tr = P.trace + (size_t)(blockIdx.x + blockIdx.y * gridDim.x) * 16;
        tr[8] = w_ring; tr[9] = w_free; tr[10] = w_h; tr[11] = clock64() - c_start; tr[15] = i_m1; tr[13] = i_m2; tr[14] = i_cm;
      }
#endif
    }
  } else {
    // ---- epilogue warps 2..9: TMEM lane quadrant q = warp % 4; the two warps of a quadrant split the columns ----
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int64_t mrow = m0 + row;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    float mean = 0.f, rstd = 1.f;
    if (mrow < P.M) {
      float s1 = 0.f, s2 = 0.f;
      for (int p = 0; p < P.ln_parts; ++p) { s1 += P.ln_stats[(mrow * P.ln_parts + p) * 2]; s2 += P.ln_stats[(mrow * P.ln_parts + p) * 2 + 1]; }
      mean = s1 * P.ln_invC;
      rstd = rsqrtf(fmaxf(fmaf(-mean, mean, s2 * P.ln_invC), 0.f) + P.ln_eps);
    }
    const float nmean = -mean;
#ifdef CSWIN_MLP_PROFILE
    long long e_acc = 0, e_hfree = 0;
#endif
    for (int j = 0; j < nsub; ++j) {
      const int b = j & 1;
      PWAIT(e_acc, mbar_wait(bar(kAcc1Full + b), (j >> 1) & 1));
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(trow + b * HC + half * 32, v);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kAcc1Free + b));        // acc1[b] may be overwritten by MMA1(j + 2)
      const float4* b4 = reinterpret_cast<const float4*>(sB1 + j * HC + half * 32);
      const float4* c4 = reinterpret_cast<const float4*>(sCs + j * HC + half * 32);
      uint32_t pk[16];
      const float2 nmean2 = make_float2(nmean, nmean), rstd2 = make_float2(rstd, rstd);
#pragma unroll
      for (int g = 0; g < 8; ++g) {                          // packed fp32 pairs: FFMA2 fold + 7-op GELU per two elements
        const float4 bb = b4[g], cc = c4[g];
        const float2 a0 = make_float2(__uint_as_float(v[g * 4 + 0]), __uint_as_float(v[g * 4 + 1]));
        const float2 a1 = make_float2(__uint_as_float(v[g * 4 + 2]), __uint_as_float(v[g * 4 + 3]));
        const float2 f0 = gelu_fast2(ffma2(rstd2, ffma2(nmean2, make_float2(cc.x, cc.y), a0), make_float2(bb.x, bb.y)));
        const float2 f1 = gelu_fast2(ffma2(rstd2, ffma2(nmean2, make_float2(cc.z, cc.w), a1), make_float2(bb.z, bb.w)));
        pk[g * 2] = pack_bf16x2(f0.x, f0.y);
        pk[g * 2 + 1] = pack_bf16x2(f1.x, f1.y);
      }
      if (j >= 2) PWAIT(e_hfree, mbar_wait(bar(kHFree + b), ((j >> 1) - 1) & 1));   // MMA2(j - 2) has finished reading H[b]
      const uint32_t hrow = smem_u32(Hs + b * 16384) + row * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c) {                          // 16-byte chunk (half*4 + c) of the 128-byte row, SW128 XOR
        const uint32_t addr = hrow + ((((half << 2) + c) ^ (row & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pk[c * 4]), "r"(pk[c * 4 + 1]), "r"(pk[c * 4 + 2]),
                     "r"(pk[c * 4 + 3]) : "memory");
      }
      fence_proxy_async();                                   // generic-proxy writes -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kHFull + b));
    }
#ifdef CSWIN_MLP_PROFILE
    if (warp == 2 && lane == 0 && P.trace != nullptr && blockIdx.x + blockIdx.y * gridDim.x < 1024) {
      unsigned long long* tr = P.trace + (size_t)(blockIdx.x + blockIdx.y * gridDim.x) * 16;
      tr[12] = e_acc;
    }
#endif
    // ---- partial acc2 -> fp32 staging (16-byte chunks XOR-swizzled by row so that row-per-lane stores are conflict-free) ----
    mbar_wait(bar(kAcc2Full), 0);
    if (warp == 2 && lane == 0) trace_stamp(P.trace, 5);
    tc_fence_after();
    const int ncol_half = C >> 1;
    const uint32_t srow = smem_u32(Stage) + (uint32_t)row * C * 4;
    for (int u = 0; u < ncol_half; u += 32) {
      const int col0 = half * ncol_half + u;
      uint32_t v[32];
      tmem_ld32(trow + 128 + col0, v);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t addr = srow + ((((col0 >> 2) + c) ^ (row & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v[c * 4]), "r"(v[c * 4 + 1]), "r"(v[c * 4 + 2]),
                     "r"(v[c * 4 + 3]) : "memory");
      }
    }
    tc_fence_before();
  }

  // ---- exchange: every CTA of the cluster has parked its partial sums ----
  if (P.spl > 1) cluster_sync_all(); else __syncthreads();
  if (tid == 0) trace_stamp(P.trace, 6);

  if (warp >= 2) {
    // this CTA finishes columns [rank * C/spl, (rank+1) * C/spl): 8 columns (16 B of bf16) per item, G lanes per row
    const int my_cols = C / P.spl, cb = (int)rank * my_cols, G = my_cols >> 3;
    const int e = tid - 64;
    const int g = e % G;
    const int col = cb + g * 8;
    const float4 bA = *reinterpret_cast<const float4*>(sB2 + col), bB = *reinterpret_cast<const float4*>(sB2 + col + 4);
    const uint32_t stage_u32 = smem_u32(Stage);
    uint32_t peer[2] = {stage_u32, stage_u32};
    if (P.spl > 1) { peer[0] = dsmem_addr(stage_u32, 0); peer[1] = dsmem_addr(stage_u32, 1); }
    for (int it = e; it < BM * G; it += 256) {
      const int r = it / G;
      const int64_t m = m0 + r;
      const uint32_t off0 = (uint32_t)r * C * 4 + ((((col >> 2)) ^ (r & 7)) << 4);
      const uint32_t off1 = (uint32_t)r * C * 4 + ((((col >> 2) + 1) ^ (r & 7)) << 4);
      float4 a0, a1;
      if (P.spl > 1) {
        a0 = ld_dsmem_f4(peer[0] + off0); a1 = ld_dsmem_f4(peer[0] + off1);
        const float4 c0 = ld_dsmem_f4(peer[1] + off0), c1 = ld_dsmem_f4(peer[1] + off1);
        a0.x += c0.x; a0.y += c0.y; a0.z += c0.z; a0.w += c0.w;
        a1.x += c1.x; a1.y += c1.y; a1.z += c1.z; a1.w += c1.w;
      } else {
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a0.x), "=f"(a0.y), "=f"(a0.z), "=f"(a0.w) : "r"(stage_u32 + off0));
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a1.x), "=f"(a1.y), "=f"(a1.z), "=f"(a1.w) : "r"(stage_u32 + off1));
      }
      float st1 = 0.f, st2 = 0.f;
      if (m < P.M) {
        const uint4 rv = *reinterpret_cast<const uint4*>(P.x + m * P.ldx + col);
        // the MLP branch is rounded to bf16 first, then added to the bf16 residual (same two roundings as Linear + residual)
        const uint32_t y0 = pack_bf16x2(a0.x + bA.x, a0.y + bA.y), y1 = pack_bf16x2(a0.z + bA.z, a0.w + bA.w);
        const uint32_t y2 = pack_bf16x2(a1.x + bB.x, a1.y + bB.y), y3 = pack_bf16x2(a1.z + bB.z, a1.w + bB.w);
        uint4 w;
        w.x = add_bf16x2(y0, rv.x); w.y = add_bf16x2(y1, rv.y); w.z = add_bf16x2(y2, rv.z); w.w = add_bf16x2(y3, rv.w);
        *reinterpret_cast<uint4*>(P.out + m * P.ldo + col) = w;
        const float e0 = bf16_lo(w.x), e1 = bf16_hi(w.x), e2 = bf16_lo(w.y), e3 = bf16_hi(w.y);
        const float e4 = bf16_lo(w.z), e5 = bf16_hi(w.z), e6 = bf16_lo(w.w), e7 = bf16_hi(w.w);
        st1 = ((e0 + e1) + (e2 + e3)) + ((e4 + e5) + (e6 + e7));
        st2 = fmaf(e0, e0, fmaf(e1, e1, fmaf(e2, e2, fmaf(e3, e3, fmaf(e4, e4, fmaf(e5, e5, fmaf(e6, e6, e7 * e7)))))));
      }
      if (P.stats_out != nullptr) {                          // G consecutive lanes hold one row's columns
        for (int o = 1; o < G; o <<= 1) { st1 += __shfl_xor_sync(0xffffffffu, st1, o); st2 += __shfl_xor_sync(0xffffffffu, st2, o); }
        if (g == 0 && m < P.M) { P.stats_out[(m * P.spl + rank) * 2] = st1; P.stats_out[(m * P.spl + rank) * 2 + 1] = st2; }
      }
    }
  }
  if (tid == 0) trace_stamp(P.trace, 7);
  // no CTA may exit (its shared memory would be released) while its peer is still reading it
  if (P.spl > 1) cluster_sync_all(); else __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols); }
}

size_t mlp_smem_bytes(int C, int slots) {
  return 1024 + (size_t)BM * C * 2 + (size_t)slots * 128 * C + 2 * 16384 + (2 * kMaxHidPerCta + 256) * 4 + kNumBars * 8 + 64;
}

struct MlpCfg { int spl, slots, tmem_cols; };
bool mlp_cfg(int C, int hidden, MlpCfg* c) {
  if (hidden != 4 * C) return false;
  if (C == 64) *c = MlpCfg{1, 6, 256};          // ring depth: the weight stream must cover ~1 us of TMA latency
  else if (C == 128) *c = MlpCfg{1, 8, 256};
  else if (C == 256) *c = MlpCfg{2, 3, 512};
  else return false;
  return true;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int mlp_tc_stats_parts(int C, int hidden) {
  MlpCfg c;
  return mlp_cfg(C, hidden, &c) ? c.spl : 0;
}

int mlp_fwd_tc(const cswin_mlp_args_t* a, cudaStream_t stream) {
  MlpCfg cfg;
  CSWIN_REQUIRE(mlp_cfg(a->C, a->hidden, &cfg), CSWIN_ERR_UNSUPPORTED,
                "mlp_fwd: the fused tcgen05 MLP exists for C in {64, 128, 256} with hidden = 4 C (use two cswin_linear_fwd calls otherwise)");
  CSWIN_REQUIRE(aligned16(a->x) && aligned16(a->w1) && aligned16(a->w2) && aligned16(a->out) && (a->ldx * 2) % 16 == 0 &&
                (a->ldo * 2) % 16 == 0 && (a->ldw1 * 2) % 16 == 0 && (a->ldw2 * 2) % 16 == 0 && a->M <= 0x7fffffff,
                CSWIN_ERR_UNSUPPORTED, "mlp_fwd: operands are not TMA-compatible (16-byte aligned pointers / row pitches)");
  CSWIN_REQUIRE(tc::encode_tiled_fn() != nullptr, CSWIN_ERR_CUDA, "mlp_fwd: the driver does not expose cuTensorMapEncodeTiled");
  if (a->M == 0) return CSWIN_OK;

  MlpTcParams P;
  memset(&P, 0, sizeof(P));
  P.b1 = a->b1; P.cs1 = a->ln_colsum; P.b2 = a->b2;
  P.ln_stats = a->ln_stats; P.ln_parts = a->ln_stats_parts; P.ln_invC = 1.0f / (float)a->C; P.ln_eps = a->ln_eps;
  P.x = (const __nv_bfloat16*)a->x; P.ldx = a->ldx;
  P.out = (__nv_bfloat16*)a->out; P.ldo = a->ldo;
  P.stats_out = a->stats_out;
  P.M = a->M; P.C = a->C; P.spl = cfg.spl; P.slots = cfg.slots; P.tmem_cols = cfg.tmem_cols;
  P.hid_per_cta = a->hidden / cfg.spl;
  P.nsub = P.hid_per_cta / HC;
  P.trace = g_trace.load(std::memory_order_relaxed);
  {
    const uint64_t dims[2] = {(uint64_t)a->C, (uint64_t)a->M}, str[1] = {(uint64_t)a->ldx * 2};
    const uint32_t box[2] = {64, BM};
    if (!tc::make_tensor_map_bf16(&P.map_x, a->x, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return CSWIN_ERR_CUDA;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a->C, (uint64_t)a->hidden}, str[1] = {(uint64_t)a->ldw1 * 2};
    const uint32_t box[2] = {64, HC};
    if (!tc::make_tensor_map_bf16(&P.map_w1, a->w1, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return CSWIN_ERR_CUDA;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a->hidden, (uint64_t)a->C}, str[1] = {(uint64_t)a->ldw2 * 2};
    const uint32_t box[2] = {HC, (uint32_t)a->C};
    if (!tc::make_tensor_map_bf16(&P.map_w2, a->w2, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return CSWIN_ERR_CUDA;
  }
  const size_t smem = mlp_smem_bytes(a->C, cfg.slots);
  static std::atomic<int> configured{0};
  if (!configured.load(std::memory_order_acquire)) {
    CSWIN_CUDA_OK(cudaFuncSetAttribute(mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured.store(1, std::memory_order_release);
  }
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)((a->M + BM - 1) / BM), (unsigned)cfg.spl);
  lc.blockDim = dim3(kThreads);
  lc.dynamicSmemBytes = smem;
  lc.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cfg.spl > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1; attr[na].val.clusterDim.y = (unsigned)cfg.spl; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  lc.attrs = attr; lc.numAttrs = na;
  CSWIN_CUDA_OK(cudaLaunchKernelEx(&lc, mlp_tc_kernel, P));
  CSWIN_LAUNCH_CHECK();
  g_tc_launches.fetch_add(1, std::memory_order_relaxed);
  return CSWIN_OK;
}

}  // namespace cswin
