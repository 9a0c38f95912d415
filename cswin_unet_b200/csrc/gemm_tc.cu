// gemm_tc.cu — bf16 Linear on tcgen05 / TMEM / TMA with the fused epilogue of cswin_linear_fwd (sm_100a).
//
//   out[m,n] = residual[m,n] + sample_scale[m / rps] * act( sum_k [a | a2][m,k] * w[n,k] + bias[n] )
//
// Replaces nn.Linear and the separate element-wise kernels of the reference around it (networks/cswin_unet.py:169,
// :177-178, Mlp :22-26 + :179, concat_linear :509-527, and the 1x1 / im2col'ed convs :216, :240-241, :264, :339, :542).
//
// One 128 x BN output tile per CTA (BN = 16..256, picked by a latency model so that the grid covers the 148 SMs at least twice where
// the problem allows), K consumed in 64-wide blocks (128-byte rows, 128-byte swizzle) through a 1..8 stage TMA -> smem ring.  Warp
// roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread tcgen05.mma issuer (elect.sync lane), warps 2..9 = epilogue.
// The two-source form ([skip | x] for the decoder's concat Linears) switches tensor maps at the K1 boundary, so the concatenated
// activation is never materialised.  Ragged M / N / K tails are handled by TMA out-of-bounds zero fill on the loads and by TMA clipping
// (or predicated stores) on the way out.
//
// Epilogue, in the accumulator layout (thread = row), per 32-column unit: tcgen05.ld -> folded LayerNorm / bias / GELU / DropPath scale
// in fp32 registers (packed FFMA2) -> bf16 -> packed residual add -> row statistics for the next folded LayerNorm -> 64-byte-swizzled
// staging box -> one TMA store per 32 x 32 unit.  Outputs whose rows are not 16-byte aligned take the older path (fp32 staging read back
// row-contiguous, predicated 16-byte stores).
//
// Forms of the same kernel (template / runtime parameters, see the comments at each):
//   kPersist      resident CTAs walk the tiles, accumulator double-buffered in TMEM, ring running across tiles (multi-wave launches
//                 with a light epilogue); kEW = 16: the one-CTA-per-SM variant with 16 epilogue warps (experiment);
//   P.conv        implicit-GEMM convolution: the A operand is fetched as strided 4-D TMA boxes of a (B, H, W, C) token image
//                 (cswin_conv_tokens_fwd: Merge_Block.conv, CARAFE.encoder — networks/cswin_unet.py:214-216, :240-241);
//   kTrain        the two training-only epilogues (aux pre-activation output, multiply by GELU'(z));
//   P.w_kn        weights read as (K, N) row-major = MN-major B operand (data gradient dA = dZ W).
#include <climits>
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace cswin {
namespace {

using namespace tc;

// -DCSWIN_GEMM_PROFILE: cycle stamps of the first epilogue unit of warp 2 into debug trace slots 8..15 (tools/trace_kernel.py)
#ifdef CSWIN_GEMM_PROFILE
#define PSTAMP(i) do { if (warp == 2 && lane == 0 && u == ch + 2 * CSWIN_GEMM_PROFILE && P.trace != nullptr && blockIdx.x + blockIdx.y * gridDim.x < 1024) \
  P.trace[(size_t)(blockIdx.x + blockIdx.y * gridDim.x) * 16 + (i)] = clock64(); } while (0)
#else
#define PSTAMP(i) do { } while (0)
#endif

constexpr int BM = 128, BK = 64;
constexpr int kMaxStages = 8;
#ifndef CSWIN_GEMM_MINB
#define CSWIN_GEMM_MINB 2                 // CTAs per SM the register budget is sized for (3 -> 64 registers, minor spills)
#endif
constexpr int kThreads = 320;                          // warp 0 TMA, warp 1 MMA + TMEM, warps 2..9 epilogue

struct alignas(64) GemmTcParams {
  CUtensorMap map_a, map_a2, map_w, map_out, map_aux;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* res; int64_t ldr;
  const float* sscale; int rps;
  __nv_bfloat16* out; int64_t ldo;
  int64_t M; int N; int K1; int K2;
  int BN, stages, nkb, nkb1, act, tmem_cols, vec_ok, bias_vec, w_kn, tma_out;
  int nt_n, ntiles, acc_stride;          // N tiles per row of tiles; persistent form: total tiles, TMEM columns per accumulator buffer
  // implicit-GEMM convolution over a (B, H, W, C) token image (cswin_conv_tokens_fwd): map_a is the 4-D image map with element
  // strides (1, stride, stride, 1); an M tile = cv_rows output pixels = cv_bh full output rows of one image (cv_nb = 1) or cv_nb
  // whole images; K block kb = 64 channels of tap kb / cv_cblk
  int conv, cv_rows, cv_tpi, cv_nb, cv_bh, cv_stride, cv_pad, cv_KW, cv_cblk;
  const float* ln_stats; int ln_parts; float ln_invC, ln_eps;
  const float* ln_cs; const float* bias_f32; float* stats_out;
  int aux, aux_off;                      // training epilogue: also store the pre-activation (map_aux); staging offset in smem
  unsigned long long* trace;
};

// GELU'(x) for a packed pair: tc::gelu_grad_fast (clamped fit, shared with the streaming activation kernels in backward.cu)
__device__ __forceinline__ float2 gelu_grad_pair(uint32_t zz) {
  return make_float2(gelu_grad_fast(bf16_lo(zz)), gelu_grad_fast(bf16_hi(zz)));
}

// kTrain adds the two training-only epilogues (TMA-store fast path only):
//   aux : out = GELU(z) AND z itself goes to a second tensor (fc1 of the MLP: the backward needs the pre-activation);
//   act 2: out = acc * GELU'(residual) — the data gradient of fc2 multiplied by GELU'(z) in place of a separate pass.
//
// kPersist: the CTA walks tiles `blockIdx.x + i * gridDim.x` (N tile fastest, so that CTAs running at the same time share A rows
// in L2).  The operand ring runs on across tiles, the accumulator is double-buffered in TMEM (2 x acc_stride columns) and the
// epilogue staging no longer aliases the ring: the TMA / MMA warps work on tile i+1 while the epilogue warps drain tile i, and
// barrier init / TMEM allocation / descriptor prefetch are paid once per CTA.  Used when a launch is more than one wave of CTAs.
//
// kEW = 16 (persistent form only): ONE CTA per SM with 16 epilogue warps (640 threads), 128 x BN <= 256 tiles, both 256-column halves of
// TMEM as the two accumulator buffers and a 3..8 stage ring — the same 16 epilogue warps per SM as two 320-thread CTAs, but they never
// wait for operands: while they drain tile i the producer already has all K blocks of tile i + 1 in flight.
template <bool kFold, bool kStats, bool kTrain = false, bool kPersist = false, int kEW = 8>
__global__ void __launch_bounds__(64 + 32 * kEW, kEW == 16 ? 1 : CSWIN_GEMM_MINB) linear_tc_kernel(const __grid_constant__ GemmTcParams P) {
  constexpr int kUS = kEW / 4;                           // unit stride: epilogue warps per TMEM lane quadrant
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128-byte swizzle; plain pointer arithmetic keeps the shared address space (LDS/STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int BN = P.BN, S = P.stages;
  const uint32_t a_bytes = BM * BK * 2, w_bytes = (uint32_t)BN * BK * 2;
  uint8_t* As = smem;                                   // [S][128][64] bf16, SW128
  uint8_t* Ws = As + (size_t)S * a_bytes;               // [S][BN][64]
  // epilogue staging (8 warps x 32 rows x 64 B, bf16, swizzled) re-uses the operand ring: it is only touched after
  // bar_acc, i.e. after every TMA load has landed and every MMA has finished reading shared memory
  const size_t ring = (size_t)S * (a_bytes + w_bytes);
  uint8_t* Epi = kPersist ? smem + ring : smem;                          // persistent form: own 16 KB behind the ring
  constexpr int kCB = kPersist ? 2 : 1;                                  // column-constant buffers (one per accumulator buffer)
  float* sBias = reinterpret_cast<float*>(smem + ring + (kPersist ? kEW * 2048 : 0));   // [kCB][256] fp32
  float* sCs = sBias + kCB * 256;                                        // [kCB][256] folded-LayerNorm column sums
  float* sStat = sCs + kCB * 256;                                        // [128][2] per-row (sum, sum^2) of this tile's output
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + 256);             // full[8], empty[8], acc_full[2], acc_empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  if (tid == 0) trace_stamp(P.trace, 0);                                   // kernel entry
  const int nt_n = P.nt_n;
  // first (or only) tile of this CTA
  int64_t m0 = kPersist ? (int64_t)((int)blockIdx.x / nt_n) * BM : (int64_t)blockIdx.x * (P.conv ? P.cv_rows : BM);
  const uint32_t a_tx = P.conv ? (uint32_t)P.cv_rows * 128u : a_bytes;      // bytes one A box delivers
  int n0 = kPersist ? ((int)blockIdx.x % nt_n) * BN : (int)blockIdx.y * BN;
  int n_tile = kPersist ? (int)blockIdx.x % nt_n : (int)blockIdx.y;

  auto full = [&](int s) { return smem_u32(&bars[s]); };
  auto empty = [&](int s) { return smem_u32(&bars[kMaxStages + s]); };
  auto acc_full = [&](int b) { return smem_u32(&bars[2 * kMaxStages + b]); };
  auto acc_empty = [&](int b) { return smem_u32(&bars[2 * kMaxStages + 2 + b]); };
  const uint32_t bar_acc = acc_full(0);

  // One elected thread initialises the barriers and immediately requests the W tiles of the first ring pass: weights do not
  // depend on the previous kernel (PDL) nor on the rest of this CTA's prologue (TMEM allocation, bias loads).
  const int npre = P.nkb < S ? P.nkb : S;
  if (warp == 0 && elect_one()) {
    for (int s = 0; s < S; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    mbar_init(acc_full(0), 1);
    if (kPersist) { mbar_init(acc_full(1), 1); mbar_init(acc_empty(0), kEW); mbar_init(acc_empty(1), kEW); }
    fence_barrier_init();
    fence_proxy_async();
    tma_prefetch_desc(&P.map_w);
    for (int kb = 0; kb < npre; ++kb) {
      mbar_expect_tx(full(kb), a_tx + w_bytes);
      if (!P.w_kn) tma_load_2d(smem_u32(Ws + (size_t)kb * w_bytes), &P.map_w, full(kb), kb * BK, n0);
      else for (int j = 0; j * 64 < BN; ++j)             // (K, N) weight: [64 k][64 n] boxes = MN-major B blocks
        tma_load_2d(smem_u32(Ws + (size_t)kb * w_bytes + j * 8192), &P.map_w, full(kb), n0 + 64 * j, kb * BK);
    }
    tma_prefetch_desc(&P.map_a);
    if (P.K2 > 0) tma_prefetch_desc(&P.map_a2);
  }
  if (warp == 1) { tmem_alloc(smem_u32(tmem_slot), (uint32_t)P.tmem_cols); tmem_relinquish(); }
  float bias_r = 0.f, cs_r = 0.f;                       // per-column constants: loaded now, parked in smem by the epilogue warps
  if (!kPersist && warp >= 2) {                         // (the loads stay in flight across the CTA barrier)
    const int j = tid - 64, n = n0 + j;
    if (j < BN && n < P.N) {
      bias_r = P.bias_f32 != nullptr ? P.bias_f32[n] : P.bias != nullptr ? __bfloat162float(P.bias[n]) : 0.f;
      if (kFold) cs_r = P.ln_cs[n];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) trace_stamp(P.trace, 1);                                   // prologue done
  pdl_wait();                                           // activations (A, residual) and the output buffer are safe from here

  if (warp == 0) {
    if (elect_one()) {                                  // ---- TMA producer (elect.sync keeps the warp-uniform datapath) ----
      int s = 0; uint32_t ph = 1;                       // ring slot and pass parity, advanced without div / mod
      if (kPersist) {
        int g = 0;                                      // ring uses so far (W of the first `npre` was requested in the prologue)
        for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
          const int tm0 = (tile / nt_n) * BM, tn0 = (tile % nt_n) * BN;
          for (int kb = 0; kb < P.nkb; ++kb, ++g) {
            if (g >= npre) {
              if (g >= S) mbar_wait(empty(s), ph);
              mbar_expect_tx(full(s), a_bytes + w_bytes);
              if (!P.w_kn) tma_load_2d(smem_u32(Ws + (size_t)s * w_bytes), &P.map_w, full(s), kb * BK, tn0);
              else for (int j = 0; j * 64 < BN; ++j)
                tma_load_2d(smem_u32(Ws + (size_t)s * w_bytes + j * 8192), &P.map_w, full(s), tn0 + 64 * j, kb * BK);
            }
            if (kb < P.nkb1) tma_load_2d(smem_u32(As + (size_t)s * a_bytes), &P.map_a, full(s), kb * BK, tm0);
            else             tma_load_2d(smem_u32(As + (size_t)s * a_bytes), &P.map_a2, full(s), (kb - P.nkb1) * BK, tm0);
            if (++s == S) { s = 0; ph ^= 1; }
          }
        }
      } else
      {
      // implicit conv: image / first output row of this tile
      const int cb0 = !P.conv ? 0 : P.cv_nb > 1 ? (int)blockIdx.x * P.cv_nb : (int)blockIdx.x / P.cv_tpi;
      const int cy0 = !P.conv || P.cv_nb > 1 ? 0 : ((int)blockIdx.x % P.cv_tpi) * P.cv_bh * P.cv_stride - P.cv_pad;
      for (int kb = 0; kb < P.nkb; ++kb) {
        if (kb >= S) {
          mbar_wait(empty(s), ph);
          mbar_expect_tx(full(s), a_tx + w_bytes);
          if (!P.w_kn) tma_load_2d(smem_u32(Ws + (size_t)s * w_bytes), &P.map_w, full(s), kb * BK, n0);
          else for (int j = 0; j * 64 < BN; ++j)
            tma_load_2d(smem_u32(Ws + (size_t)s * w_bytes + j * 8192), &P.map_w, full(s), n0 + 64 * j, kb * BK);
        }
        if (P.conv) {                                    // tap (ky, kx), 64 channels: a strided box of the token image, zero padding = TMA OOB fill
          const int tap = kb / P.cv_cblk, c0 = (kb - tap * P.cv_cblk) * BK;
          const int ky = tap / P.cv_KW, kx = tap - ky * P.cv_KW;
          tma_load_4d(smem_u32(As + (size_t)s * a_bytes), &P.map_a, full(s), c0, kx - P.cv_pad, cy0 + (P.cv_nb > 1 ? ky - P.cv_pad : ky), cb0);
        }
        else if (kb < P.nkb1) tma_load_2d(smem_u32(As + (size_t)s * a_bytes), &P.map_a, full(s), kb * BK, (int)m0);
        else                  tma_load_2d(smem_u32(As + (size_t)s * a_bytes), &P.map_a2, full(s), (kb - P.nkb1) * BK, (int)m0);
        if (++s == S) { s = 0; ph ^= 1; }
      }
      }
      trace_stamp(P.trace, 2);                          // all TMA issued
    }
  } else if (warp == 1) {
    if (elect_one()) {                                  // ---- MMA issuer ----
      const uint32_t idesc = make_idesc_bf16(BM, BN, 0, P.w_kn);
      const uint32_t wstep = P.w_kn ? (2048 >> 4) : 2;   // 16 contraction rows: 2 KB (MN-major) or 32 B (K-major) further
      int s = 0; uint32_t ph = 0;
      int it = 0;
      for (int tile = kPersist ? (int)blockIdx.x : 0; tile < (kPersist ? P.ntiles : 1); tile += kPersist ? (int)gridDim.x : 1, ++it) {
        const int buf = it & 1;
        if (kPersist && it >= 2) {                      // the epilogue warps have drained this buffer's previous tile
          mbar_wait(acc_empty(buf), (uint32_t)(((it >> 1) - 1) & 1));
          tc_fence_after();
        }
        const uint32_t d_tmem = tmem_base + (kPersist ? (uint32_t)(buf * P.acc_stride) : 0u);
        for (int kb = 0; kb < P.nkb; ++kb) {
          mbar_wait(full(s), ph);
          if (kb == 0 && it == 0) trace_stamp(P.trace, 3);    // first operands landed
          tc_fence_after();
          const uint64_t ad = make_smem_desc(smem_u32(As + (size_t)s * a_bytes), 16, 1024, kLayoutSw128);
          const uint64_t wd = make_smem_desc(smem_u32(Ws + (size_t)s * w_bytes), P.w_kn ? 8192 : 16, 1024, kLayoutSw128);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) mma_ss(d_tmem, ad + 2 * k, wd + (uint64_t)wstep * k, idesc, (kb | k) != 0);
          tc_commit(empty(s));                          // smem slot reusable once these MMAs have read it
          if (++s == S) { s = 0; ph ^= 1; }
        }
        tc_commit(acc_full(buf));                       // accumulator complete
      }
      trace_stamp(P.trace, 4);                          // all MMAs issued
    }
  } else {
    // ---- epilogue: warps 2..9.  TMEM lane quadrant q = warp % 4 (hardware rule); the two warps of a quadrant take the
    //      even / odd 32-column units.  Per unit: tcgen05.ld -> bias / GELU / DropPath scale in registers -> bf16 ->
    //      swizzled 2 KB staging tile -> row-contiguous read-back so the residual load and the store are coalesced 16 B.
    const int q = warp & 3;
    const int ch = (warp - 2) >> 2;
    if (!kPersist) {
      sBias[tid - 64] = bias_r;
      if (kFold) sCs[tid - 64] = cs_r;
    }
    if (kStats && tid - 64 < 256) sStat[tid - 64] = 0.f;
    if (!kPersist) asm volatile("bar.sync 1, %0;" ::"n"(32 * kEW) : "memory");      // the epilogue warps only
    uint8_t* stg = Epi + (warp - 2) * 2048;
    const uint32_t stg_u32 = smem_u32(stg);
    const int nunits = (BN + 31) >> 5;
    const int rows_valid = P.conv ? P.cv_rows : BM;     // implicit conv: an M tile may hold fewer than 128 output pixels
    int it = 0;
    for (int tile = kPersist ? (int)blockIdx.x : 0; tile < (kPersist ? P.ntiles : 1); tile += kPersist ? (int)gridDim.x : 1, ++it) {
    const int buf = kPersist ? (it & 1) : 0;
    if (kPersist) {                                     // this tile's coordinates and column constants
      n_tile = tile % nt_n; m0 = (int64_t)(tile / nt_n) * BM; n0 = n_tile * BN;
      const int j = tid - 64, n = n0 + j;
      float b = 0.f, c = 0.f;
      if (j < BN && n < P.N) {
        b = P.bias_f32 != nullptr ? P.bias_f32[n] : P.bias != nullptr ? __bfloat162float(P.bias[n]) : 0.f;
        if (kFold) c = P.ln_cs[n];
      }
      if (j < 256) {
        sBias[buf * 256 + j] = b;
        if (kFold) sCs[buf * 256 + j] = c;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEW) : "memory");    // also orders the previous tile's statistics hand-off
    }
    const float* sB = sBias + buf * 256;
    const float* sC = sCs + buf * 256;
    if (P.res != nullptr && P.vec_ok) {                 // pull this thread's residual segments towards L1 while the MMAs run
      for (int u = ch; u < ((BN + 31) >> 5); u += kUS) {
        const int n = n0 + u * 32 + (lane & 3) * 8;
#pragma unroll
        for (int pass = 0; pass < 4; ++pass) {
          const int64_t m = m0 + q * 32 + pass * 8 + (lane >> 2);
          if (m < P.M && n < P.N) asm volatile("prefetch.global.L1 [%0];" ::"l"(P.res + m * P.ldr + n));
        }
      }
    }
    const int64_t mrow = m0 + q * 32 + lane;            // accumulator row held by this thread in phase 1
    const float sc = (P.sscale != nullptr && mrow < P.M) ? P.sscale[mrow / P.rps] : 1.0f;
    float ln_mean = 0.f, ln_rstd = 1.f;                 // folded LayerNorm: row statistics from the producer's partial sums
    if (kFold && mrow < P.M) {
      float s1 = 0.f, s2 = 0.f;
      for (int p = 0; p < P.ln_parts; ++p) { s1 += P.ln_stats[(mrow * P.ln_parts + p) * 2]; s2 += P.ln_stats[(mrow * P.ln_parts + p) * 2 + 1]; }
      ln_mean = s1 * P.ln_invC;
      ln_rstd = rsqrtf(fmaxf(fmaf(-ln_mean, ln_mean, s2 * P.ln_invC), 0.f) + P.ln_eps);
    }
    mbar_wait(acc_full(buf), kPersist ? (uint32_t)((it >> 1) & 1) : 0u);
    if (warp == 2 && lane == 0 && it == 0) trace_stamp(P.trace, 5);  // accumulator ready
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (kPersist ? (uint32_t)(buf * P.acc_stride) : 0u);
    if (P.tma_out) {
      // ---- fast path (16-byte aligned rows, N % 8 == 0): everything happens in the accumulator layout (thread = row): bias /
      //      GELU / scale -> bf16 -> residual add (packed bf16x2) -> row statistics -> 64B-swizzled staging box -> one TMA
      //      store per 32 x 32 unit.  No shared-memory read-back, no per-thread global stores, ragged edges clipped by TMA.
      for (int u = ch; u < nunits; u += kUS) {
        uint32_t v[32];
        PSTAMP(8);
        tmem_ld32(trow + u * 32, v);
        uint4 rv[4];
        const int ncol = n0 + u * 32;
        const bool has_res = P.res != nullptr && mrow < P.M;
        if (kTrain && !has_res) { rv[0] = rv[1] = rv[2] = rv[3] = make_uint4(0, 0, 0, 0); }
        if (has_res) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            rv[c] = (ncol + c * 8 < P.N) ? *reinterpret_cast<const uint4*>(P.res + mrow * P.ldr + ncol + c * 8) : make_uint4(0, 0, 0, 0);
        }
        tmem_wait_ld();
        PSTAMP(9);
        if (kPersist && u + kUS >= nunits) {                // last unit of this warp: the accumulator buffer is free for tile i + 2
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty(buf));
        }
        const float4* b4 = reinterpret_cast<const float4*>(sB + u * 32);
        const float4* c4 = reinterpret_cast<const float4*>(sC + u * 32);
        if (u != ch) {                                    // the previous unit's TMA store must have finished reading the staging box
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
        }
        PSTAMP(10);
        float st1 = 0.f, st2 = 0.f;
        const float2 nmean2 = make_float2(-ln_mean, -ln_mean), rstd2 = make_float2(ln_rstd, ln_rstd), sc2 = make_float2(sc, sc);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float2 f[4];                                    // 8 columns as 4 packed fp32 pairs (FFMA2 / FADD2)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 bb = b4[c * 2 + h];
            const float2 a0 = make_float2(__uint_as_float(v[c * 8 + h * 4 + 0]), __uint_as_float(v[c * 8 + h * 4 + 1]));
            const float2 a1 = make_float2(__uint_as_float(v[c * 8 + h * 4 + 2]), __uint_as_float(v[c * 8 + h * 4 + 3]));
            if (kFold) {                                  // rstd * (acc - mean * colsum) + bias'
              const float4 cc = c4[c * 2 + h];
              f[h * 2 + 0] = ffma2(rstd2, ffma2(nmean2, make_float2(cc.x, cc.y), a0), make_float2(bb.x, bb.y));
              f[h * 2 + 1] = ffma2(rstd2, ffma2(nmean2, make_float2(cc.z, cc.w), a1), make_float2(bb.z, bb.w));
            } else {
              f[h * 2 + 0] = fadd2(a0, make_float2(bb.x, bb.y));
              f[h * 2 + 1] = fadd2(a1, make_float2(bb.z, bb.w));
            }
          }
          if (kTrain && P.aux) {                            // pre-activation z (bf16) into the second staging box
            const uint32_t aaddr = stg_u32 + (uint32_t)P.aux_off + lane * 64 + (((c ^ (lane >> 1)) & 3) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(aaddr), "r"(pack_bf16x2(f[0].x, f[0].y)), "r"(pack_bf16x2(f[1].x, f[1].y)),
                         "r"(pack_bf16x2(f[2].x, f[2].y)), "r"(pack_bf16x2(f[3].x, f[3].y)) : "memory");
          }
          if (P.act == 1) {
#pragma unroll
            for (int e = 0; e < 4; ++e) f[e] = gelu_fast2(f[e]);
          }
          if (kTrain && P.act == 2) {                       // dZ = dH o GELU'(z), z = `residual` operand
            f[0] = fmul2(f[0], gelu_grad_pair(rv[c].x)); f[1] = fmul2(f[1], gelu_grad_pair(rv[c].y));
            f[2] = fmul2(f[2], gelu_grad_pair(rv[c].z)); f[3] = fmul2(f[3], gelu_grad_pair(rv[c].w));
          }
          if (P.sscale != nullptr) {
#pragma unroll
            for (int e = 0; e < 4; ++e) f[e] = fmul2(f[e], sc2);
          }
          uint4 x = make_uint4(pack_bf16x2(f[0].x, f[0].y), pack_bf16x2(f[1].x, f[1].y), pack_bf16x2(f[2].x, f[2].y), pack_bf16x2(f[3].x, f[3].y));
          if (has_res && !(kTrain && P.act == 2)) {
            x.x = add_bf16x2(x.x, rv[c].x); x.y = add_bf16x2(x.y, rv[c].y);
            x.z = add_bf16x2(x.z, rv[c].z); x.w = add_bf16x2(x.w, rv[c].w);
          }
          if (kStats && ncol + c * 8 < P.N) {             // (sum, sum^2) of the bf16 values this row contributes
            const float e0 = bf16_lo(x.x), e1 = bf16_hi(x.x), e2 = bf16_lo(x.y), e3 = bf16_hi(x.y);
            const float e4 = bf16_lo(x.z), e5 = bf16_hi(x.z), e6 = bf16_lo(x.w), e7 = bf16_hi(x.w);
            st1 += ((e0 + e1) + (e2 + e3)) + ((e4 + e5) + (e6 + e7));
            st2 = fmaf(e0, e0, fmaf(e1, e1, fmaf(e2, e2, fmaf(e3, e3, fmaf(e4, e4, fmaf(e5, e5, fmaf(e6, e6, fmaf(e7, e7, st2))))))));
          }
          const uint32_t addr = stg_u32 + lane * 64 + (((c ^ (lane >> 1)) & 3) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w) : "memory");
        }
        if (kStats) { atomicAdd(&sStat[(q * 32 + lane) * 2], st1); atomicAdd(&sStat[(q * 32 + lane) * 2 + 1], st2); }
        PSTAMP(11);
        fence_proxy_async();
        __syncwarp();
        PSTAMP(12);
        if (lane == 0) {
          tma_store_2d(&P.map_out, stg_u32, ncol, (int)(m0 + q * 32));
          if (kTrain && P.aux) tma_store_2d(&P.map_aux, stg_u32 + (uint32_t)P.aux_off, ncol, (int)(m0 + q * 32));
          tma_store_commit();
        }
        PSTAMP(13);
      }
      if (lane == 0) tma_store_wait_read();               // shared memory may be released / re-used after this
      __syncwarp();
      if (kPersist && ch >= nunits && lane == 0) mbar_arrive(acc_empty(buf));   // BN <= 32: this warp had no unit, it still signs off
    } else
    for (int u = ch; u < nunits; u += kUS) {
      uint32_t v[32];
      PSTAMP(8);
      tmem_ld32(trow + u * 32, v);
      tmem_wait_ld();
      PSTAMP(9);
      const float4* b4 = reinterpret_cast<const float4*>(sB + u * 32);
      const float4* c4 = reinterpret_cast<const float4*>(sC + u * 32);
#pragma unroll
      for (int c = 0; c < 4; ++c) {                     // four 16-byte chunks of 8 columns
        float f[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 bb = b4[c * 2 + h];
          if (kFold) {                                  // rstd * (acc - mean * colsum) + bias'
            const float4 cc = c4[c * 2 + h];
            f[h * 4 + 0] = fmaf(ln_rstd, fmaf(-ln_mean, cc.x, __uint_as_float(v[c * 8 + h * 4 + 0])), bb.x);
            f[h * 4 + 1] = fmaf(ln_rstd, fmaf(-ln_mean, cc.y, __uint_as_float(v[c * 8 + h * 4 + 1])), bb.y);
            f[h * 4 + 2] = fmaf(ln_rstd, fmaf(-ln_mean, cc.z, __uint_as_float(v[c * 8 + h * 4 + 2])), bb.z);
            f[h * 4 + 3] = fmaf(ln_rstd, fmaf(-ln_mean, cc.w, __uint_as_float(v[c * 8 + h * 4 + 3])), bb.w);
          } else {
            f[h * 4 + 0] = __uint_as_float(v[c * 8 + h * 4 + 0]) + bb.x;
            f[h * 4 + 1] = __uint_as_float(v[c * 8 + h * 4 + 1]) + bb.y;
            f[h * 4 + 2] = __uint_as_float(v[c * 8 + h * 4 + 2]) + bb.z;
            f[h * 4 + 3] = __uint_as_float(v[c * 8 + h * 4 + 3]) + bb.w;
          }
        }
        if (P.act == 1) {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = gelu_fast(f[e]);
        }
        if (P.sscale != nullptr) {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] *= sc;
        }
        const uint32_t addr = stg_u32 + lane * 64 + (((c ^ (lane >> 1)) & 3) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pack_bf16x2(f[0], f[1])),
                     "r"(pack_bf16x2(f[2], f[3])), "r"(pack_bf16x2(f[4], f[5])), "r"(pack_bf16x2(f[6], f[7])) : "memory");
      }
      __syncwarp();
      PSTAMP(10);
      // 8 rows x 64 B per pass, 4 lanes per row; the four passes' shared-memory reads and residual loads are all issued
      // before the first use so their latencies overlap
      const int c = lane & 3;
      const int n = n0 + u * 32 + c * 8;
      const bool col_ok = n < P.N && u * 32 + c * 8 < BN;
      const bool vec = P.vec_ok && n + 8 <= P.N;
      uint4 w[4], rv[4];
#pragma unroll
      for (int pass = 0; pass < 4; ++pass) {
        const int r = pass * 8 + (lane >> 2);
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[pass].x), "=r"(w[pass].y), "=r"(w[pass].z), "=r"(w[pass].w)
                     : "r"(stg_u32 + r * 64 + (((c ^ (r >> 1)) & 3) << 4)));
      }
      if (P.res != nullptr && vec && col_ok) {
#pragma unroll
        for (int pass = 0; pass < 4; ++pass) {
          const int64_t m = m0 + q * 32 + pass * 8 + (lane >> 2);
          rv[pass] = m < P.M ? *reinterpret_cast<const uint4*>(P.res + m * P.ldr + n) : make_uint4(0, 0, 0, 0);
        }
      }
      PSTAMP(11);
#pragma unroll
      for (int pass = 0; pass < 4; ++pass) {
        const int r = pass * 8 + (lane >> 2);
        const int64_t m = m0 + q * 32 + r;
        if (pass == 0 && P.res != nullptr) { PSTAMP(12 + (int)((rv[0].x & 1) & 0)); }
        uint4 x = w[pass];
        float st1 = 0.f, st2 = 0.f;
        if (m < P.M && col_ok && q * 32 + r < rows_valid) {
          __nv_bfloat16* dst = P.out + m * P.ldo + n;
          if (vec) {
            if (P.res != nullptr) {
              const uint4 y = rv[pass];
              x.x = add_bf16x2(x.x, y.x); x.y = add_bf16x2(x.y, y.y);     // bf16 + bf16 -> bf16, one rounding (HADD2.BF16)
              x.z = add_bf16x2(x.z, y.z); x.w = add_bf16x2(x.w, y.w);
            }
            *reinterpret_cast<uint4*>(dst) = x;
            if (kStats) {                               // (sum, sum^2) of the bf16 values just stored, for the next folded LN
              const float e0 = bf16_lo(x.x), e1 = bf16_hi(x.x), e2 = bf16_lo(x.y), e3 = bf16_hi(x.y);
              const float e4 = bf16_lo(x.z), e5 = bf16_hi(x.z), e6 = bf16_lo(x.w), e7 = bf16_hi(x.w);
              st1 = ((e0 + e1) + (e2 + e3)) + ((e4 + e5) + (e6 + e7));
              st2 = fmaf(e0, e0, fmaf(e1, e1, fmaf(e2, e2, fmaf(e3, e3, fmaf(e4, e4, fmaf(e5, e5, fmaf(e6, e6, e7 * e7)))))));
            }
          } else {
            const uint32_t ww[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              if (n + e < P.N) {
                float t = (e & 1) ? bf16_hi(ww[e >> 1]) : bf16_lo(ww[e >> 1]);
                if (P.res != nullptr) t += __bfloat162float(P.res[m * P.ldr + n + e]);
                dst[e] = __float2bfloat16_rn(t);
              }
            }
          }
        }
        if (kStats) {                                   // the 4 lanes of a row -> one shared-memory accumulate per row
          st1 += __shfl_xor_sync(0xffffffffu, st1, 1); st2 += __shfl_xor_sync(0xffffffffu, st2, 1);
          st1 += __shfl_xor_sync(0xffffffffu, st1, 2); st2 += __shfl_xor_sync(0xffffffffu, st2, 2);
          if (c == 0) { atomicAdd(&sStat[(q * 32 + r) * 2], st1); atomicAdd(&sStat[(q * 32 + r) * 2 + 1], st2); }
        }
      }
      PSTAMP(13);
      __syncwarp();
    }
    if (kStats) {
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEW) : "memory");      // all epilogue warps have accumulated their units
      const int r = tid - 64;
      if (r < BM) {
        if (m0 + r < P.M && r < rows_valid) {
          float* dst = P.stats_out + ((m0 + r) * nt_n + n_tile) * 2;
          dst[0] = sStat[r * 2]; dst[1] = sStat[r * 2 + 1];
        }
        if (kPersist) { sStat[r * 2] = 0.f; sStat[r * 2 + 1] = 0.f; }   // visible to the others through the next tile's barrier
      }
    }
    }   // tile loop
    if (warp == 2 && lane == 0) trace_stamp(P.trace, 6);  // epilogue done
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) trace_stamp(P.trace, 7);                                   // exit
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
bool no_tma_out() { static const bool v = [] { const char* e = getenv("CSWIN_GEMM_TMA_OUT"); return e && e[0] == '0'; }(); return v; }  // A/B switch

size_t smem_bytes(int bn, int stages) {
  const size_t ring = (size_t)stages * (BM * BK * 2 + (size_t)bn * BK * 2);     // >= 18 KB > the aliased 16 KB staging
  return 1024 + ring + 3 * 1024 /* bias, column sums, row stats */ + 256 /* barriers, TMEM slot */;
}
int tmem_cols_for(int bn) { return bn <= 64 ? 64 : bn <= 128 ? 128 : 256; }      // the epilogue reads whole 64-col groups

size_t smem_bytes_persist(int bn, int stages, int ew = 8) {                       // + own staging, second column-constant buffer
  return smem_bytes(bn, stages) + (size_t)ew * 2048 + 2 * 1024;
}
// CSWIN_GEMM_PERSIST=0 switches the persistent form off (A/B runs); =2 forces it wherever it is legal
int persist_mode() { static const int v = [] { const char* e = getenv("CSWIN_GEMM_PERSIST"); return e ? atoi(e) : 1; }(); return v; }

struct TileCfg { int bn, stages; int persist, grid; int ew = 8; };

// Tile-shape choice.  At cswin_tiny sizes a Linear is a handful of waves at most, so the launch is latency- not
// throughput-bound: the model below (constants fitted to L2-warm CUDA-event timings on B200, microseconds) trades the
// number of waves against per-tile latency = fixed setup + K-loop (faster with a deeper ring) + epilogue (per 64
// columns; GELU costs ~2x).  The ring depth is whatever fits once the CTAs that must share an SM are accounted for.
// CSWIN_GEMM_BN=<n> forces BN for experiments.
TileCfg pick_tile(int64_t M, int N, int nkb, int act, int sms, bool w_kn, bool allow_persist = true) {
  static const int forced = [] { const char* e = getenv("CSWIN_GEMM_BN"); return e ? atoi(e) : 0; }();
  // CSWIN_GEMM_SMEM_CAP_KB: per-CTA shared-memory ceiling, so that a CTA of the NEXT kernel (PDL) fits next to the resident ones
  static const int env_cap = [] { const char* e = getenv("CSWIN_GEMM_SMEM_CAP_KB"); return e ? atoi(e) : -1; }();
  const size_t smem_cap = (size_t)(env_cap >= 0 ? env_cap : g_gemm_smem_cap_kb.load(std::memory_order_relaxed)) * 1024;   // env wins (A/B runs)
  const int n16 = w_kn ? ((N + 63) & ~63) : ((N + 15) & ~15);      // (K,N) weights are fetched in 64-column boxes
  const int64_t mt = (M + BM - 1) / BM;
  const int cands[] = {64, 96, 128, 192, 256};
  TileCfg best{n16 < 64 ? n16 : 64, 1, 0, 0};
  double best_t = 1e30;
  int64_t best_waves = 1;
  for (int bn : cands) {
    if (forced >= 16 && forced <= 256 && forced % 16 == 0) bn = forced;
    if (w_kn && bn % 64) continue;
    if (bn > n16) bn = n16 > 256 ? 256 : n16;
    if (N % bn != 0 && bn > 64 && bn != n16) continue;
    const int64_t tiles = mt * ((N + bn - 1) / bn);
    int resident = 512 / tmem_cols_for(bn);
    if (resident > CSWIN_GEMM_MINB) resident = CSWIN_GEMM_MINB;   // register budget: CSWIN_GEMM_MINB CTAs of 320 threads per SM
    while (resident > 1 && smem_bytes(bn, 2) * resident > 220 * 1024) --resident;
    const int64_t waves = (tiles + (int64_t)sms * resident - 1) / ((int64_t)sms * resident);
    int64_t share = (tiles + sms - 1) / sms;                    // CTAs that will actually share an SM
    if (share > resident) share = resident;
    int st = 1;
    while (st < kMaxStages && st < nkb && smem_bytes(bn, st + 1) * share <= 220 * 1024 &&
           (smem_cap == 0 || smem_bytes(bn, st + 1) <= smem_cap)) ++st;
    const double t_tile = 2.0 + nkb * (0.10 + 0.7 / st) + (bn / 64.0) * (act ? 3.5 : 1.6);
    const double t = waves * t_tile;
    if (t < best_t - 1e-9) { best_t = t; best = TileCfg{bn, st, 0, 0}; best_waves = waves; }
    if (forced) break;
  }
  // More than one wave of CTAs: persistent form, 2 CTAs per SM, each with a double-buffered accumulator (2 x <= 128 columns of
  // TMEM, so 4 x 128 per SM) and as deep an operand ring as fits next to the second CTA.
  // Measured on B200 (tools/check_persist.py): wins where the epilogue is light and the tile shape is kept (N <= 128 or a deep K
  // loop: 301056x64x64 40 -> 29 us, 301056x16x64 23 -> 13, 75264x128x512 28 -> 24); loses where BN would have to shrink from
  // 192 / 256 (more tiles re-reading A) and for the GELU epilogues, which are instruction-issue bound either way — those keep
  // the one-tile form.
  const int pm = allow_persist ? persist_mode() : 0;
  if (pm == 2 || (pm != 0 && best_waves > 1 && best.bn <= 128 && act == 0)) {
    int bn = best.bn > 128 ? 128 : best.bn;
    if (best.bn > 128 && N % 128 != 0) bn = (N % 96 == 0 && !w_kn) ? 96 : 64;
    if (!(w_kn && bn % 64)) {
      const size_t lim = smem_cap != 0 && smem_cap < 112 * 1024 ? smem_cap : 112 * 1024;
      int st = 1;
      while (st < kMaxStages && st < 2 * nkb && smem_bytes_persist(bn, st + 1) <= lim) ++st;
      const int64_t tiles = mt * ((N + bn - 1) / bn);
      if (smem_bytes_persist(bn, st) <= 113 * 1024 && tiles <= 0x7fffffff) {
        const int64_t slots = (int64_t)sms * 2;
        best = TileCfg{bn, st, 1, (int)(tiles < slots ? tiles : slots)};
      }
    }
  }
  // The one-CTA-per-SM persistent form with 16 epilogue warps (kEW = 16): EXPERIMENT, only with CSWIN_GEMM_PERSIST=3.  Measured on
  // B200 (profiles/r02_linear_persist_forms.log): 18816x1024x256 GELU 21.4 vs 20.7 us, 18816x768x256 13.9 vs 14.1, 75264x512x128 GELU
  // 39.6 vs 34.2, 301056x192x64 56.6 vs 42.7 — never ahead.  The epilogue warps are bound by their own dependent chains (0.4-0.5 IPC per
  // scheduler with 4 warps each, 96 registers per thread cap the SM at 20 warps), not by waiting for operands, so overlapping the
  // mainloop with the epilogue buys nothing once two CTAs share an SM.
  if (pm == 3) {
    int bn2 = 0;
    if (n16 <= 256) bn2 = n16;
    else { const int c2[] = {256, 192, 128}; for (int c : c2) if (N % c == 0) { bn2 = c; break; } }
    if (bn2 != 0 && w_kn && bn2 % 64) bn2 = 0;
    if (bn2 != 0) {
      const int64_t tiles2 = mt * ((N + bn2 - 1) / bn2);
      if ((tiles2 >= 2 * (int64_t)sms || (pm == 3 && tiles2 >= sms)) && tiles2 <= 0x7fffffff) {
        int st = 1;
        while (st < kMaxStages && st < 2 * nkb && smem_bytes_persist(bn2, st + 1, 16) <= 227 * 1024) ++st;
        if (smem_bytes_persist(bn2, st, 16) <= 227 * 1024) {
          best = TileCfg{bn2, st, 1, (int)(tiles2 < sms ? tiles2 : sms)};
          best.ew = 16;
        }
      }
    }
  }
  return best;
}

}  // namespace

int linear_tc_stats_parts(int64_t M, int N, int K, int act) {
  const TileCfg cfg = pick_tile(M, N, (K + BK - 1) / BK, act, sm_count(), false);
  return (N + cfg.bn - 1) / cfg.bn;
}

// geometry of an implicit-GEMM convolution over a (B, H, W, C) token image (cswin_conv_tokens_fwd); a->a = image base, a->lda = token stride
struct ConvGeom { int B, H, W, C, KH, KW, stride, pad, OH, OW; int64_t x_bs; };

static int linear_fwd_tc_impl(const cswin_linear_args_t* a, const ConvGeom* cv, cudaStream_t stream, bool* handled);

int linear_fwd_tc(const cswin_linear_args_t* a, cudaStream_t stream, bool* handled) { return linear_fwd_tc_impl(a, nullptr, stream, handled); }

// out (B OH OW, N) = conv(x (B, H W, C) token image, w (N, KH KW C) in (ky, kx, c) order) + bias: the Linear kernel with its A operand
// fetched as strided 4-D TMA boxes of the image (no column matrix).  *handled = false when the shape is outside the envelope.
int conv_tokens_fwd_tc(const void* x, int64_t x_bs, int64_t x_ts, const void* w, int64_t ldw, const void* bias, void* out, int64_t ldo,
                       int B, int H, int W, int C, int N, int KH, int KW, int stride, int pad, cudaStream_t stream, bool* handled) {
  *handled = false;
  ConvGeom g{B, H, W, C, KH, KW, stride, pad, (H + 2 * pad - KH) / stride + 1, (W + 2 * pad - KW) / stride + 1, x_bs};
  if (C % BK != 0 || g.OH <= 0 || g.OW <= 0 || g.OW > BM || stride < 1 || stride > 8 || g.OW * stride > 256) return CSWIN_OK;
  cswin_linear_args_t a = {};
  a.a = x; a.lda = x_ts; a.K1 = KH * KW * C; a.w = w; a.ldw = ldw; a.bias = bias; a.out = out; a.ldo = ldo;
  a.M = (int64_t)B * g.OH * g.OW; a.N = N;
  return linear_fwd_tc_impl(&a, &g, stream, handled);
}

static int linear_fwd_tc_impl(const cswin_linear_args_t* a, const ConvGeom* cv, cudaStream_t stream, bool* handled) {
  *handled = false;
  if (a->ln_gamma != nullptr) return CSWIN_OK;                               // LayerNorm prologue: SIMT kernel (host calls LN first on the bf16 path)
  const int K = a->K1 + a->K2;
  if (!aligned16(a->a) || !aligned16(a->w) || (a->lda * 2) % 16 || (a->ldw * 2) % 16) return CSWIN_OK;
  // K itself may be ragged (TMA zero-fills past the tensor extent); only the row pitches above must be 16-byte multiples
  if (a->a2 && (!aligned16(a->a2) || (a->lda2 * 2) % 16 || a->K1 % BK || a->K2 % 8)) return CSWIN_OK;
  if (a->M > 0x7fffffff) return CSWIN_OK;
  if (tc::encode_tiled_fn() == nullptr) return CSWIN_OK;

  GemmTcParams P;
  P.bias = (const __nv_bfloat16*)a->bias;
  P.res = (const __nv_bfloat16*)a->residual; P.ldr = a->ldr;
  P.sscale = a->sample_scale; P.rps = a->rows_per_sample > 0 ? a->rows_per_sample : 1;
  P.out = (__nv_bfloat16*)a->out; P.ldo = a->ldo;
  P.M = a->M; P.N = a->N; P.K1 = a->K1; P.K2 = a->K2; P.act = a->act;
  P.nkb1 = (a->K1 + BK - 1) / BK;
  P.nkb = P.nkb1 + (a->K2 + BK - 1) / BK;
  P.w_kn = a->w_layout;
  // (the training epilogues — aux output, act 2 — exist in the one-tile form only)
  TileCfg cfg = pick_tile(a->M, a->N, P.nkb, a->act, sm_count(), a->w_layout != 0, !(a->aux_out != nullptr || a->act == 2));
  P.conv = 0; P.cv_rows = BM; P.cv_tpi = 1; P.cv_nb = 1; P.cv_bh = 1; P.cv_stride = 1; P.cv_pad = 0; P.cv_KW = 1; P.cv_cblk = 1;
  if (cv != nullptr) {
    // M tile = whole images (OH OW nb <= 128) or bh full output rows of one image (bh | OH, bh OW <= 128)
    int nb = 1, bh = cv->OH;
    if (cv->OH * cv->OW <= BM) { nb = BM / (cv->OH * cv->OW); if (nb > cv->B) nb = cv->B; if (nb > 256) nb = 256; }
    else { bh = BM / cv->OW; while (bh > 1 && cv->OH % bh) --bh; }
    if (bh * cv->stride > 256) return CSWIN_OK;
    P.conv = 1; P.cv_nb = nb; P.cv_bh = bh; P.cv_rows = nb * bh * cv->OW; P.cv_tpi = cv->OH / bh;
    P.cv_stride = cv->stride; P.cv_pad = cv->pad; P.cv_KW = cv->KW; P.cv_cblk = cv->C / BK;
    cfg.persist = 0;
    // the grid is (M tiles of cv_rows rows) x (N tiles): re-pick BN for the real tile count
    cfg = pick_tile(((a->M + P.cv_rows - 1) / P.cv_rows) * BM, a->N, P.nkb, a->act, sm_count(), false);
    cfg.persist = 0;
  }
  P.BN = cfg.bn;
  P.stages = cfg.stages;
  P.tmem_cols = tmem_cols_for(P.BN);
  P.nt_n = (a->N + P.BN - 1) / P.BN;
  P.ntiles = (int)(((a->M + BM - 1) / BM) * P.nt_n);
  P.acc_stride = P.tmem_cols;
  P.bias_vec = a->bias != nullptr && aligned16(a->bias);
  P.ln_stats = a->ln_stats; P.ln_parts = a->ln_stats_parts; P.ln_invC = a->ln_C > 0 ? 1.0f / (float)a->ln_C : 0.f;
  P.ln_eps = a->ln_eps; P.ln_cs = a->ln_colsum; P.bias_f32 = a->bias_f32; P.stats_out = a->stats_out;
  P.trace = g_trace.load(std::memory_order_relaxed);
  P.vec_ok = aligned16(a->out) && (a->ldo * 2) % 16 == 0 &&
             (a->residual == nullptr || (aligned16(a->residual) && (a->ldr * 2) % 16 == 0));
  P.tma_out = P.vec_ok && a->N % 8 == 0 && !no_tma_out() && P.cv_rows == BM;     // partial conv tiles: predicated stores
  if (P.tma_out) {
    const uint64_t dims[2] = {(uint64_t)a->N, (uint64_t)a->M}, str[1] = {(uint64_t)a->ldo * 2};
    const uint32_t box[2] = {32, 32};
    if (!tc::make_tensor_map_bf16(&P.map_out, a->out, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B)) return CSWIN_ERR_CUDA;
  }
  if (a->stats_out != nullptr && !P.vec_ok) { set_error("linear_fwd: stats_out needs 16-byte aligned output rows"); return CSWIN_ERR_UNSUPPORTED; }

  if (cv != nullptr) {          // (channel, x, y, image) view of the token image, traversed with the conv stride along x and y
    const uint64_t dims[4] = {(uint64_t)cv->C, (uint64_t)cv->W, (uint64_t)cv->H, (uint64_t)cv->B};
    const uint64_t str[3] = {(uint64_t)a->lda * 2, (uint64_t)a->lda * 2 * cv->W, (uint64_t)cv->x_bs * 2};
    const uint32_t s = (uint32_t)cv->stride;
    const uint32_t box[4] = {BK, (uint32_t)cv->OW * s, (uint32_t)P.cv_bh * s, (uint32_t)P.cv_nb};
    const uint32_t es[4] = {1, s, s, 1};
    if ((cv->x_bs * 2) % 16) return CSWIN_OK;
    if (!tc::make_tensor_map_bf16(&P.map_a, a->a, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, es)) return CSWIN_ERR_CUDA;
  } else {
    const uint64_t dims[2] = {(uint64_t)a->K1, (uint64_t)a->M};
    const uint64_t str[1] = {(uint64_t)a->lda * 2};
    const uint32_t box[2] = {BK, BM};
    if (!tc::make_tensor_map_bf16(&P.map_a, a->a, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return CSWIN_ERR_CUDA;
  }
  if (a->a2) {
    const uint64_t dims[2] = {(uint64_t)a->K2, (uint64_t)a->M};
    const uint64_t str[1] = {(uint64_t)a->lda2 * 2};
    const uint32_t box[2] = {BK, BM};
    if (!tc::make_tensor_map_bf16(&P.map_a2, a->a2, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return CSWIN_ERR_CUDA;
  } else {
    P.map_a2 = P.map_a;
  }
  {
    const uint64_t dims_nk[2] = {(uint64_t)K, (uint64_t)a->N}, dims_kn[2] = {(uint64_t)a->N, (uint64_t)K};
    const uint64_t str[1] = {(uint64_t)a->ldw * 2};
    const uint32_t box_nk[2] = {BK, (uint32_t)P.BN}, box_kn[2] = {64, BK};
    if (!tc::make_tensor_map_bf16(&P.map_w, a->w, 2, a->w_layout ? dims_kn : dims_nk, str, a->w_layout ? box_kn : box_nk,
                                  CU_TENSOR_MAP_SWIZZLE_128B)) return CSWIN_ERR_CUDA;
  }

  // training epilogues (aux pre-activation output, act 2 = multiply by GELU'(residual)): TMA-store fast path, plain (unfolded) form
  const bool train = a->aux_out != nullptr || a->act == 2;
  P.aux = 0; P.aux_off = 0;
  size_t smem = smem_bytes(P.BN, P.stages);
  if (train) {
    if (!P.tma_out || a->ln_stats != nullptr || a->stats_out != nullptr) return CSWIN_OK;       // caller reports "unsupported"
    if (a->act == 2 && a->residual == nullptr) return CSWIN_OK;
    if (a->aux_out != nullptr) {
      if (!aligned16(a->aux_out) || (a->ld_aux * 2) % 16) return CSWIN_OK;
      const uint64_t dims[2] = {(uint64_t)a->N, (uint64_t)a->M}, str[1] = {(uint64_t)a->ld_aux * 2};
      const uint32_t box[2] = {32, 32};
      if (!tc::make_tensor_map_bf16(&P.map_aux, a->aux_out, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B)) return CSWIN_ERR_CUDA;
      P.aux = 1;
      P.aux_off = (int)((smem - 1024 + 1023) / 1024 * 1024);   // second 16 KB staging area behind everything else (the main one
      smem = 1024 + (size_t)P.aux_off + 8 * 2048;               // aliases the ring); 1 KB aligned: the swizzle is address-based
    }
  }
  using Kern = void (*)(const GemmTcParams);
  static const Kern kerns[13] = {linear_tc_kernel<false, false>, linear_tc_kernel<true, false>, linear_tc_kernel<false, true>,
                                 linear_tc_kernel<true, true>, linear_tc_kernel<false, false, true>,
                                 linear_tc_kernel<false, false, false, true>, linear_tc_kernel<true, false, false, true>,
                                 linear_tc_kernel<false, true, false, true>, linear_tc_kernel<true, true, false, true>,
                                 linear_tc_kernel<false, false, false, true, 16>, linear_tc_kernel<true, false, false, true, 16>,
                                 linear_tc_kernel<false, true, false, true, 16>, linear_tc_kernel<true, true, false, true, 16>};
  static std::atomic<int> configured{0};
  if (!configured.load(std::memory_order_acquire)) {
    for (Kern k : kerns) CSWIN_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured.store(1, std::memory_order_release);
  }
  dim3 grid((unsigned)((a->M + P.cv_rows - 1) / P.cv_rows), (unsigned)((a->N + P.BN - 1) / P.BN));
  const int variant = (a->ln_stats != nullptr ? 1 : 0) | (a->stats_out != nullptr ? 2 : 0);
  unsigned threads = kThreads;
  Kern kern = train ? kerns[4] : kerns[variant];
  if (cfg.persist && !train && P.tma_out) {             // persistent form: 1-D grid of resident CTAs, two accumulator buffers
    kern = kerns[(cfg.ew == 16 ? 9 : 5) + variant];
    grid = dim3((unsigned)cfg.grid);
    P.tmem_cols = 2 * P.acc_stride;
    smem = smem_bytes_persist(P.BN, P.stages, cfg.ew);
    threads = 64 + 32 * cfg.ew;
  }
  CSWIN_CUDA_OK(launch_pdl(kern, grid, dim3(threads), smem, stream, P));
  CSWIN_LAUNCH_CHECK();
  g_tc_launches.fetch_add(1, std::memory_order_relaxed);
  *handled = true;
  return CSWIN_OK;
}

}  // namespace cswin
