// attention_bwd_tc.cu — backward of the fused LePE stripe attention, bf16, tcgen05 / TMEM / TMA (sm_100a).
//
// Autograd of LePEAttention.forward (networks/cswin_unet.py:82-109; formulas: SURVEY.md Appendix A), both branches in
// one launch.  Same tiling as the forward (attention_tc.cu): 128 query rows = 128 TMEM lanes, 256 threads (two per
// row); windows of N <= 64 tokens are packed two per tile.  Per tile, five tensor-core contractions with every
// operand taken from shared memory in the layout it already has:
//     S  = Q K^T, dP = G V^T                 (K-major operands, TMA-delivered 64-byte-swizzled rows)      -> TMEM
//     P  = exp(scale S - lse)  (lse saved by the forward), delta = rowsum(P o dP), dS = P o (dP - delta)   (threads)
//     P and dS are written once, as bf16 [q][kv] rows with the 128-byte swizzle, and then serve as
//       dV = P^T  G   (A MN-major = P read "transposed" for free, B = G MN-major)
//       dK = dS^T Q   (A MN-major,                          B = Q MN-major)     (x scale in the epilogue)
//       dQ = dS   K   (A K-major,                           B = K MN-major)     (x scale in the epilogue)
//   + LePE: dV += depthwise-conv-transpose(G) (window-local zero padding), d w / d b reduced per tile -> fp32 atomics.
// Rows / keys beyond the window and the off-diagonal blocks of a packed tile are exact zeros in P and dS, and the
// padded rows of Q, K, V, G are zero-filled, so they contribute nothing.
// The grid is PERSISTENT (at most 2 CTAs per SM, each walking tiles blockIdx.x, +gridDim.x, ...): barriers and TMEM are
// set up once per CTA, the LePE weights are re-staged only when the head of the CTA's tiles changes.
//
// d get_v.weight / d get_v.bias do NOT come from this kernel: per tile they are a reduction over tokens per channel —
// (channel, token-group) threads, 2-byte swizzled smem reads, then 320 same-address fp32 atomics per (window, head) —
// which cost 6 us of every tile's critical path and, at stage 1 / batch 24, 860 k atomics on 640 addresses (an L2 sector
// retires one atomic per ~14 ns: 194 us per launch).  lepe_param_grad_kernel below computes them instead straight from
// the (B, H, W, C) tensors: a CTA owns 16 channels x a token range of one image, a thread owns 4 channels and walks
// tokens with 8-byte loads keeping the 9 tap sums + the bias sum in registers, so the reduction over tokens costs
// nothing until the end (shuffles, shared-memory atomics, 40 red.global.add.v4.f32 per CTA).
#include "common.cuh"
#include "tc_common.cuh"

namespace cswin {
namespace {

using namespace tc;

struct BwdBranch {
  __nv_bfloat16 *dq, *dk, *dv;
  const __nv_bfloat16* cw;
  const float* lse;
  int64_t dq_bs, dq_ts, dk_bs, dk_ts, dv_bs, dv_ts;
  int heads, hs, ws, nww, nwin, N, tile_begin, nprob;
};
struct alignas(64) BwdParams {
  CUtensorMap map[2][4];      // [branch][q, k, v, g]
  BwdBranch br[2];
  int nb, reso, total_tiles;
  float scale, scale_log2e;
  unsigned long long* trace;
};

constexpr int kRows = 128, kThr = 256, kRowB = 64;
constexpr int kOpB = kRows * kRowB;                 // 8 KB per [128][32] bf16 operand
constexpr int kPB = 2 * kRows * 128;                // 32 KB per [128][128] bf16 matrix (two 64-column blocks)
constexpr uint32_t kTmem = 256;
constexpr int kSmem = 4 * kOpB + 2 * kPB + 2 * 9 * 32 * 4 + 2 * 128 * 4 + 64 + 1024;

__device__ __forceinline__ uint32_t sw64(uint32_t base, int row, int chunk) {       // [row][4 chunks of 16 B]
  return base + row * 64 + (((chunk ^ (row >> 1)) & 3) << 4);
}
__device__ __forceinline__ uint32_t sw128(uint32_t base, int row, int chunk) {      // [row][8 chunks of 16 B]
  return base + row * 128 + (((chunk ^ row) & 7) << 4);
}
__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&o)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]), "=r"(o[8]),
        "=r"(o[9]), "=r"(o[10]), "=r"(o[11]), "=r"(o[12]), "=r"(o[13]), "=r"(o[14]), "=r"(o[15])
      : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(kThr, 2) lepe_attn_bwd_tc_kernel(const __grid_constant__ BwdParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* Qs = smem;
  uint8_t* Ks = Qs + kOpB;
  uint8_t* Vs = Ks + kOpB;
  uint8_t* Gs = Vs + kOpB;
  uint8_t* Ps = Gs + kOpB;                 // [2 column blocks][128 rows][128 B]
  uint8_t* Ds = Ps + kPB;
  float* Wt = reinterpret_cast<float*>(Ds + kPB);       // [2][9][32]
  float* Xd = Wt + 2 * 9 * 32;                       // [2][128]      delta exchange
  uint64_t* bars = reinterpret_cast<uint64_t*>(Xd + 2 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int tid = threadIdx.x, warp = tid >> 5;
  pdl_trigger();
  if (tid == 0) trace_stamp(P.trace, 0);
  const int row = tid & 127, half = tid >> 7;
  const uint32_t bar_tma = smem_u32(&bars[0]), bar_s = smem_u32(&bars[1]), bar_o = smem_u32(&bars[2]);
  if (warp == 0) { tmem_alloc(smem_u32(tmem_slot), kTmem); tmem_relinquish(); }
  if (tid == 32) { mbar_init(bar_tma, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1); fence_barrier_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t gsb = smem_u32(Gs), vsb = smem_u32(Vs), ps = smem_u32(Ps), dsb = smem_u32(Ds);

  int acc_bi = -1, acc_h0 = 0;                           // (branch, first head) whose LePE weights are staged in Wt

  uint32_t ph = 0;                                       // mbarrier phase parity of this iteration
  for (int gt = blockIdx.x; gt < P.total_tiles; gt += gridDim.x, ph ^= 1) {
    const int bi = (P.nb > 1 && gt >= P.br[1].tile_begin) ? 1 : 0;
    const BwdBranch& br = P.br[bi];
    const int tile = gt - br.tile_begin;
    const int N = br.N, hs = br.hs, ws = br.ws;
    const int slots = (N <= 64) ? 2 : 1;
    const int slot_rows = kRows / slots;
    const int p0 = tile * slots;
    const int np = min(slots, br.nprob - p0);
    const int kext = (slots == 2) ? 128 : ((N + 15) & ~15);
    const int slot = row / slot_rows;
    const int n = row - slot * slot_rows;
    const bool valid = slot < np && n < N;

    int mb, mih, miw, mhead;
    {
      int local = p0 + min(slot, np - 1);
      mhead = local % br.heads; local /= br.heads;
      const int win = local % br.nwin;
      mb = local / br.nwin;
      mih = win / br.nww; miw = win - mih * br.nww;
    }
    const int r_ = n / ws, c_ = n - r_ * ws;
    const int64_t tok = (int64_t)(mih * hs + r_) * P.reso + (miw * ws + c_);

    // ---- per-tile prologue (the previous tile's smem / TMEM reads ended at its closing barrier) ----
    const int h0 = p0 % br.heads;
    const bool rekey = (bi != acc_bi) || (h0 != acc_h0);  // CTA-uniform
    if (rekey) {
      if (tid < slots * 36) {                               // LePE weights of the tile's head(s): Wt[slot][tap][ch]
        const int s = tid / 36, i = tid - s * 36;
        const int hd = (h0 + s) % br.heads;
        const uint4 raw = *reinterpret_cast<const uint4*>(br.cw + (size_t)hd * 288 + i * 8);
        const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int idx = i * 8 + e, ch = idx / 9, t = idx - ch * 9;
          Wt[(s * 9 + t) * 32 + ch] = (e & 1) ? bf16_hi(w4[e >> 1]) : bf16_lo(w4[e >> 1]);
        }
      }
      acc_bi = bi; acc_h0 = h0;
    }
    // zero the rows of Q, K, V, G that TMA does not write (they are contraction rows of dQ / dK / dV)
    for (int i = tid; i < 4 * kRows * 4; i += kThr) {
      const int r = (i >> 2) & 127;
      const int s = r / slot_rows, rn = r - s * slot_rows;
      if (s >= np || rn >= N) *reinterpret_cast<uint4*>(smem + i * 16) = make_uint4(0, 0, 0, 0);
    }
    const float lse = valid ? br.lse[((int64_t)mb * P.reso * P.reso + tok) * br.heads + mhead] : 0.f;
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) trace_stamp(P.trace, 1);
    pdl_wait();                                          // d out (previous kernel's output) and the gradient buffers are safe from here

    if (warp == 0 && elect_one()) {     // one elected lane, warp-uniform datapath for the TMA / tcgen05 issue
      mbar_expect_tx(bar_tma, (uint32_t)(np * 4 * N * kRowB));
      for (int s = 0; s < np; ++s) {
        int local = p0 + s;
        const int hd = local % br.heads; local /= br.heads;
        const int win = local % br.nwin;
        const int b = local / br.nwin;
        const int ih = win / br.nww, iw = win - ih * br.nww;
        const uint32_t off = s * slot_rows * kRowB;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          tma_load_4d(smem_u32(smem + j * kOpB) + off, &P.map[bi][j], bar_tma, hd * 32, iw * ws, ih * hs, b);
      }
      mbar_wait(bar_tma, ph);
      tc_fence_after();
      const uint32_t idesc = make_idesc_bf16(128, kext, 0, 0);
      const uint64_t qd = make_smem_desc(smem_u32(Qs), 16, 512, kLayoutSw64), kd = make_smem_desc(smem_u32(Ks), 16, 512, kLayoutSw64);
      const uint64_t gd = make_smem_desc(smem_u32(Gs), 16, 512, kLayoutSw64), vd = make_smem_desc(smem_u32(Vs), 16, 512, kLayoutSw64);
      mma_ss(tmem_base, qd, kd, idesc, false);
      mma_ss(tmem_base, qd + 2, kd + 2, idesc, true);
      mma_ss(tmem_base + 128, gd, vd, idesc, false);
      mma_ss(tmem_base + 128, gd + 2, vd + 2, idesc, true);
      tc_commit(bar_s);
    }
    mbar_wait(bar_s, ph);
    if (tid == 0) trace_stamp(P.trace, 2);
    tc_fence_after();

    // ---- P, delta, dS for row `row`, key columns [kbeg, kbeg + hcols) of my slot ----
    const int hcols = slot_rows >> 1;                     // 32 or 64
    const int kbeg = half * hcols;
    const int cbeg = slot * slot_rows + kbeg;             // tile column of my first key
    const int nch = hcols >> 5;
    const float lse2 = lse * 1.4426950408889634f;
    float dsum = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (c < nch && kbeg + 32 * c < kext) {
        uint32_t s[32], d[32];
        tmem_ld32(trow + cbeg + 32 * c, s);
        tmem_ld32(trow + 128 + cbeg + 32 * c, d);
        tmem_wait_ld();
        const int lim = valid ? N - (kbeg + 32 * c) : 0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < lim) dsum = fmaf(ex2a(fmaf(__uint_as_float(s[j]), P.scale_log2e, -lse2)), __uint_as_float(d[j]), dsum);
      }
    }
    Xd[half * 128 + row] = dsum;
    __syncthreads();
    const float delta = dsum + Xd[(half ^ 1) * 128 + row];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (c < nch) {
        const int col = cbeg + 32 * c;                    // tile column of this chunk (multiple of 32)
        uint32_t pp[16], dd[16];
        if (kbeg + 32 * c < kext) {
          uint32_t s[32], d[32];
          tmem_ld32(trow + col, s);
          tmem_ld32(trow + 128 + col, d);
          tmem_wait_ld();
          const int lim = valid ? N - (kbeg + 32 * c) : 0;
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float p0_ = 0.f, p1_ = 0.f, d0 = 0.f, d1 = 0.f;
            if (j < lim) { p0_ = ex2a(fmaf(__uint_as_float(s[j]), P.scale_log2e, -lse2)); d0 = p0_ * (__uint_as_float(d[j]) - delta); }
            if (j + 1 < lim) { p1_ = ex2a(fmaf(__uint_as_float(s[j + 1]), P.scale_log2e, -lse2)); d1 = p1_ * (__uint_as_float(d[j + 1]) - delta); }
            pp[j >> 1] = pack_bf16x2(p0_, p1_);
            dd[j >> 1] = pack_bf16x2(d0, d1);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) { pp[j] = 0u; dd[j] = 0u; }
        }
        const int blk = col >> 6, ch0 = (col & 63) >> 3;  // 64-column block, first 16-byte chunk inside the 128-byte row
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          sts128(sw128(ps + blk * 16384, row, ch0 + k), pp[4 * k], pp[4 * k + 1], pp[4 * k + 2], pp[4 * k + 3]);
          sts128(sw128(dsb + blk * 16384, row, ch0 + k), dd[4 * k], dd[4 * k + 1], dd[4 * k + 2], dd[4 * k + 3]);
        }
      }
    }
    if (slots == 2) {                                     // the other problem's key columns of my row: zeros
      const int col = (1 - slot) * 64 + half * 32;
      const int blk = col >> 6, ch0 = (col & 63) >> 3;
#pragma unroll
      for (int k = 0; k < 4; ++k) { sts128(sw128(ps + blk * 16384, row, ch0 + k), 0, 0, 0, 0); sts128(sw128(dsb + blk * 16384, row, ch0 + k), 0, 0, 0, 0); }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();                                      // S / dP fully consumed; P / dS visible to the tensor core

    if (warp == 0 && elect_one()) {     // one elected lane, warp-uniform datapath for the TMA / tcgen05 issue
      trace_stamp(P.trace, 3);
      tc_fence_after();
      const uint32_t id_mn = make_idesc_bf16(128, 32, 1, 1);      // A = P / dS read MN-major (kv rows out), B MN-major
      const uint32_t id_k = make_idesc_bf16(128, 32, 0, 1);       // A = dS K-major (q rows out), B = K MN-major
      const uint64_t pd = make_smem_desc(ps, 16384, 1024, kLayoutSw128), dd = make_smem_desc(dsb, 16384, 1024, kLayoutSw128);
      const uint64_t gd = make_smem_desc(smem_u32(Gs), 512, 512, kLayoutSw64), qd = make_smem_desc(smem_u32(Qs), 512, 512, kLayoutSw64);
      const uint64_t kd = make_smem_desc(smem_u32(Ks), 512, 512, kLayoutSw64);
      for (int k = 0; k < 8; ++k) {                               // contraction over the 128 query rows, 16 per MMA
        mma_ss(tmem_base + 64, pd + (uint64_t)k * (2048 >> 4), gd + (uint64_t)k * (1024 >> 4), id_mn, k > 0);   // dV
        mma_ss(tmem_base + 32, dd + (uint64_t)k * (2048 >> 4), qd + (uint64_t)k * (1024 >> 4), id_mn, k > 0);   // dK
      }
      const uint64_t dk_ = make_smem_desc(dsb, 16, 1024, kLayoutSw128);                 // dS as K-major A
      for (int k = 0; k < kext / 16; ++k) {                       // contraction over the keys, 16 per MMA
        const uint64_t a = dk_ + (uint64_t)((k >> 2) * (16384 >> 4) + (k & 3) * 2);
        mma_ss(tmem_base, a, kd + (uint64_t)k * (1024 >> 4), id_k, k > 0);                                      // dQ
      }
      tc_commit(bar_o);
    }

    // ---- LePE: conv-transpose of G for my token / my 16 channels, overlapped with the MMAs ----
    mbar_wait(bar_tma, ph);
    float lv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) lv[j] = 0.f;
    if (valid) {
      const float* wt = Wt + slot * 9 * 32 + half * 16;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int rr = r_ - (t / 3 - 1), cc = c_ - (t % 3 - 1);       // the output position that read me through tap t
        if (rr >= 0 && rr < hs && cc >= 0 && cc < ws) {
          const int gr = slot * slot_rows + rr * ws + cc;
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            const uint4 g4 = lds128(sw64(gsb, gr, half * 2 + ch));
            const float4 w0 = *reinterpret_cast<const float4*>(wt + t * 32 + ch * 8);
            const float4 w1 = *reinterpret_cast<const float4*>(wt + t * 32 + ch * 8 + 4);
            lv[ch * 8 + 0] = fmaf(w0.x, bf16_lo(g4.x), lv[ch * 8 + 0]); lv[ch * 8 + 1] = fmaf(w0.y, bf16_hi(g4.x), lv[ch * 8 + 1]);
            lv[ch * 8 + 2] = fmaf(w0.z, bf16_lo(g4.y), lv[ch * 8 + 2]); lv[ch * 8 + 3] = fmaf(w0.w, bf16_hi(g4.y), lv[ch * 8 + 3]);
            lv[ch * 8 + 4] = fmaf(w1.x, bf16_lo(g4.z), lv[ch * 8 + 4]); lv[ch * 8 + 5] = fmaf(w1.y, bf16_hi(g4.z), lv[ch * 8 + 5]);
            lv[ch * 8 + 6] = fmaf(w1.z, bf16_lo(g4.w), lv[ch * 8 + 6]); lv[ch * 8 + 7] = fmaf(w1.w, bf16_hi(g4.w), lv[ch * 8 + 7]);
          }
        }
      }
    }
    if (tid == 0) trace_stamp(P.trace, 4);
    if (tid == 0) trace_stamp(P.trace, 5);
    mbar_wait(bar_o, ph);
    if (tid == 0) trace_stamp(P.trace, 6);
    tc_fence_after();
    {
      uint32_t q16[16], k16[16], v16[16];
      tmem_ld16(trow + half * 16, q16);
      tmem_ld16(trow + 32 + half * 16, k16);
      tmem_ld16(trow + 64 + half * 16, v16);
      tmem_wait_ld();
      if (valid) {
        const int64_t cho = mhead * 32 + half * 16;
        __nv_bfloat16* pq = br.dq + (int64_t)mb * br.dq_bs + tok * br.dq_ts + cho;
        __nv_bfloat16* pk = br.dk + (int64_t)mb * br.dk_bs + tok * br.dk_ts + cho;
        __nv_bfloat16* pv = br.dv + (int64_t)mb * br.dv_bs + tok * br.dv_ts + cho;
        uint32_t a[8], b[8], c[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a[j] = pack_bf16x2(__uint_as_float(q16[2 * j]) * P.scale, __uint_as_float(q16[2 * j + 1]) * P.scale);
          b[j] = pack_bf16x2(__uint_as_float(k16[2 * j]) * P.scale, __uint_as_float(k16[2 * j + 1]) * P.scale);
          c[j] = pack_bf16x2(__uint_as_float(v16[2 * j]) + lv[2 * j], __uint_as_float(v16[2 * j + 1]) + lv[2 * j + 1]);
        }
        *reinterpret_cast<uint4*>(pq) = make_uint4(a[0], a[1], a[2], a[3]); *reinterpret_cast<uint4*>(pq + 8) = make_uint4(a[4], a[5], a[6], a[7]);
        *reinterpret_cast<uint4*>(pk) = make_uint4(b[0], b[1], b[2], b[3]); *reinterpret_cast<uint4*>(pk + 8) = make_uint4(b[4], b[5], b[6], b[7]);
        *reinterpret_cast<uint4*>(pv) = make_uint4(c[0], c[1], c[2], c[3]); *reinterpret_cast<uint4*>(pv + 8) = make_uint4(c[4], c[5], c[6], c[7]);
      }
    }
    tc_fence_before();
    __syncthreads();                                      // closes the tile: smem operands and TMEM are free again
    if (tid == 0) trace_stamp(P.trace, 7);
  }
  if (warp == 0) tmem_dealloc(tmem_base, kTmem);
}

// ------------------------------------------------------------------------------------------------------------------
// Wide variant: windows of 128 < N <= 256 tokens (the 512^2 configuration's 32x8 / 8x32 / 16x16 stripes, 14x14 windows ...).
// One CTA per (window, head) problem (persistent over problems), 1 CTA per SM: Q, K, V, G of the whole window in shared memory
// (4 x 16 KB, one TMA box each), processed as 2 query tiles x 2 key halves of 128 x 128:
//   pass A (per query tile): S = Q K^T and dP = G V^T per key half into TMEM columns [0,128) / [128,256), the row's
//           delta = sum_j P_ij dP_ij accumulated over both halves (the row's lse comes from the forward);
//   pass B: S, dP again per key half, P / dS (bf16, 128-byte swizzle) into shared memory, then
//           dV[half] += P^T G_tile, dK[half] += dS^T Q_tile (accumulated over the two query tiles in TMEM columns
//           [352,416) / [288,352)), dQ_tile += dS K_half (columns [256,288), stored after the second half).
// S and dP are recomputed instead of kept: at 128 x 256 fp32 each they would fill all 512 TMEM columns on their own.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kWOpB = 256 * kRowB;                  // 16 KB per [256][32] bf16 operand
constexpr int kSmemWideBwd = 4 * kWOpB + 2 * kPB + 9 * 32 * 4 + 2 * 128 * 4 + 64 + 1024;
constexpr uint32_t kColDQ = 256, kColDK = 288, kColDV = 352;

__global__ void __launch_bounds__(kThr, 1) lepe_attn_bwd_wide_tc_kernel(const __grid_constant__ BwdParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* Qs = smem;
  uint8_t* Ks = Qs + kWOpB;
  uint8_t* Vs = Ks + kWOpB;
  uint8_t* Gs = Vs + kWOpB;
  uint8_t* Ps = Gs + kWOpB;                 // [2 column blocks][128 rows][128 B]
  uint8_t* Ds = Ps + kPB;
  float* Wt = reinterpret_cast<float*>(Ds + kPB);       // [9][32]
  float* Xd = Wt + 9 * 32;                              // [2][128]      delta exchange
  uint64_t* bars = reinterpret_cast<uint64_t*>(Xd + 2 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & 127, half = tid >> 7;
  const uint32_t bar_tma = smem_u32(&bars[0]), bar_s = smem_u32(&bars[1]), bar_o = smem_u32(&bars[2]);
  if (warp == 0) { tmem_alloc(smem_u32(tmem_slot), 512u); tmem_relinquish(); }
  if (tid == 32) { mbar_init(bar_tma, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1); fence_barrier_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();                                         // only now: this CTA owns all 512 TMEM columns of its SM, a dependent CTA
                                                         // that got them first would wait for this grid while this CTA waits for TMEM
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t qsb = smem_u32(Qs), ksb = smem_u32(Ks), vsb = smem_u32(Vs), gsb = smem_u32(Gs), ps = smem_u32(Ps), dsb = smem_u32(Ds);

  uint32_t ph_tma = 0, ph_s = 0, ph_o = 0;               // mbarrier phase parities
  int w_bi = -1, w_hd = -1;                              // (branch, head) whose LePE weights are staged
  for (int gt = blockIdx.x; gt < P.total_tiles; gt += gridDim.x) {
    const int bi = (P.nb > 1 && gt >= P.br[1].tile_begin) ? 1 : 0;
    const BwdBranch& br = P.br[bi];
    const int N = br.N, hs = br.hs, ws = br.ws;
    const int kext = (N + 15) & ~15;
    int local = gt - br.tile_begin;
    const int mhead = local % br.heads; local /= br.heads;
    const int win = local % br.nwin;
    const int mb = local / br.nwin;
    const int mih = win / br.nww, miw = win - mih * br.nww;

    if (bi != w_bi || mhead != w_hd) {                   // LePE weights of this head: Wt[tap][ch]   (CTA-uniform branch)
      if (tid < 36) {
        const uint4 raw = *reinterpret_cast<const uint4*>(br.cw + (size_t)mhead * 288 + tid * 8);
        const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int idx = tid * 8 + e, ch = idx / 9, t = idx - ch * 9;
          Wt[t * 32 + ch] = (e & 1) ? bf16_hi(w4[e >> 1]) : bf16_lo(w4[e >> 1]);
        }
      }
      w_bi = bi; w_hd = mhead;
    }
    // rows [N, 256) of Q, K, V, G are contraction / output rows of the MMAs but not written by TMA: zero them
    for (int i = tid; i < 4 * 256 * 4; i += kThr) {
      const int r = (i >> 2) & 255;
      if (r >= N) *reinterpret_cast<uint4*>(smem + i * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
    __syncthreads();
    pdl_wait();
    if (warp == 0 && elect_one()) {      // one elected lane, warp-uniform datapath for the TMA / tcgen05 issue
      mbar_expect_tx(bar_tma, (uint32_t)(4 * N * kRowB));
#pragma unroll
      for (int j = 0; j < 4; ++j)
        tma_load_4d(smem_u32(smem + j * kWOpB), &P.map[bi][j], bar_tma, mhead * 32, miw * ws, mih * hs, mb);
    }
    mbar_wait(bar_tma, ph_tma); ph_tma ^= 1;
    tc_fence_after();

    bool pending_o = false;                              // a dV / dK / dQ commit whose completion nobody has waited for yet
#pragma unroll 1
    for (int qt = 0; qt < 2; ++qt) {
      const int qn = qt * 128 + row;                     // window-local token of my query row
      const bool qvalid = qn < N;
      const int qr = qn / ws, qc = qn - qr * ws;
      const int64_t qtok = (int64_t)(mih * hs + qr) * P.reso + (miw * ws + qc);
      const float lse2 = (qvalid ? br.lse[((int64_t)mb * P.reso * P.reso + qtok) * br.heads + mhead] : 0.f) * 1.4426950408889634f;
      float delta = 0.f;
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {             // 0: delta, 1: P / dS and the three gradient contractions
#pragma unroll 1
        for (int kh = 0; kh < 2; ++kh) {
          const int kx = min(128, kext - 128 * kh);      // key columns of this half fed to the MMAs (multiple of 16, >= 16)
          if (warp == 0 && elect_one()) {      // one elected lane, warp-uniform datapath for the TMA / tcgen05 issue
            tc_fence_after();
            const uint32_t idesc = make_idesc_bf16(128, kx, 0, 0);
            const uint64_t qd = make_smem_desc(qsb + qt * 8192, 16, 512, kLayoutSw64), kd = make_smem_desc(ksb + kh * 8192, 16, 512, kLayoutSw64);
            const uint64_t gd = make_smem_desc(gsb + qt * 8192, 16, 512, kLayoutSw64), vd = make_smem_desc(vsb + kh * 8192, 16, 512, kLayoutSw64);
            mma_ss(tmem_base, qd, kd, idesc, false);
            mma_ss(tmem_base, qd + 2, kd + 2, idesc, true);
            mma_ss(tmem_base + 128, gd, vd, idesc, false);
            mma_ss(tmem_base + 128, gd + 2, vd + 2, idesc, true);
            tc_commit(bar_s);
          }
          mbar_wait(bar_s, ph_s); ph_s ^= 1;
          tc_fence_after();
          // my 64 key columns of this half: [half * 64, +64), two chunks of 32
          if (pass == 0) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int k0 = half * 64 + 32 * c;
              if (k0 < kx) {
                uint32_t s[32], d[32];
                tmem_ld32(trow + k0, s);
                tmem_ld32(trow + 128 + k0, d);
                tmem_wait_ld();
                const int lim = qvalid ? N - (kh * 128 + k0) : 0;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j < lim) delta = fmaf(ex2a(fmaf(__uint_as_float(s[j]), P.scale_log2e, -lse2)), __uint_as_float(d[j]), delta);
              }
            }
            tc_fence_before();
            __syncthreads();                             // S / dP consumed: the next MMA pair may overwrite them
          } else {
            if (pending_o) { mbar_wait(bar_o, ph_o); ph_o ^= 1; pending_o = false; }   // P / dS of the previous half have been read
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int k0 = half * 64 + 32 * c;         // column inside the 128-column tile (multiple of 32)
              uint32_t pp[16], dd[16];
              if (k0 < kx) {
                uint32_t s[32], d[32];
                tmem_ld32(trow + k0, s);
                tmem_ld32(trow + 128 + k0, d);
                tmem_wait_ld();
                const int lim = qvalid ? N - (kh * 128 + k0) : 0;
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                  float p0_ = 0.f, p1_ = 0.f, d0 = 0.f, d1 = 0.f;
                  if (j < lim) { p0_ = ex2a(fmaf(__uint_as_float(s[j]), P.scale_log2e, -lse2)); d0 = p0_ * (__uint_as_float(d[j]) - delta); }
                  if (j + 1 < lim) { p1_ = ex2a(fmaf(__uint_as_float(s[j + 1]), P.scale_log2e, -lse2)); d1 = p1_ * (__uint_as_float(d[j + 1]) - delta); }
                  pp[j >> 1] = pack_bf16x2(p0_, p1_);
                  dd[j >> 1] = pack_bf16x2(d0, d1);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) { pp[j] = 0u; dd[j] = 0u; }
              }
              const int blk = k0 >> 6, ch0 = (k0 & 63) >> 3;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                sts128(sw128(ps + blk * 16384, row, ch0 + k), pp[4 * k], pp[4 * k + 1], pp[4 * k + 2], pp[4 * k + 3]);
                sts128(sw128(dsb + blk * 16384, row, ch0 + k), dd[4 * k], dd[4 * k + 1], dd[4 * k + 2], dd[4 * k + 3]);
              }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();                             // S / dP consumed; P / dS visible to the tensor core
            if (warp == 0 && elect_one()) {      // one elected lane, warp-uniform datapath for the TMA / tcgen05 issue
              tc_fence_after();
              const uint32_t id_mn = make_idesc_bf16(128, 32, 1, 1);      // A = P / dS read MN-major (kv rows out), B MN-major
              const uint32_t id_k = make_idesc_bf16(128, 32, 0, 1);       // A = dS K-major (q rows out), B = K MN-major
              const uint64_t pd = make_smem_desc(ps, 16384, 1024, kLayoutSw128), dd = make_smem_desc(dsb, 16384, 1024, kLayoutSw128);
              const uint64_t gd = make_smem_desc(gsb + qt * 8192, 512, 512, kLayoutSw64), qd = make_smem_desc(qsb + qt * 8192, 512, 512, kLayoutSw64);
              const uint64_t kd = make_smem_desc(ksb + kh * 8192, 512, 512, kLayoutSw64);
              for (int k = 0; k < 8; ++k) {                               // contraction over the 128 query rows of this tile
                mma_ss(tmem_base + kColDV + 32 * kh, pd + (uint64_t)k * (2048 >> 4), gd + (uint64_t)k * (1024 >> 4), id_mn, (qt | k) != 0);
                mma_ss(tmem_base + kColDK + 32 * kh, dd + (uint64_t)k * (2048 >> 4), qd + (uint64_t)k * (1024 >> 4), id_mn, (qt | k) != 0);
              }
              const uint64_t dk_ = make_smem_desc(dsb, 16, 1024, kLayoutSw128);         // dS as K-major A
              for (int k = 0; k < kx / 16; ++k) {                         // contraction over the keys of this half
                const uint64_t a = dk_ + (uint64_t)((k >> 2) * (16384 >> 4) + (k & 3) * 2);
                mma_ss(tmem_base + kColDQ, a, kd + (uint64_t)k * (1024 >> 4), id_k, (kh | k) != 0);
              }
              tc_commit(bar_o);
            }
            pending_o = true;
          }
        }
        if (pass == 0) {                                 // the two threads of a row meet
          Xd[half * 128 + row] = delta;
          __syncthreads();
          delta += Xd[(half ^ 1) * 128 + row];
          __syncthreads();                               // Xd free for the next query tile
        }
      }
      // dQ of this query tile (both key halves accumulated)
      mbar_wait(bar_o, ph_o); ph_o ^= 1; pending_o = false;
      tc_fence_after();
      {
        uint32_t q16[16];
        tmem_ld16(trow + kColDQ + half * 16, q16);
        tmem_wait_ld();
        if (qvalid) {
          __nv_bfloat16* pq = br.dq + (int64_t)mb * br.dq_bs + qtok * br.dq_ts + mhead * 32 + half * 16;
          uint32_t a[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] = pack_bf16x2(__uint_as_float(q16[2 * j]) * P.scale, __uint_as_float(q16[2 * j + 1]) * P.scale);
          *reinterpret_cast<uint4*>(pq) = make_uint4(a[0], a[1], a[2], a[3]); *reinterpret_cast<uint4*>(pq + 8) = make_uint4(a[4], a[5], a[6], a[7]);
        }
      }
      tc_fence_before();
      __syncthreads();                                   // dQ columns free for the next query tile
    }
    // dK, dV of both key halves (+ the LePE conv-transpose of G into dV)
#pragma unroll 1
    for (int kh = 0; kh < 2; ++kh) {
      const int n = kh * 128 + row;
      const bool valid = n < N;
      const int r_ = n / ws, c_ = n - r_ * ws;
      float lv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) lv[j] = 0.f;
      if (valid) {
        const float* wt = Wt + half * 16;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int rr = r_ - (t / 3 - 1), cc = c_ - (t % 3 - 1);       // the output position that read me through tap t
          if (rr >= 0 && rr < hs && cc >= 0 && cc < ws) {
            const int gr = rr * ws + cc;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
              const uint4 g4 = lds128(sw64(gsb, gr, half * 2 + ch));
              const float4 w0 = *reinterpret_cast<const float4*>(wt + t * 32 + ch * 8);
              const float4 w1 = *reinterpret_cast<const float4*>(wt + t * 32 + ch * 8 + 4);
              lv[ch * 8 + 0] = fmaf(w0.x, bf16_lo(g4.x), lv[ch * 8 + 0]); lv[ch * 8 + 1] = fmaf(w0.y, bf16_hi(g4.x), lv[ch * 8 + 1]);
              lv[ch * 8 + 2] = fmaf(w0.z, bf16_lo(g4.y), lv[ch * 8 + 2]); lv[ch * 8 + 3] = fmaf(w0.w, bf16_hi(g4.y), lv[ch * 8 + 3]);
              lv[ch * 8 + 4] = fmaf(w1.x, bf16_lo(g4.z), lv[ch * 8 + 4]); lv[ch * 8 + 5] = fmaf(w1.y, bf16_hi(g4.z), lv[ch * 8 + 5]);
              lv[ch * 8 + 6] = fmaf(w1.z, bf16_lo(g4.w), lv[ch * 8 + 6]); lv[ch * 8 + 7] = fmaf(w1.w, bf16_hi(g4.w), lv[ch * 8 + 7]);
            }
          }
        }
      }
      uint32_t k16[16], v16[16];
      tmem_ld16(trow + kColDK + 32 * kh + half * 16, k16);
      tmem_ld16(trow + kColDV + 32 * kh + half * 16, v16);
      tmem_wait_ld();
      if (valid) {
        const int64_t tok = (int64_t)(mih * hs + r_) * P.reso + (miw * ws + c_);
        const int64_t cho = mhead * 32 + half * 16;
        __nv_bfloat16* pk = br.dk + (int64_t)mb * br.dk_bs + tok * br.dk_ts + cho;
        __nv_bfloat16* pv = br.dv + (int64_t)mb * br.dv_bs + tok * br.dv_ts + cho;
        uint32_t b[8], c[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          b[j] = pack_bf16x2(__uint_as_float(k16[2 * j]) * P.scale, __uint_as_float(k16[2 * j + 1]) * P.scale);
          c[j] = pack_bf16x2(__uint_as_float(v16[2 * j]) + lv[2 * j], __uint_as_float(v16[2 * j + 1]) + lv[2 * j + 1]);
        }
        *reinterpret_cast<uint4*>(pk) = make_uint4(b[0], b[1], b[2], b[3]); *reinterpret_cast<uint4*>(pk + 8) = make_uint4(b[4], b[5], b[6], b[7]);
        *reinterpret_cast<uint4*>(pv) = make_uint4(c[0], c[1], c[2], c[3]); *reinterpret_cast<uint4*>(pv + 8) = make_uint4(c[4], c[5], c[6], c[7]);
      }
    }
    tc_fence_before();
    __syncthreads();                                     // closes the problem: shared-memory operands and TMEM are free again
  }
  if (warp == 0) tmem_dealloc(tmem_base, 512u);
}

// ------------------------------------------------------------------------------------------------------------------
// d get_v.weight[c, tap] = sum over (b, token) of G[b, tok, c] * V[b, tok + tap, c]  (neighbour inside the same window),
// d get_v.bias[c] = sum G[b, tok, c]   — autograd of the depthwise conv in LePEAttention.get_lepe (cswin_unet.py:67-80).
// ------------------------------------------------------------------------------------------------------------------
struct PgBranch {
  const __nv_bfloat16 *g, *v;
  float *dcw, *dcb;
  int64_t g_bs, g_ts, v_bs, v_ts;
  int hs, ws, ncg;                       // window, number of channel groups (of 4 << lq channels)
  uint32_t m_hs, m_ws;                   // ceil(2^32 / d) magic numbers (exact for n * d < 2^32)
};
struct PgParams {
  PgBranch br[2];
  int nb, reso, L, per, cluster, vec_atomics, lq;       // lq = log2(channel quads per token = lanes along the channels)
  uint32_t m_reso;
};
constexpr int kPgThr = 256;

__device__ __forceinline__ uint32_t fdiv(uint32_t n, uint32_t magic, int d) { return d == 1 ? n : __umulhi(n, magic); }

// grid = (token ranges of an image [cluster dimension], image, channel group over both branches).  A CTA is
// (quads = 8 / 16 / 32 lanes along the channels) x (256 / quads token lanes); a thread owns 4 channels.
__global__ void __launch_bounds__(kPgThr, 3) lepe_param_grad_kernel(const __grid_constant__ PgParams P) {
  __shared__ __align__(16) float4 slab[10 * kPgThr];     // [tap | bias][thread]: every thread's partial sums
  float* part = reinterpret_cast<float*>(slab);          // later: the CTA's sums in the layout of d w / d b
  const int tid = threadIdx.x;
  pdl_trigger();
  const int bi = (P.nb > 1 && (int)blockIdx.z >= P.br[0].ncg) ? 1 : 0;
  const PgBranch& br = P.br[bi];
  const int cg = blockIdx.z - (bi ? P.br[0].ncg : 0);
  const int b = blockIdx.y;
  const int pos0 = blockIdx.x * P.per, pos1 = min(P.L, pos0 + P.per);
  const int quads = 1 << P.lq, W = quads * 4, lanes = kPgThr >> P.lq;
  const int q4 = tid & (quads - 1), tl = tid >> P.lq;
  const int hs = br.hs, ws = br.ws, reso = P.reso;
  const __nv_bfloat16* gp = br.g + (int64_t)b * br.g_bs + cg * W + q4 * 4;
  const __nv_bfloat16* vp = br.v + (int64_t)b * br.v_bs + cg * W + q4 * 4;
  float2 acc[10][2];
#pragma unroll
  for (int t = 0; t < 10; ++t) acc[t][0] = acc[t][1] = make_float2(0.f, 0.f);
  pdl_wait();                                            // G is the previous kernel's output
  for (int pos = pos0 + tl; pos < pos1; pos += lanes) {
    const uint32_t y = fdiv((uint32_t)pos, P.m_reso, reso), x = (uint32_t)pos - y * reso;
    const int r = (int)(y - fdiv(y, br.m_hs, hs) * hs), c = (int)(x - fdiv(x, br.m_ws, ws) * ws);
    const __nv_bfloat16* vc = vp + (int64_t)pos * br.v_ts;
    // all ten loads are issued unconditionally (a neighbour outside the window reads the centre and is zeroed): one
    // memory round trip per token
    const uint2 gw = *reinterpret_cast<const uint2*>(gp + (int64_t)pos * br.g_ts);
    uint2 vw[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int dr = t / 3 - 1, dc = t % 3 - 1;
      const bool ok = r + dr >= 0 && r + dr < hs && c + dc >= 0 && c + dc < ws;
      vw[t] = *reinterpret_cast<const uint2*>(vc + (ok ? (int64_t)(dr * reso + dc) * br.v_ts : (int64_t)0));
      if (!ok) vw[t] = make_uint2(0u, 0u);
    }
    const float2 g0 = make_float2(bf16_lo(gw.x), bf16_hi(gw.x)), g1 = make_float2(bf16_lo(gw.y), bf16_hi(gw.y));
    acc[9][0] = fadd2(acc[9][0], g0); acc[9][1] = fadd2(acc[9][1], g1);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      acc[t][0] = ffma2(g0, make_float2(bf16_lo(vw[t].x), bf16_hi(vw[t].x)), acc[t][0]);
      acc[t][1] = ffma2(g1, make_float2(bf16_lo(vw[t].y), bf16_hi(vw[t].y)), acc[t][1]);
    }
  }
  // token lanes meet through shared memory (no atomics, no shuffles): every thread parks its 10 float4 sums, then one
  // thread per (tap, quad) adds up the token lanes
#pragma unroll
  for (int t = 0; t < 10; ++t) slab[t * kPgThr + tid] = make_float4(acc[t][0].x, acc[t][0].y, acc[t][1].x, acc[t][1].y);
  __syncthreads();
  float4 res[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int u = tid + k * kPgThr;                      // unit = (tap t, quad q); 10 * quads <= 320 units
    res[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (u < 10 * quads) {
      const int t = u >> P.lq, q = u & (quads - 1);
      for (int l = 0; l < lanes; ++l) {
        const float4 o = slab[t * kPgThr + (l << P.lq) + q];
        res[k].x += o.x; res[k].y += o.y; res[k].z += o.z; res[k].w += o.w;
      }
    }
  }
  __syncthreads();                                       // slab consumed; re-used as part[]
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int u = tid + k * kPgThr;
    if (u < 10 * quads) {
      const int t = u >> P.lq, q = u & (quads - 1);
      const float r4[4] = {res[k].x, res[k].y, res[k].z, res[k].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ch = q * 4 + e;
        part[t == 9 ? W * 9 + ch : ch * 9 + t] = r4[e];  // [W ch][9 taps] then [W] bias = the slices of d w / d b
      }
    }
  }
  // the CTAs of a cluster (token ranges of the same image and channel group) meet in rank 0 through distributed shared
  // memory: one set of global atomics (16 bytes each) per cluster
  const uint32_t rank = P.cluster > 1 ? cluster_ctarank() : 0u;
  if (P.cluster > 1) cluster_sync_all(); else __syncthreads();
  if (rank == 0) {
    for (int i = tid; i < 10 * quads; i += kPgThr) {
      float4 v = *reinterpret_cast<const float4*>(part + i * 4);
      for (uint32_t rk = 1; rk < (uint32_t)P.cluster; ++rk) {
        const float4 o = ld_dsmem_f4(dsmem_addr(smem_u32(part + i * 4), rk));
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      float* dst = i < 9 * quads ? br.dcw + (int64_t)cg * W * 9 + i * 4 : br.dcb + cg * W + (i - 9 * quads) * 4;
      if (P.vec_atomics) {
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      } else {
        atomicAdd(dst, v.x); atomicAdd(dst + 1, v.y); atomicAdd(dst + 2, v.z); atomicAdd(dst + 3, v.w);
      }
    }
  }
  if (P.cluster > 1) cluster_sync_all();                 // peers keep their shared memory alive until rank 0 has read it
}

bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// Eligibility of the streaming parameter-gradient kernel: head_dim-32 branches (C_b % 32 == 0), 8-byte aligned rows.
int lepe_param_grad_tc(const cswin_lepe_branch_grad_t* gs, int nb, int B, int reso, cudaStream_t stream, bool* handled) {
  *handled = false;
  for (int i = 0; i < nb; ++i) {
    const cswin_lepe_branch_grad_t& g = gs[i];
    const cswin_lepe_branch_t& s = g.fwd;
    if (!s.v || !g.dout || !g.dconv_w || !g.dconv_b) return CSWIN_OK;
    if (s.C_b <= 0 || s.C_b % 32 || s.H_sp <= 0 || s.W_sp <= 0 || reso % s.H_sp || reso % s.W_sp || reso > 1024) return CSWIN_OK;
    const int64_t st[] = {s.v_bs, s.v_ts, g.do_bs, g.do_ts};
    for (int64_t v : st) if (v <= 0 || (v * 2) % 8 != 0) return CSWIN_OK;
    if ((reinterpret_cast<uintptr_t>(s.v) & 7) || (reinterpret_cast<uintptr_t>(g.dout) & 7)) return CSWIN_OK;
  }
  if (B > 65535) return CSWIN_OK;
    PgParams G;
    G.nb = nb; G.reso = reso; G.L = reso * reso;
    G.m_reso = (uint32_t)((1ull << 32) / (uint64_t)reso + 1);
    // lanes along the channels: 32 quads (128 channels) when every branch allows it, else 16, else 8 (C_b = 32 heads)
    int lq = 5;
    for (int i = 0; i < nb; ++i) while (gs[i].fwd.C_b % (4 << lq)) --lq;
    const int W = 4 << lq, lanes = kPgThr >> lq;
    G.lq = lq;
    int ncg = 0, vec = 1;
    for (int i = 0; i < nb; ++i) {
      const cswin_lepe_branch_grad_t& g = gs[i];
      const cswin_lepe_branch_t& s = g.fwd;
      PgBranch& d = G.br[i];
      d.g = (const __nv_bfloat16*)g.dout; d.v = (const __nv_bfloat16*)s.v; d.dcw = g.dconv_w; d.dcb = g.dconv_b;
      d.g_bs = g.do_bs; d.g_ts = g.do_ts; d.v_bs = s.v_bs; d.v_ts = s.v_ts;
      d.hs = s.H_sp; d.ws = s.W_sp; d.ncg = s.C_b / W;
      d.m_hs = (uint32_t)((1ull << 32) / (uint64_t)s.H_sp + 1); d.m_ws = (uint32_t)((1ull << 32) / (uint64_t)s.W_sp + 1);
      ncg += d.ncg;
      if (!al16(g.dconv_w) || !al16(g.dconv_b)) vec = 0;
    }
    if (nb == 1) G.br[1] = G.br[0];
    G.vec_atomics = vec;
    // 2 tokens per thread (4, 8, ... when that would need more than ~12 CTAs per SM); the token ranges of one (image,
    // channel group) form a cluster of up to 8 CTAs
    int tpt = 2, ranges = 1, cluster = 1;
    for (;; tpt *= 2) {
      ranges = std::max(1, (G.L + lanes * tpt - 1) / (lanes * tpt));
      if ((int64_t)ranges * B * ncg <= 12ll * sm_count() || ranges == 1) break;
    }
    if (ranges >= 8) { ranges = (ranges + 7) / 8 * 8; cluster = 8; }
    else { while (cluster < ranges) cluster <<= 1; ranges = cluster; }
    G.per = (G.L + ranges - 1) / ranges; G.cluster = cluster;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)ranges, (unsigned)B, (unsigned)ncg);
    lc.blockDim = dim3(kPgThr);
    lc.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (cluster > 1) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = (unsigned)cluster; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
      ++na;
    }
    if (pdl_enabled()) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    lc.attrs = attr; lc.numAttrs = na;
    CSWIN_CUDA_OK(cudaLaunchKernelEx(&lc, lepe_param_grad_kernel, G));
    CSWIN_LAUNCH_CHECK();
  *handled = true;                                       // (a streaming kernel: not counted as a tcgen05 launch)
  return CSWIN_OK;
}

int lepe_attention_bwd_tc(const cswin_lepe_branch_grad_t* gs, int nb, int B, int reso, float scale, cudaStream_t stream,
                          bool* handled) {
  *handled = false;
  int n_wide = 0;                                        // branches with 128 < N <= 256: wide kernel (all or none)
  for (int i = 0; i < nb; ++i) {
    const cswin_lepe_branch_grad_t& g = gs[i];
    const cswin_lepe_branch_t& s = g.fwd;
    if (!s.q || !s.k || !s.v || !s.conv_w || !s.lse || !g.dout || !g.dq || !g.dk || !g.dv || ((g.dconv_w != nullptr) != (g.dconv_b != nullptr))) return CSWIN_OK;
    if (s.heads <= 0 || s.C_b != s.heads * 32) return CSWIN_OK;
    if (s.H_sp <= 0 || s.W_sp <= 0 || reso % s.H_sp || reso % s.W_sp || s.H_sp * s.W_sp > 256 || s.H_sp > 256 || s.W_sp > 256) return CSWIN_OK;
    if (s.H_sp * s.W_sp > 128) ++n_wide;
    const int64_t st[] = {s.q_bs, s.q_ts, s.k_bs, s.k_ts, s.v_bs, s.v_ts, g.do_bs, g.do_ts, g.dq_bs, g.dq_ts, g.dk_bs, g.dk_ts, g.dv_bs, g.dv_ts};
    for (int64_t v : st) if (v <= 0 || (v * 2) % 16 != 0) return CSWIN_OK;
    if (!al16(s.q) || !al16(s.k) || !al16(s.v) || !al16(g.dout) || !al16(g.dq) || !al16(g.dk) || !al16(g.dv) || !al16(s.conv_w)) return CSWIN_OK;
  }
  if (tc::encode_tiled_fn() == nullptr) return CSWIN_OK;
  if (n_wide != 0 && n_wide != nb) return CSWIN_OK;
  const bool wide = n_wide != 0;
  if (gs[0].dconv_w != nullptr) {                        // parameter gradients first: if their kernel declines, so does this path
    bool pg = false;
    const int rc = lepe_param_grad_tc(gs, nb, B, reso, stream, &pg);
    if (rc != CSWIN_OK) return rc;
    if (!pg) return CSWIN_OK;
  }
  BwdParams P;
  P.nb = nb; P.reso = reso; P.scale = scale; P.scale_log2e = scale * 1.4426950408889634f;
  P.trace = g_trace.load(std::memory_order_relaxed);
  int tiles = 0;
  for (int i = 0; i < nb; ++i) {
    const cswin_lepe_branch_grad_t& g = gs[i];
    const cswin_lepe_branch_t& s = g.fwd;
    BwdBranch& d = P.br[i];
    d.dq = (__nv_bfloat16*)g.dq; d.dk = (__nv_bfloat16*)g.dk; d.dv = (__nv_bfloat16*)g.dv;
    d.cw = (const __nv_bfloat16*)s.conv_w; d.lse = s.lse;
    d.dq_bs = g.dq_bs; d.dq_ts = g.dq_ts; d.dk_bs = g.dk_bs; d.dk_ts = g.dk_ts; d.dv_bs = g.dv_bs; d.dv_ts = g.dv_ts;
    d.heads = s.heads; d.hs = s.H_sp; d.ws = s.W_sp; d.nww = reso / s.W_sp;
    d.nwin = (reso / s.H_sp) * (reso / s.W_sp); d.N = s.H_sp * s.W_sp; d.nprob = B * d.nwin * d.heads;
    d.tile_begin = tiles;
    const int slots = d.N <= 64 ? 2 : 1;
    tiles += wide ? d.nprob : (d.nprob + slots - 1) / slots;
    const void* ptr[4] = {s.q, s.k, s.v, g.dout};
    const int64_t bs[4] = {s.q_bs, s.k_bs, s.v_bs, g.do_bs}, ts[4] = {s.q_ts, s.k_ts, s.v_ts, g.do_ts};
    for (int j = 0; j < 4; ++j) {
      const uint64_t dims[4] = {(uint64_t)s.C_b, (uint64_t)reso, (uint64_t)reso, (uint64_t)B};
      const uint64_t str[3] = {(uint64_t)ts[j] * 2, (uint64_t)ts[j] * 2 * reso, (uint64_t)bs[j] * 2};
      const uint32_t box[4] = {32, (uint32_t)s.W_sp, (uint32_t)s.H_sp, 1};
      if (!tc::make_tensor_map_bf16(&P.map[i][j], ptr[j], 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B)) return CSWIN_ERR_CUDA;
    }
  }
  if (nb == 1) P.br[1] = P.br[0];
  P.total_tiles = tiles;
  // persistent grid: 2 CTAs per SM; rounded down to a multiple of the tile period of the heads, so that the tiles a CTA
  // walks (stride = grid) keep their heads and the LePE weights are staged once per CTA
  int grid = tiles;
  const int cap = 2 * sm_count();
  if (tiles > cap) {
    int period = 1;                                       // heads are powers of two in every CSWin configuration
    for (int i = 0; i < nb; ++i) {
      const int h = P.br[i].heads, sl = P.br[i].N <= 64 ? 2 : 1;
      const int need = (sl == 2 && h % 2 == 0) ? h / 2 : h;
      if ((need & (need - 1)) == 0) period = std::max(period, need);
    }
    grid = (period <= cap) ? cap / period * period : cap;
  }
  static std::atomic<bool> configured{false};
  if (!configured.exchange(true)) {
    CSWIN_CUDA_OK(cudaFuncSetAttribute(lepe_attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    CSWIN_CUDA_OK(cudaFuncSetAttribute(lepe_attn_bwd_wide_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemWideBwd));
  }
  if (wide) CSWIN_CUDA_OK(launch_pdl(lepe_attn_bwd_wide_tc_kernel, dim3(std::min(tiles, sm_count())), dim3(kThr), (size_t)kSmemWideBwd, stream, P));
  else CSWIN_CUDA_OK(launch_pdl(lepe_attn_bwd_tc_kernel, dim3(grid), dim3(kThr), (size_t)kSmem, stream, P));
  CSWIN_LAUNCH_CHECK();
  g_tc_launches.fetch_add(1, std::memory_order_relaxed);
  *handled = true;
  return CSWIN_OK;
}

}  // namespace cswin
