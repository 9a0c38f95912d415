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
#include "common.cuh"
#include "tc_common.cuh"

namespace cswin {
namespace {

using namespace tc;

struct BwdBranch {
  __nv_bfloat16 *dq, *dk, *dv;
  const __nv_bfloat16* cw;
  const float* lse;
  float *dcw, *dcb;
  int64_t dq_bs, dq_ts, dk_bs, dk_ts, dv_bs, dv_ts;
  int heads, hs, ws, nww, nwin, N, tile_begin, nprob;
};
struct alignas(64) BwdParams {
  CUtensorMap map[2][4];      // [branch][q, k, v, g]
  BwdBranch br[2];
  int nb, reso;
  float scale, scale_log2e;
  unsigned long long* trace;
};

constexpr int kRows = 128, kThr = 256, kRowB = 64;
constexpr int kOpB = kRows * kRowB;                 // 8 KB per [128][32] bf16 operand
constexpr int kPB = 2 * kRows * 128;                // 32 KB per [128][128] bf16 matrix (two 64-column blocks)
constexpr uint32_t kTmem = 256;
constexpr int kSmem = 4 * kOpB + 2 * kPB + 2 * 9 * 32 * 4 + 2 * 10 * 32 * 4 + 2 * 128 * 4 + 64 + 1024;

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
__device__ __forceinline__ float lds_bf16(uint32_t addr) {
  unsigned short h;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(addr));
  return __uint_as_float((uint32_t)h << 16);
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
  float* Part = Wt + 2 * 9 * 32;                        // [2][10][32]   d w (9 taps) and d b of the tile's heads
  float* Xd = Part + 2 * 10 * 32;                       // [2][128]      delta exchange
  uint64_t* bars = reinterpret_cast<uint64_t*>(Xd + 2 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int tid = threadIdx.x, warp = tid >> 5;
  pdl_trigger();
  if (tid == 0) trace_stamp(P.trace, 0);
  const int row = tid & 127, half = tid >> 7;
  const int bi = (P.nb > 1 && (int)blockIdx.x >= P.br[1].tile_begin) ? 1 : 0;
  const BwdBranch& br = P.br[bi];
  const int tile = blockIdx.x - br.tile_begin;
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

  const uint32_t bar_tma = smem_u32(&bars[0]), bar_s = smem_u32(&bars[1]), bar_o = smem_u32(&bars[2]);
  if (warp == 0) { tmem_alloc(smem_u32(tmem_slot), kTmem); tmem_relinquish(); }
  if (tid == 32) { mbar_init(bar_tma, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1); fence_barrier_init(); }
  // zero the rows of Q, K, V, G that TMA does not write (they are contraction rows of dQ / dK / dV)
  for (int i = tid; i < 4 * kRows * 4; i += kThr) {
    const int r = (i >> 2) & 127;
    const int s = r / slot_rows, rn = r - s * slot_rows;
    if (s >= np || rn >= N) *reinterpret_cast<uint4*>(smem + i * 16) = make_uint4(0, 0, 0, 0);
  }
  for (int i = tid; i < 2 * 10 * 32; i += kThr) Part[i] = 0.f;
  if (tid < np * 36) {                                   // LePE weights of the tile's head(s): Wt[slot][tap][ch]
    const int s = tid / 36, i = tid - s * 36;
    const int hd = (p0 + s) % br.heads;
    const uint4 raw = *reinterpret_cast<const uint4*>(br.cw + (size_t)hd * 288 + i * 8);
    const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int idx = i * 8 + e, ch = idx / 9, t = idx - ch * 9;
      Wt[(s * 9 + t) * 32 + ch] = (e & 1) ? bf16_hi(w4[e >> 1]) : bf16_lo(w4[e >> 1]);
    }
  }
  const float lse = valid ? br.lse[((int64_t)mb * P.reso * P.reso + tok) * br.heads + mhead] : 0.f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) trace_stamp(P.trace, 1);
  pdl_wait();                                            // d out (previous kernel's output) and the gradient buffers are safe from here

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
    mbar_wait(bar_tma, 0);
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
  mbar_wait(bar_s, 0);
  if (tid == 0) trace_stamp(P.trace, 2);
  tc_fence_after();

  // ---- P, delta, dS for row `row`, key columns [kbeg, kbeg + hcols) of my slot ----
  const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
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
  const uint32_t ps = smem_u32(Ps), dsb = smem_u32(Ds);
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
  mbar_wait(bar_tma, 0);
  float lv[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) lv[j] = 0.f;
  const uint32_t gsb = smem_u32(Gs), vsb = smem_u32(Vs);
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
  // ---- d w[ch][tap] / d b[ch]: thread = (channel tid & 31, token group tid >> 5); every thread walks its share of the
  //      window's tokens keeping the 9 tap sums + the bias sum in registers, groups are merged with shared-memory atomics
  for (int s = 0; s < np; ++s) {
    const int ch = tid & 31;
    const int cchunk = ch >> 3, coff = (ch & 7) * 2;
    float acc[10];
#pragma unroll
    for (int t = 0; t < 10; ++t) acc[t] = 0.f;
    for (int nn = tid >> 5; nn < N; nn += kThr / 32) {
      const int r = nn / ws, c = nn - r * ws;
      const float g = lds_bf16(sw64(gsb, s * slot_rows + nn, cchunk) + coff);
      acc[9] += g;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int rr = r + t / 3 - 1, cc = c + t % 3 - 1;
        if (rr >= 0 && rr < hs && cc >= 0 && cc < ws)
          acc[t] = fmaf(g, lds_bf16(sw64(vsb, s * slot_rows + rr * ws + cc, cchunk) + coff), acc[t]);
      }
    }
#pragma unroll
    for (int t = 0; t < 10; ++t) atomicAdd(&Part[(s * 10 + t) * 32 + ch], acc[t]);
  }
  __syncthreads();
  for (int o = tid; o < np * 320; o += kThr) {
    const int s = o / 320, rem = o - s * 320;
    const int t = rem >> 5, ch = rem & 31;                  // t = 9 -> bias
    const int hd = (p0 + s) % br.heads;
    const float v = Part[(s * 10 + t) * 32 + ch];
    if (t == 9) atomicAdd(br.dcb + hd * 32 + ch, v);
    else atomicAdd(br.dcw + (int64_t)(hd * 32 + ch) * 9 + t, v);
  }

  if (tid == 0) trace_stamp(P.trace, 5);
  mbar_wait(bar_o, 0);
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
  __syncthreads();
  if (tid == 0) trace_stamp(P.trace, 7);
  if (warp == 0) tmem_dealloc(tmem_base, kTmem);
}

bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int lepe_attention_bwd_tc(const cswin_lepe_branch_grad_t* gs, int nb, int B, int reso, float scale, cudaStream_t stream,
                          bool* handled) {
  *handled = false;
  for (int i = 0; i < nb; ++i) {
    const cswin_lepe_branch_grad_t& g = gs[i];
    const cswin_lepe_branch_t& s = g.fwd;
    if (!s.q || !s.k || !s.v || !s.conv_w || !s.lse || !g.dout || !g.dq || !g.dk || !g.dv || !g.dconv_w || !g.dconv_b) return CSWIN_OK;
    if (s.heads <= 0 || s.C_b != s.heads * 32) return CSWIN_OK;
    if (s.H_sp <= 0 || s.W_sp <= 0 || reso % s.H_sp || reso % s.W_sp || s.H_sp * s.W_sp > 128 || s.H_sp > 256 || s.W_sp > 256) return CSWIN_OK;
    const int64_t st[] = {s.q_bs, s.q_ts, s.k_bs, s.k_ts, s.v_bs, s.v_ts, g.do_bs, g.do_ts, g.dq_bs, g.dq_ts, g.dk_bs, g.dk_ts, g.dv_bs, g.dv_ts};
    for (int64_t v : st) if (v <= 0 || (v * 2) % 16 != 0) return CSWIN_OK;
    if (!al16(s.q) || !al16(s.k) || !al16(s.v) || !al16(g.dout) || !al16(g.dq) || !al16(g.dk) || !al16(g.dv) || !al16(s.conv_w)) return CSWIN_OK;
  }
  if (tc::encode_tiled_fn() == nullptr) return CSWIN_OK;
  BwdParams P;
  P.nb = nb; P.reso = reso; P.scale = scale; P.scale_log2e = scale * 1.4426950408889634f;
  P.trace = g_trace.load(std::memory_order_relaxed);
  int tiles = 0;
  for (int i = 0; i < nb; ++i) {
    const cswin_lepe_branch_grad_t& g = gs[i];
    const cswin_lepe_branch_t& s = g.fwd;
    BwdBranch& d = P.br[i];
    d.dq = (__nv_bfloat16*)g.dq; d.dk = (__nv_bfloat16*)g.dk; d.dv = (__nv_bfloat16*)g.dv;
    d.cw = (const __nv_bfloat16*)s.conv_w; d.lse = s.lse; d.dcw = g.dconv_w; d.dcb = g.dconv_b;
    d.dq_bs = g.dq_bs; d.dq_ts = g.dq_ts; d.dk_bs = g.dk_bs; d.dk_ts = g.dk_ts; d.dv_bs = g.dv_bs; d.dv_ts = g.dv_ts;
    d.heads = s.heads; d.hs = s.H_sp; d.ws = s.W_sp; d.nww = reso / s.W_sp;
    d.nwin = (reso / s.H_sp) * (reso / s.W_sp); d.N = s.H_sp * s.W_sp; d.nprob = B * d.nwin * d.heads;
    d.tile_begin = tiles;
    const int slots = d.N <= 64 ? 2 : 1;
    tiles += (d.nprob + slots - 1) / slots;
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
  static std::atomic<bool> configured{false};
  if (!configured.exchange(true))
    CSWIN_CUDA_OK(cudaFuncSetAttribute(lepe_attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  CSWIN_CUDA_OK(launch_pdl(lepe_attn_bwd_tc_kernel, dim3(tiles), dim3(kThr), (size_t)kSmem, stream, P));
  CSWIN_LAUNCH_CHECK();
  g_tc_launches.fetch_add(1, std::memory_order_relaxed);
  *handled = true;
  return CSWIN_OK;
}

}  // namespace cswin
