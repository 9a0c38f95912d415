// attention_tc.cu — fused LePE cross-shaped-window attention, bf16, tcgen05 / TMEM / TMA (sm_100a).
//
// Replaces LePEAttention.forward and everything it calls (networks/cswin_unet.py:59-109, 184-202) plus the branch
// concat of CSWinBlock.forward (:172-176).  Both stripe branches of a block run in ONE launch.
//
// Work unit ("tile") = 128 query rows = 128 TMEM lanes, served by 256 threads (two threads per row, each owning
// half of the row's key columns / output channels):
//     windows of N <= 64 tokens   : two (window, head) problems per tile (rows 0..63 and 64..127),
//     windows of 64 < N <= 128    : one problem per tile.
// Per tile
//   1. TMA gathers the stripe window straight out of the (B, H, W, 3C) qkv tensor: one 4-D box
//      (32 channels, W_sp, H_sp, 1) per operand lands as [token][32 ch] rows of 64 B in the 64-byte-swizzled
//      UMMA layout — the img2windows / head-split copies of the reference become address arithmetic in the
//      tensor map;
//   2. S = Q K^T: two tcgen05.mma (M128 x N<=128 x K16, both operands K-major from smem), fp32 in TMEM columns
//      [0,128).  With two problems per tile the off-diagonal 64x64 blocks are computed and ignored;
//   3. softmax: the two threads of row i read their halves of S row i from TMEM (tcgen05.ld 32x32b), exchange the
//      row max through smem, exponentiate in fp32 (ex2.approx with the qk scale folded in) and write bf16 P back
//      into TMEM columns [0,64) (zeros for padded keys and for the other problem's keys);
//   4. O = P V: tcgen05.mma with A = P from TMEM and B = V from smem (MN-major, the layout TMA delivered),
//      N = 32, accumulating in TMEM columns [64,96);
//   5. while that runs, each thread computes the LePE depthwise 3x3 conv of its token for its 16 channels from the
//      same V tile in smem (zero padding at the WINDOW border), then adds it to O / rowsum and stores 32 B
//      straight into the (B, L, C) concat layout (windows2img + cat are address arithmetic).
// HBM traffic = q, k, v read once + out written once; the N x N scores never leave TMEM.
//
// Wide variant (kWide, windows of 128 < N <= 256 tokens — the 512^2 configuration's 32x8 / 8x32 / 16x16 stripes): a
// (window, head) problem is two tiles of 128 query rows, each against all N keys: K / V tiles of 256 rows, S = 128 x 256
// fp32 in TMEM columns [0,256), P (bf16) back into columns [0,128), O into columns [128,160) (S is dead by then); each of
// the two threads of a row owns 128 key columns.  Needs 128 % W_sp == 0 so that a query tile is a rectangular TMA box.
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace cswin {
namespace {

using namespace tc;

struct TcBranch {
  __nv_bfloat16* out; const __nv_bfloat16* cw; const __nv_bfloat16* cb; float* lse;
  int64_t o_bs, o_ts;
  int heads, hs, ws, nww, nwin, N, tile_begin, nprob;
  uint32_t m_heads, m_nwin, m_nww, m_ws;      // floor(2^32 / d) + 1: x / d == __umulhi(x, m) for the index ranges of this kernel
};
struct alignas(64) TcParams {
  CUtensorMap map[2][3];     // [branch][q,k,v]
  TcBranch br[2];
  int nb, reso;
  float scale, scale_log2e;
  unsigned long long* trace;
};

constexpr int kTileRows = 128;
constexpr int kThreads = 256;
constexpr int kRowBytes = 64;                         // 32 bf16 channels of one head
constexpr int kOperandBytes = kTileRows * kRowBytes;  // 8 KB
constexpr uint32_t kTmemCols = 128;
//                         Q,K,V               Wt[2][9][32]     Bc[2][32]    xchg[2][128]   barriers   align slack
constexpr int kSmemBytes = 3 * kOperandBytes + 2 * 9 * 32 * 4 + 2 * 32 * 4 + 2 * 128 * 4 + 64 + 1024;
constexpr int kSmemBytesWide = 5 * kOperandBytes + 2 * 9 * 32 * 4 + 2 * 32 * 4 + 2 * 128 * 4 + 64 + 1024;   // K, V: 256 rows

__device__ __forceinline__ uint32_t v_chunk_addr(uint32_t vbase, int row, int chunk) {
  return vbase + row * kRowBytes + (((chunk ^ (row >> 1)) & 3) << 4);          // Swizzle<2,4,3> (64-byte swizzle)
}
// x / d for d > 1 through the precomputed multiplier (exact while x * d < 2^32); d == 1 passes x through
__device__ __forceinline__ int fast_div(int x, int d, uint32_t m) { return d == 1 ? x : (int)__umulhi((uint32_t)x, m); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// kMinB = CTAs per SM the register budget is sized for.  3 (80 registers) for launches of at most one wave: the tile's own latency is
// what counts (stage 3 at batch 24: 7.4 us vs 9.2 us with the spills of the 64-register build); 4 (64 registers, 252 bytes of spill
// stores) for multi-wave launches, where a fourth resident tile hides more latency than the spills cost (batch 96: 757 -> 713 us for the
// 26 launches of a forward).  The launcher picks by tile count.
template <bool kWide, int kMinB>
__global__ void __launch_bounds__(kThreads, kMinB) lepe_attn_fwd_tc_kernel(const __grid_constant__ TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the shared array: keeps the shared address space (LDS / STS, not generic LD / ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int kKvBytes = kWide ? 2 * kOperandBytes : kOperandBytes;   // K / V tile: 256 or 128 rows
  constexpr int kKvRows = kWide ? 256 : 128;
  constexpr int kCh = kWide ? 4 : 2;                                    // 32-column chunks of S per thread
  constexpr uint32_t kOCol = kWide ? 128 : 64;                          // TMEM column of the O accumulator
  uint8_t* Qs = smem;
  uint8_t* Ks = smem + kOperandBytes;
  uint8_t* Vs = Ks + kKvBytes;
  __nv_bfloat16* Wt = reinterpret_cast<__nv_bfloat16*>(Vs + kKvBytes);  // [2][9][32] bf16 (tap-major copy of the (32, 3, 3) conv weights)
  float* Bc = reinterpret_cast<float*>(Wt + 2 * 9 * 32 * 2);           // [2][32]   (the region is sized for the former fp32 copy)
  float* Xmax = Bc + 2 * 32;                                          // [2 halves][128 rows]
  float* Xsum = Xmax;                                                 // reused after the max exchange
  uint64_t* bars = reinterpret_cast<uint64_t*>(Xmax + 2 * 128);       // tma, s, o
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int tid = threadIdx.x, warp = tid >> 5;
  pdl_trigger();
  if (tid == 0) trace_stamp(P.trace, 0);
  const int row = tid & 127;                   // tile row == TMEM lane
  const int half = tid >> 7;                   // which half of the row's columns / channels this thread owns
  const int bi = (P.nb > 1 && (int)blockIdx.x >= P.br[1].tile_begin) ? 1 : 0;
  const TcBranch& br = P.br[bi];
  const int tile = blockIdx.x - br.tile_begin;
  const int N = br.N, hs = br.hs, ws = br.ws;
  const int slots = (!kWide && N <= 64) ? 2 : 1;
  const int slot_rows = slots == 2 ? kTileRows / 2 : kTileRows;
  const int qt = kWide ? (tile & 1) : 0;                             // wide: which 128-row query tile of the problem
  const int p0 = kWide ? (tile >> 1) : tile * slots;
  const int np = min(slots, br.nprob - p0);
  const int kext = (slots == 2) ? 128 : ((N + 15) & ~15);            // kv extent fed to the P.V MMA

  const int slot = slots == 2 ? row >> 6 : 0;  // warp-uniform
  const int n = row - slot * slot_rows + qt * kTileRows;             // token index inside the window

  // (batch, window, head) of my slot, and of both slots for the TMA-issuing thread
  int mb, mih, miw, mhead;
  {
    const int local = p0 + min(slot, np - 1);
    const int lw = fast_div(local, br.heads, br.m_heads);
    mhead = local - lw * br.heads;
    mb = fast_div(lw, br.nwin, br.m_nwin);
    const int win = lw - mb * br.nwin;
    mih = fast_div(win, br.nww, br.m_nww); miw = win - mih * br.nww;
  }

  const uint32_t bar_tma = smem_u32(&bars[0]), bar_s = smem_u32(&bars[1]), bar_o = smem_u32(&bars[2]);
  if (warp == 0) { tmem_alloc(smem_u32(tmem_slot), kWide ? 256u : kTmemCols); tmem_relinquish(); }
  if (tid == 32) { mbar_init(bar_tma, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1); fence_barrier_init(); }

  // zero the V rows the P.V MMA reads but TMA does not write (0 * stale-NaN would poison O)
  for (int i = tid; i < kKvRows * 4; i += kThreads) {
    const int r = i >> 2;
    const int s = (kWide || slots == 1) ? 0 : r >> 6, rn = r - s * slot_rows;
    if (r < kext && (s >= np || rn >= N)) *reinterpret_cast<uint4*>(Vs + i * 16) = make_uint4(0, 0, 0, 0);
  }
  // stage the LePE weights of the head(s) of this tile: 288 contiguous bf16 per head -> Wt[slot][tap][ch] (still bf16: FHFMA.BF16)
  if (tid < np * 36) {
    const int s = tid >= 36 ? 1 : 0, i = tid - s * 36;                // 36 x 16-byte chunks per head
    const int local = p0 + s;
    const int hd = local - fast_div(local, br.heads, br.m_heads) * br.heads;
    const uint4 raw = *reinterpret_cast<const uint4*>(br.cw + (size_t)hd * 288 + i * 8);
    const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
    unsigned short* wt16 = reinterpret_cast<unsigned short*>(Wt);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int idx = i * 8 + e;                                       // = ch * 9 + tap
      const int ch = (idx * 7282) >> 16, t = idx - ch * 9;             // idx / 9 for idx < 288
      wt16[(s * 9 + t) * 32 + ch] = (unsigned short)((e & 1) ? (w4[e >> 1] >> 16) : (w4[e >> 1] & 0xffffu));
    }
  } else if (tid >= 128 && tid < 128 + np * 32) {
    const int s = (tid - 128) >> 5, ch = tid & 31;
    const int hd = (p0 + s) - fast_div(p0 + s, br.heads, br.m_heads) * br.heads;
    Bc[s * 32 + ch] = __bfloat162float(br.cb[hd * 32 + ch]);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) trace_stamp(P.trace, 1);                     // prologue done
  pdl_wait();                                                // q, k, v (previous kernel's output) and `out` are safe from here

  if (warp == 0 && elect_one()) {     // one elected lane, warp-uniform datapath for the TMA / tcgen05 issue
    mbar_expect_tx(bar_tma, kWide ? (uint32_t)((2 * N + kTileRows) * kRowBytes) : (uint32_t)(np * 3 * N * kRowBytes));
    if (kWide) {                      // q: rows [128 qt, +128) of the window (a box of 128 / W_sp window rows); k, v: the window
      const int lw = fast_div(p0, br.heads, br.m_heads);
      const int hd = p0 - lw * br.heads;
      const int b = fast_div(lw, br.nwin, br.m_nwin);
      const int win = lw - b * br.nwin;
      const int ih = fast_div(win, br.nww, br.m_nww), iw = win - ih * br.nww;
      tma_load_4d(smem_u32(Qs), &P.map[bi][0], bar_tma, hd * 32, iw * ws, ih * hs + qt * fast_div(kTileRows, ws, br.m_ws), b);
      tma_load_4d(smem_u32(Ks), &P.map[bi][1], bar_tma, hd * 32, iw * ws, ih * hs, b);
      tma_load_4d(smem_u32(Vs), &P.map[bi][2], bar_tma, hd * 32, iw * ws, ih * hs, b);
    }
    for (int s = 0; s < (kWide ? 0 : np); ++s) {
      const int local = p0 + s;
      const int lw = fast_div(local, br.heads, br.m_heads);
      const int hd = local - lw * br.heads;
      const int b = fast_div(lw, br.nwin, br.m_nwin);
      const int win = lw - b * br.nwin;
      const int ih = fast_div(win, br.nww, br.m_nww), iw = win - ih * br.nww;
      const int c0 = hd * 32, c1 = iw * ws, c2 = ih * hs, c3 = b;
      const uint32_t off = s * slot_rows * kRowBytes;
      tma_load_4d(smem_u32(Qs) + off, &P.map[bi][0], bar_tma, c0, c1, c2, c3);
      tma_load_4d(smem_u32(Ks) + off, &P.map[bi][1], bar_tma, c0, c1, c2, c3);
      tma_load_4d(smem_u32(Vs) + off, &P.map[bi][2], bar_tma, c0, c1, c2, c3);
    }
    mbar_wait(bar_tma, 0);
    trace_stamp(P.trace, 2);                                 // q,k,v landed
    tc_fence_after();
    const uint64_t qd = make_smem_desc(smem_u32(Qs), 16, 8 * kRowBytes, kLayoutSw64);
    const uint64_t kd = make_smem_desc(smem_u32(Ks), 16, 8 * kRowBytes, kLayoutSw64);
    const uint32_t idesc = make_idesc_bf16(128, kext, 0, 0);
    mma_ss(tmem_base, qd, kd, idesc, false);
    mma_ss(tmem_base, qd + 2, kd + 2, idesc, true);          // +32 B along K inside the swizzled row
    tc_commit(bar_s);
  }
  mbar_wait(bar_s, 0);
  if (tid == 0) trace_stamp(P.trace, 3);                     // S ready
  tc_fence_after();

  // ---- softmax: this thread owns columns [cbeg, cbeg + hcols) of S row `row` (slot-local key index kbeg..) ----
  const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const int hcols = kWide ? 128 : (slot_rows >> 1);          // 32 (two problems), 64 (one problem) or 128 (wide)
  const int kbeg = half * hcols;                             // first slot-local key of my half
  const int cbeg = slot * slot_rows + kbeg;                  // first S column of my half
  const int nch = hcols >> 5;                                // chunks of 32 columns: 1, 2 or 4
  float mx = -INFINITY;
  for (int c = 0; c < nch; ++c) {
    if (kbeg + 32 * c >= kext) break;                        // (one-problem mode) chunk entirely beyond the keys
    uint32_t v[32];
    tmem_ld32(trow + cbeg + 32 * c, v);
    tmem_wait_ld();
    const int lim = N - (kbeg + 32 * c);                     // valid columns in this chunk (warp-uniform)
    if (lim >= 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
    } else {
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {                       // groups of 8 columns: whole, ragged or beyond the keys
        if (8 * g8 + 8 <= lim) {
#pragma unroll
          for (int j = 8 * g8; j < 8 * g8 + 8; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        } else if (8 * g8 < lim) {
#pragma unroll
          for (int j = 8 * g8; j < 8 * g8 + 8; ++j) if (j < lim) mx = fmaxf(mx, __uint_as_float(v[j]));
        }
      }
    }
  }
  Xmax[half * 128 + row] = mx;
  __syncthreads();
  mx = fmaxf(mx, Xmax[(half ^ 1) * 128 + row]);
  const float mxs = mx * P.scale_log2e;
  float sum = 0.f;
  float2 sum2 = make_float2(0.f, 0.f);
  const float2 sl2 = make_float2(P.scale_log2e, P.scale_log2e), nmxs2 = make_float2(-mxs, -mxs);
  uint32_t pk[kCh][16];                                      // bf16 P of my columns, held until every S read is done
#pragma unroll
  for (int c = 0; c < kCh; ++c) {
    if (c < nch && kbeg + 32 * c < kext) {
      uint32_t v[32];
      tmem_ld32(trow + cbeg + 32 * c, v);
      tmem_wait_ld();
      const int lim = N - (kbeg + 32 * c);
      if (lim >= 32) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {                    // packed pairs: one FFMA2 + one FADD2 per two scores
          const float2 t = ffma2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sl2, nmxs2);
          const float2 e = make_float2(ex2_approx(t.x), ex2_approx(t.y));
          sum2 = fadd2(sum2, e);                             // fp32 denominator (the reference normalises before rounding P)
          pk[c][j >> 1] = pack_bf16x2(e.x, e.y);
        }
      } else {
        // ragged chunk: whole groups of 8 columns take the packed path, the one ragged group is predicated, groups beyond the
        // keys only write zeros — `lim` is warp-uniform, so these are real (non-divergent) branches, not predicated-off work
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          if (8 * g8 + 8 <= lim) {
#pragma unroll
            for (int j = 8 * g8; j < 8 * g8 + 8; j += 2) {
              const float2 t = ffma2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sl2, nmxs2);
              const float2 e = make_float2(ex2_approx(t.x), ex2_approx(t.y));
              sum2 = fadd2(sum2, e);
              pk[c][j >> 1] = pack_bf16x2(e.x, e.y);
            }
          } else if (8 * g8 < lim) {
#pragma unroll
            for (int j = 8 * g8; j < 8 * g8 + 8; j += 2) {
              const float e0 = (j < lim) ? ex2_approx(fmaf(__uint_as_float(v[j]), P.scale_log2e, -mxs)) : 0.f;
              const float e1 = (j + 1 < lim) ? ex2_approx(fmaf(__uint_as_float(v[j + 1]), P.scale_log2e, -mxs)) : 0.f;
              sum += e0 + e1;
              pk[c][j >> 1] = pack_bf16x2(e0, e1);
            }
          } else {
#pragma unroll
            for (int j = 8 * g8; j < 8 * g8 + 8; j += 2) pk[c][j >> 1] = 0u;
          }
        }
      }
    }
  }
  // P (bf16, TMEM columns [0,64)) aliases S columns that the partner thread of this row may still be reading:
  // publish only after every thread has its S values in registers.
  __syncthreads();
#pragma unroll
  for (int c = 0; c < kCh; ++c)
    if (c < nch && kbeg + 32 * c < kext) tmem_st16(trow + ((cbeg + 32 * c) >> 1), pk[c]);
  if (slots == 2) {                                          // keys of the other problem: P = 0 (16 of its 32 cols each)
    uint32_t z[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) z[j] = 0u;
    tmem_st16(trow + ((1 - slot) * 32) + half * 16, z);
  }
  sum += sum2.x + sum2.y;
  Xsum[half * 128 + row] = sum;
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();

  if (warp == 0 && elect_one()) {     // one elected lane, warp-uniform datapath for the TMA / tcgen05 issue
    trace_stamp(P.trace, 4);                                 // P published
    tc_fence_after();
    const uint64_t vd = make_smem_desc(smem_u32(Vs), 8 * kRowBytes, 8 * kRowBytes, kLayoutSw64);
    const uint32_t idesc = make_idesc_bf16(128, 32, 0, 1);   // B = V is MN-major
    for (int k = 0; k < kext / 16; ++k)
      mma_ts(tmem_base + kOCol, tmem_base + 8 * k, vd + (uint64_t)k * ((16 * kRowBytes) >> 4), idesc, k > 0);
    tc_commit(bar_o);
  }
  sum += Xsum[(half ^ 1) * 128 + row];

  // ---- LePE for my token, channels [16*half, 16*half+16), overlapped with the P.V MMA ----
  mbar_wait(bar_tma, 0);                                     // (already complete) acquire the TMA-written V tile
  const bool valid = slot < np && n < N;
  const int r = fast_div(n, ws, br.m_ws), c = n - r * ws;
  float2 lp[8];                                              // 16 channels as packed fp32 pairs (FFMA2)
  {
    const float2* bc = reinterpret_cast<const float2*>(Bc + min(slot, np - 1) * 32 + half * 16);
#pragma unroll
    for (int j = 0; j < 8; ++j) lp[j] = bc[j];
  }
  if (valid) {
    const uint32_t vbase = smem_u32(Vs);
    const uint32_t wbase = smem_u32(Wt) + (uint32_t)(slot * 9 * 32 + half * 16) * 2u;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int rr = r + t / 3 - 1, cc = c + t % 3 - 1;
      if (rr >= 0 && rr < hs && cc >= 0 && cc < ws) {
        const int vr = slot * slot_rows + rr * ws + cc;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {                     // 8 channels: one 16-byte V chunk x one 16-byte weight chunk, 8 FHFMA.BF16
          uint4 vv, ww;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(vv.x), "=r"(vv.y), "=r"(vv.z), "=r"(vv.w)
                       : "r"(v_chunk_addr(vbase, vr, half * 2 + ch)));
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(ww.x), "=r"(ww.y), "=r"(ww.z), "=r"(ww.w)
                       : "r"(wbase + (uint32_t)(t * 64 + ch * 16)));
          lp[ch * 4 + 0].x = fhfma_lo(vv.x, ww.x, lp[ch * 4 + 0].x); lp[ch * 4 + 0].y = fhfma_hi(vv.x, ww.x, lp[ch * 4 + 0].y);
          lp[ch * 4 + 1].x = fhfma_lo(vv.y, ww.y, lp[ch * 4 + 1].x); lp[ch * 4 + 1].y = fhfma_hi(vv.y, ww.y, lp[ch * 4 + 1].y);
          lp[ch * 4 + 2].x = fhfma_lo(vv.z, ww.z, lp[ch * 4 + 2].x); lp[ch * 4 + 2].y = fhfma_hi(vv.z, ww.z, lp[ch * 4 + 2].y);
          lp[ch * 4 + 3].x = fhfma_lo(vv.w, ww.w, lp[ch * 4 + 3].x); lp[ch * 4 + 3].y = fhfma_hi(vv.w, ww.w, lp[ch * 4 + 3].y);
        }
      }
    }
  }

  if (tid == 0) trace_stamp(P.trace, 5);                     // LePE done
  mbar_wait(bar_o, 0);
  if (tid == 0) trace_stamp(P.trace, 6);                     // O ready
  tc_fence_after();
  {
    uint32_t o[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]), "=r"(o[8]),
          "=r"(o[9]), "=r"(o[10]), "=r"(o[11]), "=r"(o[12]), "=r"(o[13]), "=r"(o[14]), "=r"(o[15])
        : "r"(trow + kOCol + half * 16) : "memory");
    tmem_wait_ld();
    if (valid) {
      const float inv = 1.0f / sum;
      const int64_t tok = (int64_t)(mih * hs + r) * P.reso + (miw * ws + c);
      __nv_bfloat16* dst = br.out + (int64_t)mb * br.o_bs + tok * br.o_ts + mhead * 32 + half * 16;
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 y = ffma2(make_float2(__uint_as_float(o[2 * j]), __uint_as_float(o[2 * j + 1])), make_float2(inv, inv), lp[j]);
        w[j] = pack_bf16x2(y.x, y.y);
      }
      *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<uint4*>(dst + 8) = make_uint4(w[4], w[5], w[6], w[7]);
      if (br.lse != nullptr && half == 0)
        br.lse[((int64_t)mb * P.reso * P.reso + tok) * br.heads + mhead] = mx * P.scale + logf(sum);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) trace_stamp(P.trace, 7);                     // exit
  if (warp == 0) tmem_dealloc(tmem_base, kWide ? 256u : kTmemCols);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int lepe_attention_fwd_tc(const cswin_lepe_branch_t* brs, int nb, int B, int reso, float scale, cudaStream_t stream,
                          bool* handled) {
  *handled = false;
  // eligibility: head_dim 32, window <= 128 tokens (or <= 256 with 128 % W_sp == 0: wide variant), TMA-compatible strides /
  // alignment; otherwise the SIMT kernel runs
  int n_wide = 0;
  for (int i = 0; i < nb; ++i) {
    const cswin_lepe_branch_t& s = brs[i];
    if (!s.q || !s.k || !s.v || !s.out || !s.conv_w || !s.conv_b) return CSWIN_OK;     // SIMT path reports the error
    if (s.heads <= 0 || s.C_b != s.heads * 32) return CSWIN_OK;
    if (s.H_sp <= 0 || s.W_sp <= 0 || reso % s.H_sp || reso % s.W_sp) return CSWIN_OK;
    if (s.H_sp * s.W_sp > 256 || s.H_sp > 256 || s.W_sp > 256) return CSWIN_OK;
    if (s.H_sp * s.W_sp > 128) { if (128 % s.W_sp) return CSWIN_OK; ++n_wide; }
    const int64_t strides[] = {s.q_bs, s.q_ts, s.k_bs, s.k_ts, s.v_bs, s.v_ts, s.o_bs, s.o_ts};
    for (int64_t st : strides) if (st <= 0 || (st * 2) % 16 != 0) return CSWIN_OK;
    if (!aligned16(s.q) || !aligned16(s.k) || !aligned16(s.v) || !aligned16(s.out) || !aligned16(s.conv_w)) return CSWIN_OK;
  }
  if (tc::encode_tiled_fn() == nullptr) return CSWIN_OK;
  if (n_wide != 0 && n_wide != nb) return CSWIN_OK;                  // one launch = one kernel variant
  const bool wide = n_wide != 0;

  TcParams P;
  P.nb = nb; P.reso = reso; P.scale = scale; P.scale_log2e = scale * 1.4426950408889634f;
  P.trace = g_trace.load(std::memory_order_relaxed);
  int tiles = 0;
  for (int i = 0; i < nb; ++i) {
    const cswin_lepe_branch_t& s = brs[i];
    TcBranch& d = P.br[i];
    d.out = (__nv_bfloat16*)s.out; d.cw = (const __nv_bfloat16*)s.conv_w; d.cb = (const __nv_bfloat16*)s.conv_b;
    d.lse = s.lse; d.o_bs = s.o_bs; d.o_ts = s.o_ts;
    d.heads = s.heads; d.hs = s.H_sp; d.ws = s.W_sp; d.nww = reso / s.W_sp;
    d.nwin = (reso / s.H_sp) * (reso / s.W_sp); d.N = s.H_sp * s.W_sp;
    d.nprob = B * d.nwin * d.heads;
    auto magic = [](int dv) { return dv > 1 ? (uint32_t)((0x100000000ull / (uint64_t)dv) + 1) : 0u; };
    d.m_heads = magic(d.heads); d.m_nwin = magic(d.nwin); d.m_nww = magic(d.nww); d.m_ws = magic(d.ws);
    d.tile_begin = tiles;
    const int slots = d.N <= 64 ? 2 : 1;
    tiles += wide ? 2 * d.nprob : (d.nprob + slots - 1) / slots;
    const void* ptr[3] = {s.q, s.k, s.v};
    const int64_t bs[3] = {s.q_bs, s.k_bs, s.v_bs}, ts[3] = {s.q_ts, s.k_ts, s.v_ts};
    for (int j = 0; j < 3; ++j) {
      const uint64_t dims[4] = {(uint64_t)s.C_b, (uint64_t)reso, (uint64_t)reso, (uint64_t)B};
      const uint64_t str[3] = {(uint64_t)ts[j] * 2, (uint64_t)ts[j] * 2 * reso, (uint64_t)bs[j] * 2};
      const uint32_t box[4] = {32, (uint32_t)s.W_sp, (uint32_t)((wide && j == 0) ? 128 / s.W_sp : s.H_sp), 1};   // wide q: one 128-row tile
      if (!tc::make_tensor_map_bf16(&P.map[i][j], ptr[j], 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B)) return CSWIN_ERR_CUDA;
    }
  }
  if (nb == 1) P.br[1] = P.br[0];
  static_assert(kSmemBytes <= 48 * 1024 && kSmemBytesWide <= 48 * 1024, "dynamic smem must stay under the no-opt-in limit");
  static const int forced_minb = [] { const char* e = getenv("CSWIN_ATTN_MINB"); return e ? atoi(e) : 0; }();     // A/B switch (3 / 4)
  const bool four = forced_minb ? forced_minb == 4 : tiles > 3 * sm_count();
  if (wide) CSWIN_CUDA_OK(launch_pdl(lepe_attn_fwd_tc_kernel<true, 2>, dim3(tiles), dim3(kThreads), (size_t)kSmemBytesWide, stream, P));
  else if (four) CSWIN_CUDA_OK(launch_pdl(lepe_attn_fwd_tc_kernel<false, 4>, dim3(tiles), dim3(kThreads), (size_t)kSmemBytes, stream, P));
  else CSWIN_CUDA_OK(launch_pdl(lepe_attn_fwd_tc_kernel<false, 3>, dim3(tiles), dim3(kThreads), (size_t)kSmemBytes, stream, P));
  CSWIN_LAUNCH_CHECK();
  g_tc_launches.fetch_add(1, std::memory_order_relaxed);
  *handled = true;
  return CSWIN_OK;
}

}  // namespace cswin
