// stage_tc.cu — a whole STAGE of CSWinBlocks as ONE persistent dataflow kernel (bf16, tcgen05 / TMEM / TMA, sm_100a).
//
// Replaces, for consecutive CSWinBlocks of one stage in inference (networks/cswin_unet.py:160-181 called from the stage
// loops :462-478 / :505-533), the five launches per block of the composed path
//     [LN1 + qkv Linear] -> [LePE attention, both branches] -> [proj + residual] -> [LN2 + fc1 + GELU] -> [fc2 + residual]
// by one launch for ALL blocks of the stage.  At cswin_tiny sizes every one of those launches is <= 2 waves and bound by
// its own latency chain (launch dependency, first TMA round trip, K loop, epilogue drain): 5 x ~7 us per block for ~5 us
// of tensor-core work.  Here the same tiles (128 x BN Linear tiles, 128-row attention tiles — same arithmetic, same
// epilogues, same operand order as gemm_tc.cu / attention_tc.cu) are executed by resident CTAs that pull tile indices from
// ONE global in-order counter and synchronise through per-row-tile / per-image completion counters in global memory:
//
//     tile order  : block 0 [qkv tiles | attention tiles | proj | fc1 | fc2], block 1 [...], ...   (row tile major inside an op)
//     dependencies: qkv(j, m)  <- fc2(j-1, m) all column chunks          attention(j, image) <- qkv(j, row tiles of the image)
//                   proj(j, m) <- attention(j, images overlapping m)     fc1(j, m) <- proj(j, m)      fc2(j, m) <- fc1(j, m)
//
// A tile is only handed out after every lower-numbered tile has been handed out, and a tile only waits for lower-numbered
// tiles, so the lowest unfinished tile is always being executed by a resident CTA: forward progress does not depend on how
// many CTAs are co-resident.  Row tile m of op k+1 starts as soon as row tile m of op k is complete — no grid-wide barrier, no
// launch gap, no wave tail; TMEM, barriers and tensor maps are set up once per CTA.  Activations move between ops through L2
// (everything of a stage is L2-resident); release = bulk-store completion + fence + red.release, acquire = ld.acquire +
// fence.proxy.async before the consumer's TMA loads.
//
// CTA = 320 threads, 2 CTAs per SM: warp 0 = tile scheduler + TMA producer, warp 1 = tcgen05.mma issuer, warps 2..9 =
// epilogue / softmax / LePE.  The operand ring (S stages of 16 KB A + <= 16 KB W) runs ACROSS tiles, the accumulator is
// double-buffered in TMEM (2 x 128 columns), so the loads and MMAs of tile t+1 overlap the epilogue of tile t; an attention
// tile takes one ring stage (Q | K | V boxes + LePE weights) and one accumulator slot (S / P / O).
#include <climits>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace cswin {
namespace {

using namespace tc;

constexpr int BM = 128, BK = 64;
constexpr int kThreads = 320;
constexpr int kMaxBlocks = 9;                 // CSWinBlocks per launch (kernel-parameter space: 640 B per block)
constexpr int kQ = 4;                         // tile-index queue between the scheduler and the MMA / epilogue roles
constexpr int kMaxStages = 4;
constexpr uint32_t kAccCols = 128;            // TMEM columns per accumulator slot (2 slots per CTA, 2 CTAs per SM = 512)
constexpr int kABytes = BM * BK * 2;          // 16 KB
constexpr int kStgBytes = 8 * 2048;           // epilogue staging: 8 warps x (32 rows x 64 B)
constexpr int kAttOperand = 128 * 64;         // Q / K / V box: 128 rows x 32 channels bf16
constexpr int kAttScratch = 3 * kAttOperand;  // Wt [2][9][32] f32, Bc [2][32] f32, Xmax [2][128] f32 behind the operands
constexpr int kAttStageBytes = kAttScratch + 2 * 9 * 32 * 4 + 2 * 32 * 4 + 2 * 128 * 4;
constexpr int kCtrlFlags = 64;                // ctrl[0] tile counter, ctrl[32] finished-CTA counter, flags from ctrl[64]

// -DCSWIN_STAGE_PROFILE: per-CTA cycle accounting of every wait (written to the debug trace buffer, tools/trace_stage.py)
#ifdef CSWIN_STAGE_PROFILE
#define PROF_DECL(...) uint32_t __VA_ARGS__
#define PROF_T0(t) const long long t = clock64()
#define PROF_ACC(t, acc) acc += (uint32_t)(clock64() - t)
#define PROF_OUT(slot, v) do { if (P.trace != nullptr && !P.tile_trace && blockIdx.x < 1024) P.trace[(size_t)blockIdx.x * 16 + (slot)] = (v); } while (0)
// CSWIN_STAGE_TILE_TRACE=1: the trace buffer is (total tiles, 10) uint64 and receives %globaltimer stamps per TILE instead
#define TSTAMP(t, slot) do { if (P.tile_trace) { unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); P.trace[(size_t)(t) * 10 + (slot)] = gt_; } } while (0)
#define TVAL(t, slot, v) do { if (P.tile_trace) P.trace[(size_t)(t) * 10 + (slot)] = (v); } while (0)
#else
#define TSTAMP(t, slot)
#define TVAL(t, slot, v)
#define PROF_DECL(...)
#define PROF_T0(t)
#define PROF_ACC(t, acc)
#define PROF_OUT(slot, v)
#endif

enum { OP_QKV = 0, OP_ATT = 1, OP_PROJ = 2, OP_FC1 = 3, OP_FC2 = 4, N_OPS = 5 };

struct GemmOp { int N, nkb, BN, nch, act, fold, has_res, pad; };
struct AttBranch { int heads, hs, ws, nww, nwin, N, slots, tiles_img, nprob_img, ch0, pad0, pad1; };

struct alignas(64) BlockW {
  CUtensorMap w[4];                                   // qkv, proj, fc1, fc2 weights (N, K) row-major, box {64, BN}
  const float* bias[4];                               // fp32 biases (qkv / fc1: b + W beta of the folded LayerNorm)
  const float* cs[2];                                 // fp32 column sums of W o gamma: qkv, fc1
  const __nv_bfloat16* cw[2]; const __nv_bfloat16* cb[2];   // LePE depthwise conv per branch
  float eps1, eps2;
};

struct alignas(64) StageParams {
  CUtensorMap a_map[4];                               // A operands {64, 128} SW128: x (qkv), att (proj), x1 (fc1), hid (fc2)
  CUtensorMap o_map[4];                               // outputs {32, 32} SW64: qkv, x1, hid, x
  CUtensorMap qkv4[2][3];                             // attention gathers (32 ch, W_sp, H_sp, 1) of q / k / v per branch
  BlockW blk[kMaxBlocks];
  GemmOp op[4];
  AttBranch br[2];
  const __nv_bfloat16* x; const __nv_bfloat16* x1; __nv_bfloat16* att;
  const float* stats_in; float* stats_x; float* stats_x1;
  int* ctrl;
  int parts_in, nb, nblk, M, L, B, reso, C;
  int mt, U, tpb, total, att_tiles_img, S, stage_bytes, pad;
  int t0[N_OPS + 1];                                  // first tile of each op inside a block (block order), t0[5] = tiles per block
  float scale, scale_log2e, invC, pad2;
  unsigned long long* trace; int tile_trace, pad3;
};

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bar_epi() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ float ld_cg_f32(const float* p) { float v; asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint4 ld_cg_u4(const void* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t v_chunk_addr(uint32_t vbase, int row, int chunk) {
  return vbase + row * 64 + (((chunk ^ (row >> 1)) & 3) << 4);          // Swizzle<2,4,3> (64-byte swizzle)
}

// bounded spin on a completion counter (a protocol bug becomes a trap after ~4 s, not a hang)
__device__ __forceinline__ void wait_flag(const int* p, int need) {
  if (ld_acquire(p) >= need) return;
  uint64_t t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  uint32_t spins = 0;
  while (ld_acquire(p) < need) {
    if ((++spins & 0xff) == 0) {
      uint64_t t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 4000000000ull) __trap();
    }
  }
}

struct Tile { int j, op, m, n; };                      // block, op (block order), row tile / image, column chunk / tile in image

__device__ __forceinline__ Tile decode(const StageParams& P, int t) {
  Tile T;
  T.j = t / P.tpb;
  const int r = t - T.j * P.tpb;
  T.op = r < P.t0[1] ? 0 : r < P.t0[2] ? 1 : r < P.t0[3] ? 2 : r < P.t0[4] ? 3 : 4;
  const int rr = r - P.t0[T.op];
  const int per = T.op == OP_ATT ? P.att_tiles_img : P.op[T.op == 0 ? 0 : T.op - 1].nch;
  T.m = rr / per;
  T.n = rr - T.m * per;
  return T;
}
__device__ __forceinline__ int gemm_index(int op) { return op == 0 ? 0 : op - 1; }   // OP_* -> index into op[] / w[] / bias[]

// the completion counters tile T waits for: flags [lo, hi] of (block fj, op fo) must reach `need`
struct Dep { const int* f; int lo, hi, need; };
__device__ __forceinline__ Dep deps_of(const StageParams& P, const Tile& T) {
  Dep d; d.f = nullptr; d.lo = 0; d.hi = -1; d.need = 0;
  int fj = T.j, fo = 0;
  if (T.op == OP_QKV) {
    if (T.j == 0) return d;
    fj = T.j - 1; fo = OP_FC2; d.lo = d.hi = T.m; d.need = P.op[3].nch;
  } else if (T.op == OP_ATT) {
    fo = OP_QKV; d.lo = (T.m * P.L) / BM; d.hi = (T.m * P.L + P.L - 1) / BM; d.need = P.op[0].nch;
  } else if (T.op == OP_PROJ) {
    fo = OP_ATT; d.lo = (T.m * BM) / P.L; d.hi = min(P.M - 1, T.m * BM + BM - 1) / P.L; d.need = P.att_tiles_img;
  } else {
    fo = T.op - 1; d.lo = d.hi = T.m; d.need = P.op[gemm_index(fo)].nch;
  }
  d.f = P.ctrl + kCtrlFlags + (fj * N_OPS + fo) * P.U;
  return d;
}
__device__ __forceinline__ void wait_deps(const Dep& d) {
  for (int u = d.lo; u <= d.hi; ++u) wait_flag(d.f + u, d.need);
}

__global__ void __launch_bounds__(kThreads, 2) stage_tc_kernel(const __grid_constant__ StageParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int S = P.S;
  uint8_t* ring = smem;                                             // [S][stage_bytes]: A at +0, W at +16 KB | Q, K, V, scratch
  uint8_t* Stg = ring + (size_t)S * P.stage_bytes;                  // [8][2048] epilogue staging boxes
  float* sBias = reinterpret_cast<float*>(Stg + kStgBytes);         // [128]
  float* sCs = sBias + 128;                                         // [128]
  float* sStat = sCs + 128;                                         // [128][2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + 256);
  // barriers: full[S], empty[S], acc_full[2], acc_empty[2], p_ready[2], o_full[2], tq_full[kQ], tq_empty[kQ]
  int* tq = reinterpret_cast<int*>(bars + 2 * kMaxStages + 8 + 2 * kQ);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tq + kQ);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  auto full = [&](int s) { return smem_u32(&bars[s]); };
  auto empty = [&](int s) { return smem_u32(&bars[kMaxStages + s]); };
  auto acc_full = [&](int a) { return smem_u32(&bars[2 * kMaxStages + a]); };
  auto acc_empty = [&](int a) { return smem_u32(&bars[2 * kMaxStages + 2 + a]); };
  auto p_ready = [&](int a) { return smem_u32(&bars[2 * kMaxStages + 4 + a]); };
  auto o_full = [&](int a) { return smem_u32(&bars[2 * kMaxStages + 6 + a]); };
  auto tq_full = [&](int i) { return smem_u32(&bars[2 * kMaxStages + 8 + i]); };
  auto tq_empty = [&](int i) { return smem_u32(&bars[2 * kMaxStages + 8 + kQ + i]); };

  if (warp == 0 && elect_one()) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(acc_full(a), 1); mbar_init(acc_empty(a), 1); mbar_init(p_ready(a), 1); mbar_init(o_full(a), 1); }
    for (int i = 0; i < kQ; ++i) { mbar_init(tq_full(i), 1); mbar_init(tq_empty(i), 9); }   // consumers: MMA thread + 8 epilogue warps
    fence_barrier_init();
    fence_proxy_async();
    for (int i = 0; i < 4; ++i) { tma_prefetch_desc(&P.a_map[i]); tma_prefetch_desc(&P.blk[0].w[i]); }
  }
  if (warp == 1) { tmem_alloc(smem_u32(tmem_slot), 2 * kAccCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                   // the counters, the activations and the workspaces are safe from here

  if (warp == 0) {
    // =============================== scheduler + TMA producer ===============================
    if (elect_one()) {
      uint32_t rc = 0;                          // ring stages produced so far
      int qi = 0; uint32_t qph = 1;
      PROF_DECL(p_dep = 0, p_empty = 0, p_tq = 0, p_tiles = 0);
      PROF_T0(p_start);
      int next = atomicAdd(P.ctrl, 1);
      for (;;) {
        PROF_T0(tq0);
        const int t = next;
        if (t < P.total) next = atomicAdd(P.ctrl, 1);      // in flight while this tile's loads are issued
        mbar_wait(tq_empty(qi), qph);
        PROF_ACC(tq0, p_tq);
        tq[qi] = t < P.total ? t : -1;
        mbar_arrive(tq_full(qi));
        if (++qi == kQ) { qi = 0; qph ^= 1; }
        if (t >= P.total) break;
        const Tile T = decode(P, t);
        const Dep d = deps_of(P, T);
        TSTAMP(t, 0); TVAL(t, 8, (unsigned long long)blockIdx.x); TVAL(t, 9, (unsigned long long)(T.j | (T.op << 8) | (T.m << 12) | ((unsigned long long)T.n << 32)));
        if (T.op == OP_ATT) {
          const AttBranch& br = P.br[(P.nb > 1 && T.n >= P.br[0].tiles_img) ? 1 : 0];
          const int bi = (P.nb > 1 && T.n >= P.br[0].tiles_img) ? 1 : 0;
          const int p0 = (T.n - (bi ? P.br[0].tiles_img : 0)) * br.slots;
          const int np = min(br.slots, br.nprob_img - p0);
          const int s = rc % S; const uint32_t ph = (rc / S) & 1;
          PROF_T0(e0);
          mbar_wait(empty(s), ph ^ 1);
          PROF_ACC(e0, p_empty);
          PROF_T0(d0);
          wait_deps(d);
          fence_proxy_async_all();
          PROF_ACC(d0, p_dep);
          TSTAMP(t, 1);
          mbar_expect_tx(full(s), (uint32_t)(np * 3 * br.N * 64));
          const uint32_t base = smem_u32(ring + (size_t)s * P.stage_bytes);
          const int slot_rows = BM / br.slots;
          for (int sl = 0; sl < np; ++sl) {
            const int local = p0 + sl;
            const int hd = local % br.heads, win = local / br.heads;
            const int ih = win / br.nww, iw = win - ih * br.nww;
            const uint32_t off = sl * slot_rows * 64;
            tma_load_4d(base + off, &P.qkv4[bi][0], full(s), hd * 32, iw * br.ws, ih * br.hs, T.m);
            tma_load_4d(base + kAttOperand + off, &P.qkv4[bi][1], full(s), hd * 32, iw * br.ws, ih * br.hs, T.m);
            tma_load_4d(base + 2 * kAttOperand + off, &P.qkv4[bi][2], full(s), hd * 32, iw * br.ws, ih * br.hs, T.m);
          }
          ++rc;
        } else {
          const int g = gemm_index(T.op);
          const GemmOp& op = P.op[g];
          const uint32_t w_bytes = (uint32_t)op.BN * BK * 2;
          bool waited = d.f == nullptr;
          for (int kb = 0; kb < op.nkb; ++kb, ++rc) {
            const int s = rc % S; const uint32_t ph = (rc / S) & 1;
            PROF_T0(e0);
            mbar_wait(empty(s), ph ^ 1);
            PROF_ACC(e0, p_empty);
            const uint32_t base = smem_u32(ring + (size_t)s * P.stage_bytes);
            mbar_expect_tx(full(s), kABytes + w_bytes);
            tma_load_2d(base + kABytes, &P.blk[T.j].w[g], full(s), kb * BK, T.n * op.BN);     // weights never wait
            if (!waited) { PROF_T0(d0); wait_deps(d); fence_proxy_async_all(); waited = true; PROF_ACC(d0, p_dep); TSTAMP(t, 1); }
            tma_load_2d(base, &P.a_map[g], full(s), kb * BK, T.m * BM);
          }
        }
        TSTAMP(t, 2);
#ifdef CSWIN_STAGE_PROFILE
        ++p_tiles;
#endif
      }
#ifdef CSWIN_STAGE_PROFILE
      PROF_OUT(0, (unsigned long long)(clock64() - p_start)); PROF_OUT(1, p_dep); PROF_OUT(2, p_empty); PROF_OUT(3, p_tq); PROF_OUT(15, p_tiles);
#endif
    }
  } else if (warp == 1) {
    // =============================== tcgen05.mma issuer ===============================
    if (elect_one()) {
      uint32_t rc = 0, seq = 0, att_par = 0;    // att_par: bit a = parity of p_ready / o_full of accumulator slot a
      int qi = 0; uint32_t qph = 0;
      PROF_DECL(m_full = 0, m_acc = 0, m_p = 0, m_tq = 0);
      PROF_T0(m_start);
      for (;;) {
        PROF_T0(q0);
        mbar_wait(tq_full(qi), qph);
        PROF_ACC(q0, m_tq);
        const int t = tq[qi];
        mbar_arrive(tq_empty(qi));
        if (++qi == kQ) { qi = 0; qph ^= 1; }
        if (t < 0) break;
        const Tile T = decode(P, t);
        const int a = seq & 1; const uint32_t aph = (seq >> 1) & 1;
        ++seq;
        const uint32_t acc = tmem_base + a * kAccCols;
        PROF_T0(a0);
        mbar_wait(acc_empty(a), aph ^ 1);
        PROF_ACC(a0, m_acc);
        tc_fence_after();
        if (T.op == OP_ATT) {
          const int bi = (P.nb > 1 && T.n >= P.br[0].tiles_img) ? 1 : 0;
          const AttBranch& br = P.br[bi];
          const int kext = br.slots == 2 ? 128 : ((br.N + 15) & ~15);
          const int s = rc % S; const uint32_t ph = (rc / S) & 1;
          ++rc;
          const uint32_t base = smem_u32(ring + (size_t)s * P.stage_bytes);
          PROF_T0(f0);
          mbar_wait(full(s), ph);
          PROF_ACC(f0, m_full);
          TSTAMP(t, 3);
          tc_fence_after();
          const uint64_t qd = make_smem_desc(base, 16, 8 * 64, kLayoutSw64);
          const uint64_t kd = make_smem_desc(base + kAttOperand, 16, 8 * 64, kLayoutSw64);
          const uint32_t idesc = make_idesc_bf16(128, kext, 0, 0);
          mma_ss(acc, qd, kd, idesc, false);
          mma_ss(acc, qd + 2, kd + 2, idesc, true);
          tc_commit(acc_full(a));                                  // S ready
          PROF_T0(p0t);
          mbar_wait(p_ready(a), (att_par >> a) & 1);                   // P (bf16) is in TMEM, padded V rows are zero
          PROF_ACC(p0t, m_p);
          tc_fence_after();
          const uint64_t vd = make_smem_desc(base + 2 * kAttOperand, 8 * 64, 8 * 64, kLayoutSw64);
          const uint32_t idesc2 = make_idesc_bf16(128, 32, 0, 1);  // B = V is MN-major
          for (int k = 0; k < kext / 16; ++k)
            mma_ts(acc + 64, acc + 8 * k, vd + (uint64_t)k * ((16 * 64) >> 4), idesc2, k > 0);
          tc_commit(o_full(a));
          TSTAMP(t, 4);
          att_par ^= 1u << a;
        } else {
          const GemmOp& op = P.op[gemm_index(T.op)];
          const uint32_t idesc = make_idesc_bf16(BM, op.BN, 0, 0);
          for (int kb = 0; kb < op.nkb; ++kb, ++rc) {
            const int s = rc % S; const uint32_t ph = (rc / S) & 1;
            PROF_T0(f0);
            mbar_wait(full(s), ph);
            PROF_ACC(f0, m_full);
            if (kb == 0) TSTAMP(t, 3);
            tc_fence_after();
            const uint32_t base = smem_u32(ring + (size_t)s * P.stage_bytes);
            const uint64_t ad = make_smem_desc(base, 16, 1024, kLayoutSw128);
            const uint64_t wd = make_smem_desc(base + kABytes, 16, 1024, kLayoutSw128);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) mma_ss(acc, ad + 2 * k, wd + 2 * k, idesc, (kb | k) != 0);
            tc_commit(empty(s));
          }
          tc_commit(acc_full(a));
          TSTAMP(t, 4);
        }
      }
#ifdef CSWIN_STAGE_PROFILE
      PROF_OUT(4, (unsigned long long)(clock64() - m_start)); PROF_OUT(5, m_full); PROF_OUT(6, m_acc); PROF_OUT(7, m_p); PROF_OUT(14, m_tq);
#endif
    }
  } else {
    // =============================== epilogue / softmax / LePE: warps 2..9 ===============================
    const int ctid = tid - 64;                  // 0..255
    const int q = warp & 3;                     // TMEM lane quadrant this warp may access (hardware rule: warp id % 4)
    const int half = (warp - 2) >> 2;           // which of the two warps of a quadrant
    uint32_t rc = 0, seq = 0, att_par = 0;
    int qi = 0; uint32_t qph = 0;
    uint8_t* stg = Stg + (warp - 2) * 2048;
    const uint32_t stg_u32 = smem_u32(stg);
    PROF_DECL(e_tq = 0, e_acc = 0, e_dep = 0, e_store = 0, e_rel = 0, e_att = 0, e_gemm = 0);
    PROF_T0(e_start);
    for (;;) {
      PROF_T0(q0);
      mbar_wait(tq_full(qi), qph);
      PROF_ACC(q0, e_tq);
      PROF_T0(tile0);
      const int t = tq[qi];
      __syncwarp();
      if (lane == 0) mbar_arrive(tq_empty(qi));
      if (++qi == kQ) { qi = 0; qph ^= 1; }
      if (t < 0) break;
      const Tile T = decode(P, t);
      const int a = seq & 1; const uint32_t aph = (seq >> 1) & 1;
      ++seq;
      const uint32_t acc = tmem_base + a * kAccCols;
      const uint32_t trow = acc + ((uint32_t)(q * 32) << 16);
      int* my_flag = P.ctrl + kCtrlFlags + (T.j * N_OPS + T.op) * P.U + T.m;

      if (T.op == OP_ATT) {
        // ---------------- one attention tile: 128 query rows (1 or 2 (window, head) problems), as attention_tc.cu ----------------
        const int bi = (P.nb > 1 && T.n >= P.br[0].tiles_img) ? 1 : 0;
        const AttBranch& br = P.br[bi];
        const int N = br.N, hs = br.hs, ws = br.ws, slots = br.slots;
        const int slot_rows = BM / slots;
        const int p0 = (T.n - (bi ? P.br[0].tiles_img : 0)) * slots;
        const int np = min(slots, br.nprob_img - p0);
        const int kext = slots == 2 ? 128 : ((N + 15) & ~15);
        const int row = q * 32 + lane;          // tile row == TMEM lane
        const int slot = row / slot_rows;       // warp-uniform
        const int n = row - slot * slot_rows;   // token inside the window
        int mih, miw, mhead;
        {
          const int local = p0 + min(slot, np - 1);
          mhead = local % br.heads;
          const int win = local / br.heads;
          mih = win / br.nww; miw = win - mih * br.nww;
        }
        const int s = rc % S; const uint32_t ph = (rc / S) & 1;
        ++rc;
        uint8_t* st = ring + (size_t)s * P.stage_bytes;
        uint8_t* Vs = st + 2 * kAttOperand;
        float* Wt = reinterpret_cast<float*>(st + kAttScratch);     // [2][9][32]
        float* Bc = Wt + 2 * 9 * 32;                                // [2][32]
        float* Xmax = Bc + 2 * 32;                                  // [2][128]
        float* Xsum = Xmax;
        mbar_wait(full(s), ph);                                     // q, k, v landed; the stage (and its scratch) is this tile's
        // zero the V rows the P.V MMA reads but TMA did not write (0 * stale-NaN would poison O)
        for (int i = ctid; i < 128 * 4; i += 256) {
          const int r = i >> 2;
          const int sl = r / slot_rows, rn = r - sl * slot_rows;
          if (r < kext && (sl >= np || rn >= N)) *reinterpret_cast<uint4*>(Vs + i * 16) = make_uint4(0, 0, 0, 0);
        }
        if (ctid < np * 36) {                                       // LePE weights of the head(s): Wt[slot][tap][ch] fp32
          const int sl = ctid / 36, i = ctid - sl * 36;
          const int hd = (p0 + sl) % br.heads;
          const uint4 raw = *reinterpret_cast<const uint4*>(P.blk[T.j].cw[bi] + (size_t)hd * 288 + i * 8);
          const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int idx = i * 8 + e;
            const int ch = idx / 9, tp = idx - ch * 9;
            Wt[(sl * 9 + tp) * 32 + ch] = (e & 1) ? bf16_hi(w4[e >> 1]) : bf16_lo(w4[e >> 1]);
          }
        } else if (ctid >= 128 && ctid < 128 + np * 32) {
          const int sl = (ctid - 128) >> 5, ch = ctid & 31;
          const int hd = (p0 + sl) % br.heads;
          Bc[sl * 32 + ch] = __bfloat162float(P.blk[T.j].cb[bi][hd * 32 + ch]);
        }
        fence_proxy_async();
        PROF_T0(w0);
        mbar_wait(acc_full(a), aph);                                // S ready
        PROF_ACC(w0, e_acc);
        if (ctid == 0) TSTAMP(t, 5);
        tc_fence_after();

        const int hcols = slot_rows >> 1;                           // 32 (two problems) or 64 (one problem)
        const int kbeg = half * hcols;
        const int cbeg = slot * slot_rows + kbeg;
        const int nch = hcols >> 5;
        float mx = -INFINITY;
        for (int c = 0; c < nch; ++c) {
          if (kbeg + 32 * c >= kext) break;
          uint32_t v[32];
          tmem_ld32(trow + cbeg + 32 * c, v);
          tmem_wait_ld();
          const int lim = N - (kbeg + 32 * c);
          if (lim >= 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < lim) mx = fmaxf(mx, __uint_as_float(v[j]));
          }
        }
        Xmax[half * 128 + row] = mx;
        bar_epi();
        mx = fmaxf(mx, Xmax[(half ^ 1) * 128 + row]);
        const float mxs = mx * P.scale_log2e;
        float sum = 0.f;
        float2 sum2 = make_float2(0.f, 0.f);
        const float2 sl2 = make_float2(P.scale_log2e, P.scale_log2e), nmxs2 = make_float2(-mxs, -mxs);
        uint32_t pk[2][16];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (c < nch && kbeg + 32 * c < kext) {
            uint32_t v[32];
            tmem_ld32(trow + cbeg + 32 * c, v);
            tmem_wait_ld();
            const int lim = N - (kbeg + 32 * c);
            if (lim >= 32) {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const float2 tt = ffma2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sl2, nmxs2);
                const float2 e = make_float2(ex2_approx(tt.x), ex2_approx(tt.y));
                sum2 = fadd2(sum2, e);
                pk[c][j >> 1] = pack_bf16x2(e.x, e.y);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const float e0 = (j < lim) ? ex2_approx(fmaf(__uint_as_float(v[j]), P.scale_log2e, -mxs)) : 0.f;
                const float e1 = (j + 1 < lim) ? ex2_approx(fmaf(__uint_as_float(v[j + 1]), P.scale_log2e, -mxs)) : 0.f;
                sum += e0 + e1;
                pk[c][j >> 1] = pack_bf16x2(e0, e1);
              }
            }
          }
        }
        bar_epi();                                                  // every S value is in registers: P may overwrite S
#pragma unroll
        for (int c = 0; c < 2; ++c)
          if (c < nch && kbeg + 32 * c < kext) tmem_st16(trow + ((cbeg + 32 * c) >> 1), pk[c]);
        if (slots == 2) {                                           // keys of the other problem: P = 0
          uint32_t z[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) z[j] = 0u;
          tmem_st16(trow + ((1 - slot) * 32) + half * 16, z);
        }
        sum += sum2.x + sum2.y;
        Xsum[half * 128 + row] = sum;
        tmem_wait_st();
        tc_fence_before();
        bar_epi();
        if (ctid == 0) mbar_arrive(p_ready(a));
        sum += Xsum[(half ^ 1) * 128 + row];

        // LePE for my token, channels [16*half, +16), overlapped with the P.V MMA
        const bool valid = slot < np && n < N;
        const int r = n / ws, c = n - r * ws;
        float2 lp[8];
        {
          const float2* bc = reinterpret_cast<const float2*>(Bc + min(slot, np - 1) * 32 + half * 16);
#pragma unroll
          for (int j = 0; j < 8; ++j) lp[j] = bc[j];
        }
        if (valid) {
          const uint32_t vbase = smem_u32(Vs);
          const float* wt = Wt + slot * 9 * 32 + half * 16;
#pragma unroll
          for (int tp = 0; tp < 9; ++tp) {
            const int rr = r + tp / 3 - 1, cc = c + tp % 3 - 1;
            if (rr >= 0 && rr < hs && cc >= 0 && cc < ws) {
              const int vr = slot * slot_rows + rr * ws + cc;
#pragma unroll
              for (int ch = 0; ch < 2; ++ch) {
                uint4 vv;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(vv.x), "=r"(vv.y), "=r"(vv.z), "=r"(vv.w)
                             : "r"(v_chunk_addr(vbase, vr, half * 2 + ch)));
                const float4 w0 = *reinterpret_cast<const float4*>(wt + tp * 32 + ch * 8);
                const float4 w1 = *reinterpret_cast<const float4*>(wt + tp * 32 + ch * 8 + 4);
                lp[ch * 4 + 0] = ffma2(make_float2(w0.x, w0.y), make_float2(bf16_lo(vv.x), bf16_hi(vv.x)), lp[ch * 4 + 0]);
                lp[ch * 4 + 1] = ffma2(make_float2(w0.z, w0.w), make_float2(bf16_lo(vv.y), bf16_hi(vv.y)), lp[ch * 4 + 1]);
                lp[ch * 4 + 2] = ffma2(make_float2(w1.x, w1.y), make_float2(bf16_lo(vv.z), bf16_hi(vv.z)), lp[ch * 4 + 2]);
                lp[ch * 4 + 3] = ffma2(make_float2(w1.z, w1.w), make_float2(bf16_lo(vv.w), bf16_hi(vv.w)), lp[ch * 4 + 3]);
              }
            }
          }
        }
        mbar_wait(o_full(a), (att_par >> a) & 1);
        att_par ^= 1u << a;
        tc_fence_after();
        {
          uint32_t o[16];
          tmem_ld16(trow + 64 + half * 16, o);
          tmem_wait_ld();
          if (valid) {
            const float inv = 1.0f / sum;
            const int64_t tok = (int64_t)(mih * hs + r) * P.reso + (miw * ws + c);
            __nv_bfloat16* dst = P.att + ((int64_t)T.m * P.L + tok) * P.C + br.ch0 + mhead * 32 + half * 16;
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
        tc_fence_before();
        bar_epi();                               // all stores issued, all TMEM / smem reads of this tile done
        if (ctid == 0) {
          mbar_arrive(empty(s));
          mbar_arrive(acc_empty(a));
          TSTAMP(t, 6);
          PROF_T0(r0);
          red_release_add(my_flag, 1);           // release: orders the tile's global stores (all warps, via the barrier above) before the count
          PROF_ACC(r0, e_rel);
          TSTAMP(t, 7);
        }
        PROF_ACC(tile0, e_att);
      } else {
        // ---------------- epilogue of one 128 x BN Linear tile (the TMA-store fast path of gemm_tc.cu) ----------------
        const int g = gemm_index(T.op);
        const GemmOp& op = P.op[g];
        const BlockW& W = P.blk[T.j];
        const int BN = op.BN, n0 = T.n * BN;
        const int64_t m0 = (int64_t)T.m * BM;
        const int nunits = BN >> 5;
        rc += op.nkb;
        // my own acquire of what this tile reads with plain loads (row statistics, residual rows)
        {
          const Dep d = deps_of(P, T);
          PROF_T0(d0);
          if (d.f != nullptr) { if (lane == 0) wait_deps(d); __syncwarp(); }
          PROF_ACC(d0, e_dep);
        }
        bar_epi();                               // the previous tile's epilogue no longer reads sBias / sCs / sStat
        if (ctid < BN) {
          sBias[ctid] = W.bias[g][n0 + ctid];
          if (op.fold) sCs[ctid] = W.cs[g == 0 ? 0 : 1][n0 + ctid];
        }
        sStat[ctid] = 0.f;
        const int64_t mrow = m0 + q * 32 + lane;
        float ln_mean = 0.f, ln_rstd = 1.f;
        if (op.fold && mrow < P.M) {
          const float* stp; int parts; float eps;
          if (g == 0) { stp = (T.j == 0) ? P.stats_in : P.stats_x; parts = (T.j == 0) ? P.parts_in : P.op[3].nch; eps = W.eps1; }
          else { stp = P.stats_x1; parts = P.op[1].nch; eps = W.eps2; }
          float s1 = 0.f, s2 = 0.f;
          for (int p = 0; p < parts; ++p) { s1 += ld_cg_f32(stp + (mrow * parts + p) * 2); s2 += ld_cg_f32(stp + (mrow * parts + p) * 2 + 1); }
          ln_mean = s1 * P.invC;
          ln_rstd = rsqrtf(fmaxf(fmaf(-ln_mean, ln_mean, s2 * P.invC), 0.f) + eps);
        }
        const __nv_bfloat16* res = op.has_res ? (g == 1 ? P.x : P.x1) : nullptr;
        const bool has_res = res != nullptr && mrow < P.M;
        const bool stats = op.has_res != 0;      // proj and fc2 emit the row statistics of their output for the next folded Linear
        bar_epi();
        PROF_T0(w0);
        mbar_wait(acc_full(a), aph);
        PROF_ACC(w0, e_acc);
        if (ctid == 0) TSTAMP(t, 5);
        tc_fence_after();
        for (int u = half; u < nunits; u += 2) {
          uint32_t v[32];
          tmem_ld32(trow + u * 32, v);
          uint4 rv[4];
          const int ncol = n0 + u * 32;
          if (has_res) {
#pragma unroll
            for (int c = 0; c < 4; ++c) rv[c] = ld_cg_u4(res + mrow * P.C + ncol + c * 8);
          }
          tmem_wait_ld();
          const float4* b4 = reinterpret_cast<const float4*>(sBias + u * 32);
          const float4* c4 = reinterpret_cast<const float4*>(sCs + u * 32);
          if (u != half) {
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
          }
          float st1 = 0.f, st2 = 0.f;
          const float2 nmean2 = make_float2(-ln_mean, -ln_mean), rstd2 = make_float2(ln_rstd, ln_rstd);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float2 f[4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float4 bb = b4[c * 2 + h];
              const float2 a0 = make_float2(__uint_as_float(v[c * 8 + h * 4 + 0]), __uint_as_float(v[c * 8 + h * 4 + 1]));
              const float2 a1 = make_float2(__uint_as_float(v[c * 8 + h * 4 + 2]), __uint_as_float(v[c * 8 + h * 4 + 3]));
              if (op.fold) {
                const float4 cc = c4[c * 2 + h];
                f[h * 2 + 0] = ffma2(rstd2, ffma2(nmean2, make_float2(cc.x, cc.y), a0), make_float2(bb.x, bb.y));
                f[h * 2 + 1] = ffma2(rstd2, ffma2(nmean2, make_float2(cc.z, cc.w), a1), make_float2(bb.z, bb.w));
              } else {
                f[h * 2 + 0] = fadd2(a0, make_float2(bb.x, bb.y));
                f[h * 2 + 1] = fadd2(a1, make_float2(bb.z, bb.w));
              }
            }
            if (op.act == 1) {
#pragma unroll
              for (int e = 0; e < 4; ++e) f[e] = gelu_fast2(f[e]);
            }
            uint4 x = make_uint4(pack_bf16x2(f[0].x, f[0].y), pack_bf16x2(f[1].x, f[1].y), pack_bf16x2(f[2].x, f[2].y), pack_bf16x2(f[3].x, f[3].y));
            if (has_res) {
              x.x = add_bf16x2(x.x, rv[c].x); x.y = add_bf16x2(x.y, rv[c].y);
              x.z = add_bf16x2(x.z, rv[c].z); x.w = add_bf16x2(x.w, rv[c].w);
            }
            if (stats) {
              const float e0 = bf16_lo(x.x), e1 = bf16_hi(x.x), e2 = bf16_lo(x.y), e3 = bf16_hi(x.y);
              const float e4 = bf16_lo(x.z), e5 = bf16_hi(x.z), e6 = bf16_lo(x.w), e7 = bf16_hi(x.w);
              st1 += ((e0 + e1) + (e2 + e3)) + ((e4 + e5) + (e6 + e7));
              st2 = fmaf(e0, e0, fmaf(e1, e1, fmaf(e2, e2, fmaf(e3, e3, fmaf(e4, e4, fmaf(e5, e5, fmaf(e6, e6, fmaf(e7, e7, st2))))))));
            }
            const uint32_t addr = stg_u32 + lane * 64 + (((c ^ (lane >> 1)) & 3) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w) : "memory");
          }
          if (stats) { atomicAdd(&sStat[(q * 32 + lane) * 2], st1); atomicAdd(&sStat[(q * 32 + lane) * 2 + 1], st2); }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&P.o_map[g], stg_u32, ncol, (int)(m0 + q * 32));
            tma_store_commit();
          }
        }
        tc_fence_before();
        PROF_T0(s0);
        if (lane == 0) tma_store_wait_all();     // this warp's output boxes are written (not merely read out of shared memory)
        __syncwarp();
        bar_epi();                               // every warp: accumulator read, statistics accumulated, stores complete
        PROF_ACC(s0, e_store);
        if (stats && ctid < BM && m0 + ctid < P.M) {
          float* dst = (g == 1 ? P.stats_x1 : P.stats_x) + ((m0 + ctid) * op.nch + T.n) * 2;
          dst[0] = sStat[ctid * 2]; dst[1] = sStat[ctid * 2 + 1];
        }
        if (ctid == 0) mbar_arrive(acc_empty(a));
        if (stats) bar_epi();
        if (ctid == 0) {
          TSTAMP(t, 6);
          PROF_T0(r0);
          red_release_add(my_flag, 1);           // the bulk stores of every warp have completed (wait_group 0 + barrier): publish
          PROF_ACC(r0, e_rel);
          TSTAMP(t, 7);
        }
        PROF_ACC(tile0, e_gemm);
      }
    }
#ifdef CSWIN_STAGE_PROFILE
    if (ctid == 0) {
      PROF_OUT(8, (unsigned long long)(clock64() - e_start)); PROF_OUT(9, e_tq); PROF_OUT(10, e_acc); PROF_OUT(11, e_dep); PROF_OUT(12, e_store);
      PROF_OUT(13, e_rel | ((unsigned long long)e_att << 32));
    }
    (void)e_gemm;
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * kAccCols);
  // the last CTA to finish clears the counters for the next launch (which reads them only after griddepcontrol.wait)
  __shared__ int s_last;
  if (tid == 0) {
    __threadfence();
    s_last = atomicAdd(P.ctrl + 32, 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {
    const int nflags = P.nblk * N_OPS * P.U;
    for (int i = tid; i < nflags; i += kThreads) P.ctrl[kCtrlFlags + i] = 0;
    __syncthreads();
    if (tid == 0) { P.ctrl[0] = 0; __threadfence(); P.ctrl[32] = 0; }
  }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }

// tile width of each Linear: a multiple of 32 that divides N, <= 128 (one accumulator slot) and <= the W half of a ring stage
int pick_bn(int N, int cap) {
  for (int bn = cap; bn >= 32; bn -= 32) if (N % bn == 0) return bn;
  return 0;
}

struct Plan {
  int bn[4], nch[4], nkb[4], N[4];
  int mt, U, tpb, t0[N_OPS + 1], att_tiles_img, S, stage_bytes;
  AttBranch br[2];
  size_t smem;
  int64_t ctrl_ints;
};

int make_plan(int B, int reso, int C, int hidden, int nb, const int* heads, const int* hs, const int* ws, int nblk, Plan* pl) {
  if (B <= 0 || reso <= 0 || C <= 0 || hidden <= 0 || nb < 1 || nb > 2 || nblk < 1) return CSWIN_ERR_INVALID;
  if (C % 64 || hidden % 64) return CSWIN_ERR_UNSUPPORTED;
  const int64_t M64 = (int64_t)B * reso * reso;
  if (M64 > (1 << 28)) return CSWIN_ERR_UNSUPPORTED;
  const int M = (int)M64;
  int ch = 0;
  int tiles_img = 0;
  for (int i = 0; i < nb; ++i) {
    AttBranch& b = pl->br[i];
    if (heads[i] <= 0 || hs[i] <= 0 || ws[i] <= 0 || reso % hs[i] || reso % ws[i] || hs[i] * ws[i] > 128) return CSWIN_ERR_UNSUPPORTED;
    b.heads = heads[i]; b.hs = hs[i]; b.ws = ws[i]; b.nww = reso / ws[i]; b.nwin = (reso / hs[i]) * (reso / ws[i]);
    b.N = hs[i] * ws[i]; b.slots = b.N <= 64 ? 2 : 1; b.nprob_img = b.nwin * b.heads;
    b.tiles_img = (b.nprob_img + b.slots - 1) / b.slots; b.ch0 = ch; b.pad0 = b.pad1 = 0;
    ch += 32 * heads[i];
    tiles_img += b.tiles_img;
  }
  if (ch != C) return CSWIN_ERR_UNSUPPORTED;                       // head_dim 32, branches cover the channels
  if (nb == 1) pl->br[1] = pl->br[0];
  const int cap = env_int("CSWIN_STAGE_BN_CAP", 128);
  const int Ns[4] = {3 * C, C, hidden, C}, Ks[4] = {C, C, C, hidden};
  const char* envs[4] = {"CSWIN_STAGE_BN_QKV", "CSWIN_STAGE_BN_PROJ", "CSWIN_STAGE_BN_FC1", "CSWIN_STAGE_BN_FC2"};
  const int dflt[4] = {pick_bn(3 * C, cap), pick_bn(C, 64), pick_bn(hidden, cap), pick_bn(C, 64)};
  int bnmax = 0;
  for (int g = 0; g < 4; ++g) {
    int bn = env_int(envs[g], dflt[g]);
    if (bn < 32 || bn > 128 || bn % 32 || Ns[g] % bn) return CSWIN_ERR_UNSUPPORTED;
    pl->bn[g] = bn; pl->nch[g] = Ns[g] / bn; pl->nkb[g] = Ks[g] / BK; pl->N[g] = Ns[g];
    bnmax = bn > bnmax ? bn : bnmax;
  }
  pl->mt = (M + BM - 1) / BM;
  pl->U = pl->mt > B ? pl->mt : B;
  pl->att_tiles_img = tiles_img;
  pl->t0[0] = 0;
  pl->t0[1] = pl->mt * pl->nch[0];
  pl->t0[2] = pl->t0[1] + B * tiles_img;
  pl->t0[3] = pl->t0[2] + pl->mt * pl->nch[1];
  pl->t0[4] = pl->t0[3] + pl->mt * pl->nch[2];
  pl->t0[5] = pl->t0[4] + pl->mt * pl->nch[3];
  pl->tpb = pl->t0[5];
  if ((int64_t)pl->tpb * nblk > INT_MAX / 2) return CSWIN_ERR_UNSUPPORTED;
  int stage = kABytes + bnmax * BK * 2;
  if (stage < kAttStageBytes) stage = kAttStageBytes;
  stage = (stage + 1023) & ~1023;
  const size_t fixed = 1024 + kStgBytes + 512 * 4 + (2 * kMaxStages + 8 + 2 * kQ) * 8 + kQ * 4 + 64;
  int S = (int)((113 * 1024 - fixed) / stage);
  if (S > kMaxStages) S = kMaxStages;
  const int forced = env_int("CSWIN_STAGE_RING", 0);
  if (forced >= 2 && forced < S) S = forced;
  if (S < 2) return CSWIN_ERR_UNSUPPORTED;
  pl->S = S; pl->stage_bytes = stage;
  pl->smem = fixed + (size_t)S * stage;
  pl->ctrl_ints = kCtrlFlags + (int64_t)kMaxBlocks * N_OPS * pl->U;
  return CSWIN_OK;
}

}  // namespace

int stage_plan(int B, int reso, int C, int hidden, int nb, const int* heads, const int* hs, const int* ws, cswin_stage_plan_t* out) {
  Plan pl;
  const int rc = make_plan(B, reso, C, hidden, nb, heads, hs, ws, 1, &pl);
  if (rc != CSWIN_OK) { memset(out, 0, sizeof(*out)); return rc; }
  out->parts_x = pl.nch[3]; out->parts_x1 = pl.nch[1]; out->ctrl_ints = pl.ctrl_ints; out->max_blocks = kMaxBlocks;
  return CSWIN_OK;
}

int stage_fwd_tc(const cswin_stage_args_t* a, cudaStream_t stream) {
  Plan pl;
  int rc = make_plan(a->B, a->reso, a->C, a->hidden, a->n_branches, a->heads, a->H_sp, a->W_sp, a->n_blocks, &pl);
  if (rc != CSWIN_OK) { set_error("stage_fwd: shape outside the persistent stage kernel's envelope (C, hidden multiples of 64, head_dim 32, windows <= 128 tokens)"); return rc; }
  if (a->n_blocks > kMaxBlocks) { set_error("stage_fwd: at most %d blocks per call", kMaxBlocks); return CSWIN_ERR_UNSUPPORTED; }
  if (a->ctrl_ints < pl.ctrl_ints) { set_error("stage_fwd: ctrl workspace too small (%lld < %lld ints)", (long long)a->ctrl_ints, (long long)pl.ctrl_ints); return CSWIN_ERR_INVALID; }
  if (tc::encode_tiled_fn() == nullptr) { set_error("stage_fwd: cuTensorMapEncodeTiled is not available"); return CSWIN_ERR_CUDA; }
  const void* ptrs[] = {a->x, a->qkv, a->att, a->x1, a->hid};
  for (const void* p : ptrs) if (!aligned16(p)) { set_error("stage_fwd: activations / workspaces must be 16-byte aligned"); return CSWIN_ERR_INVALID; }
  const int C = a->C, hidden = a->hidden, L = a->reso * a->reso, M = a->B * L;

  static thread_local StageParams P;            // 8 KB: kept off the stack
  memset(&P, 0, sizeof(P));
  {
    const void* src[4] = {a->x, a->att, a->x1, a->hid};
    const int K[4] = {C, C, C, hidden};
    for (int g = 0; g < 4; ++g) {
      const uint64_t dims[2] = {(uint64_t)K[g], (uint64_t)M}, str[1] = {(uint64_t)K[g] * 2};
      const uint32_t box[2] = {BK, BM};
      if (!tc::make_tensor_map_bf16(&P.a_map[g], src[g], 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return CSWIN_ERR_CUDA;
    }
    void* dst[4] = {a->qkv, a->x1, a->hid, a->x};
    for (int g = 0; g < 4; ++g) {
      const uint64_t dims[2] = {(uint64_t)pl.N[g], (uint64_t)M}, str[1] = {(uint64_t)pl.N[g] * 2};
      const uint32_t box[2] = {32, 32};
      if (!tc::make_tensor_map_bf16(&P.o_map[g], dst[g], 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B)) return CSWIN_ERR_CUDA;
    }
    for (int i = 0; i < a->n_branches; ++i)
      for (int j = 0; j < 3; ++j) {
        const __nv_bfloat16* base = (const __nv_bfloat16*)a->qkv + (size_t)j * C + pl.br[i].ch0;
        const uint64_t dims[4] = {(uint64_t)(32 * pl.br[i].heads), (uint64_t)a->reso, (uint64_t)a->reso, (uint64_t)a->B};
        const uint64_t str[3] = {(uint64_t)3 * C * 2, (uint64_t)3 * C * 2 * a->reso, (uint64_t)3 * C * 2 * L};
        const uint32_t box[4] = {32, (uint32_t)pl.br[i].ws, (uint32_t)pl.br[i].hs, 1};
        if (!tc::make_tensor_map_bf16(&P.qkv4[i][j], base, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B)) return CSWIN_ERR_CUDA;
      }
    if (a->n_branches == 1) for (int j = 0; j < 3; ++j) P.qkv4[1][j] = P.qkv4[0][j];
  }
  for (int j = 0; j < a->n_blocks; ++j) {
    const cswin_stage_block_t& b = a->blocks[j];
    BlockW& W = P.blk[j];
    const void* w[4] = {b.w_qkv, b.w_proj, b.w_fc1, b.w_fc2};
    const int K[4] = {C, C, C, hidden};
    for (int g = 0; g < 4; ++g) {
      if (!w[g] || !aligned16(w[g])) { set_error("stage_fwd: block %d: missing / unaligned weight %d", j, g); return CSWIN_ERR_INVALID; }
      const uint64_t dims[2] = {(uint64_t)K[g], (uint64_t)pl.N[g]}, str[1] = {(uint64_t)K[g] * 2};
      const uint32_t box[2] = {BK, (uint32_t)pl.bn[g]};
      if (!tc::make_tensor_map_bf16(&W.w[g], w[g], 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return CSWIN_ERR_CUDA;
    }
    W.bias[0] = b.b_qkv; W.bias[1] = b.b_proj; W.bias[2] = b.b_fc1; W.bias[3] = b.b_fc2;
    W.cs[0] = b.cs_qkv; W.cs[1] = b.cs_fc1;
    for (int i = 0; i < 2; ++i) { W.cw[i] = (const __nv_bfloat16*)b.lepe_w[i < a->n_branches ? i : 0]; W.cb[i] = (const __nv_bfloat16*)b.lepe_b[i < a->n_branches ? i : 0]; }
    W.eps1 = b.eps1; W.eps2 = b.eps2;
    if (!W.bias[0] || !W.bias[1] || !W.bias[2] || !W.bias[3] || !W.cs[0] || !W.cs[1] || !W.cw[0] || !W.cb[0] || !aligned16(W.cw[0]) || !aligned16(W.cw[1])) {
      set_error("stage_fwd: block %d: null bias / column-sum / LePE pointer (or unaligned LePE weight)", j); return CSWIN_ERR_INVALID;
    }
  }
  for (int g = 0; g < 4; ++g) {
    P.op[g].N = pl.N[g]; P.op[g].nkb = pl.nkb[g]; P.op[g].BN = pl.bn[g]; P.op[g].nch = pl.nch[g];
    P.op[g].act = g == 2 ? 1 : 0; P.op[g].fold = (g == 0 || g == 2) ? 1 : 0; P.op[g].has_res = (g == 1 || g == 3) ? 1 : 0;
  }
  P.br[0] = pl.br[0]; P.br[1] = pl.br[1];
  P.x = (const __nv_bfloat16*)a->x; P.x1 = (const __nv_bfloat16*)a->x1; P.att = (__nv_bfloat16*)a->att;
  P.stats_in = a->stats_in; P.stats_x = a->stats_x; P.stats_x1 = a->stats_x1; P.ctrl = a->ctrl;
  P.parts_in = a->stats_in_parts; P.nb = a->n_branches; P.nblk = a->n_blocks; P.M = M; P.L = L; P.B = a->B; P.reso = a->reso; P.C = C;
  P.mt = pl.mt; P.U = pl.U; P.tpb = pl.tpb; P.total = pl.tpb * a->n_blocks; P.att_tiles_img = pl.att_tiles_img;
  P.S = pl.S; P.stage_bytes = pl.stage_bytes;
  for (int i = 0; i <= N_OPS; ++i) P.t0[i] = pl.t0[i];
  P.trace = g_trace.load(std::memory_order_relaxed);
  P.tile_trace = P.trace != nullptr && env_int("CSWIN_STAGE_TILE_TRACE", 0) ? 1 : 0;
  P.scale = a->scale; P.scale_log2e = a->scale * 1.4426950408889634f; P.invC = 1.0f / (float)C;
  if (!a->stats_in || a->stats_in_parts <= 0 || !a->stats_x || !a->stats_x1 || !a->ctrl) { set_error("stage_fwd: null statistics / ctrl pointer"); return CSWIN_ERR_INVALID; }

  static std::atomic<int> configured{0};
  if (!configured.load(std::memory_order_acquire)) {
    CSWIN_CUDA_OK(cudaFuncSetAttribute(stage_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
    configured.store(1, std::memory_order_release);
  }
  int grid = env_int("CSWIN_STAGE_CTAS_PER_SM", 2) * sm_count();
  if (env_int("CSWIN_STAGE_CTAS", 0) > 0) grid = env_int("CSWIN_STAGE_CTAS", 0);
  if (grid > P.total) grid = P.total;
  CSWIN_CUDA_OK(launch_pdl(stage_tc_kernel, dim3(grid), dim3(kThreads), pl.smem, stream, P));
  CSWIN_LAUNCH_CHECK();
  g_tc_launches.fetch_add(1, std::memory_order_relaxed);
  return CSWIN_OK;
}

}  // namespace cswin
