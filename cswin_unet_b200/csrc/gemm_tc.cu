// gemm_tc.cu — bf16 Linear on tcgen05 / TMEM / TMA with the fused epilogue of cswin_linear_fwd (sm_100a).
//
//   out[m,n] = residual[m,n] + sample_scale[m / rps] * act( sum_k [a | a2][m,k] * w[n,k] + bias[n] )
//
// Replaces nn.Linear and the separate element-wise kernels of the reference around it (networks/cswin_unet.py:169,
// :177-178, Mlp :22-26 + :179, concat_linear :509-527, and the 1x1 / im2col'ed convs :216, :240-241, :264, :339, :542).
//
// One 128 x BN output tile per CTA (BN = 16..256, picked so that the grid covers the 148 SMs at least twice where the
// problem allows), K consumed in 64-wide blocks (128-byte rows, 128-byte swizzle) through a 1..4 stage TMA -> smem
// ring.  Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread tcgen05.mma issuer,
// warps 2..5 = epilogue.  The two-source form ([skip | x] for the decoder's concat Linears) switches tensor maps
// at the K1 boundary, so the concatenated activation is never materialised.  Ragged M / N / K tails are handled by
// TMA out-of-bounds zero fill on the loads and by predicated stores.
//
// Epilogue (per warp = 32 accumulator rows, 64 columns at a time): tcgen05.ld -> bias / GELU(erf) / DropPath scale in
// fp32 registers -> fp32 staging tile in shared memory (XOR-swizzled 16-byte chunks, conflict-free) -> read back
// row-contiguous so that the residual load and the output store are fully coalesced 16-byte accesses.
#include <climits>
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace cswin {
namespace {

using namespace tc;

constexpr int BM = 128, BK = 64;
constexpr int kMaxStages = 8;
constexpr int kThreads = 320;                          // warp 0 TMA, warp 1 MMA + TMEM, warps 2..9 epilogue

struct alignas(64) GemmTcParams {
  CUtensorMap map_a, map_a2, map_w;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* res; int64_t ldr;
  const float* sscale; int rps;
  __nv_bfloat16* out; int64_t ldo;
  int64_t M; int N; int K1; int K2;
  int BN, stages, nkb, nkb1, act, tmem_cols, vec_ok, bias_vec, w_kn;
  unsigned long long* trace;
};

// GELU(erf) for the bf16 epilogue: x * Phi(x) with Phi(x) = 0.5 (1 + tanh(x (c0 + c1 x^2 + c2 x^4 + c3 x^6))), the odd
// degree-7 minimax fit of atanh(erf(x / sqrt 2)) (max |dPhi| = 6.6e-6, max |dGELU| = 2.4e-5 over all x) evaluated with one
// MUFU op (tanh.approx, relative error 2^-11): total error <= 2.5e-4 |x|, i.e. >= 16x below bf16 resolution, for
// 7 FMA-pipe instructions + 1 MUFU instead of erff's ~40.  The fp32 SIMT path keeps erff.
__device__ __forceinline__ float gelu_fast(float x) {
  const float x2 = x * x;
  float p = fmaf(x2, -1.36882761e-05f, -1.94451094e-04f);
  p = fmaf(p, x2, 3.65466544e-02f);
  p = fmaf(p, x2, 7.97820264e-01f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(p * x));
  const float h = 0.5f * x;
  return fmaf(h, t, h);
}

__global__ void __launch_bounds__(kThreads, 2) linear_tc_kernel(const __grid_constant__ GemmTcParams P) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128-byte swizzle; plain pointer arithmetic keeps the shared address space (LDS/STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int BN = P.BN, S = P.stages;
  const uint32_t a_bytes = BM * BK * 2, w_bytes = (uint32_t)BN * BK * 2;
  uint8_t* As = smem;                                   // [S][128][64] bf16, SW128
  uint8_t* Ws = As + (size_t)S * a_bytes;               // [S][BN][64]
  // epilogue staging (8 warps x 32 rows x 64 B, bf16, swizzled) re-uses the operand ring: it is only touched after
  // bar_acc, i.e. after every TMA load has landed and every MMA has finished reading shared memory
  uint8_t* Epi = smem;
  const size_t ring = (size_t)S * (a_bytes + w_bytes);
  float* sBias = reinterpret_cast<float*>(smem + ring);                  // [256] fp32
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 256);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  if (tid == 0) trace_stamp(P.trace, 0);                                   // kernel entry
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  auto full = [&](int s) { return smem_u32(&bars[s]); };
  auto empty = [&](int s) { return smem_u32(&bars[kMaxStages + s]); };
  const uint32_t bar_acc = smem_u32(&bars[2 * kMaxStages]);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
    tma_prefetch_desc(&P.map_a); tma_prefetch_desc(&P.map_w);
    if (P.K2 > 0) tma_prefetch_desc(&P.map_a2);
  }
  if (warp == 1) { tmem_alloc(smem_u32(tmem_slot), (uint32_t)P.tmem_cols); tmem_relinquish(); }
  if (warp >= 2) {                                      // bias of this tile's columns -> smem (fp32)
    const int j = tid - 64, n = n0 + j;
    sBias[j] = (P.bias != nullptr && j < BN && n < P.N) ? __bfloat162float(P.bias[n]) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) trace_stamp(P.trace, 1);                                   // prologue done

  // weights do not depend on the previous kernel: the W tiles of the first ring pass are requested before the PDL wait
  const int npre = P.nkb < S ? P.nkb : S;
  if (warp == 0 && lane == 0) {
    for (int kb = 0; kb < npre; ++kb) {
      mbar_expect_tx(full(kb), a_bytes + w_bytes);
      if (!P.w_kn) tma_load_2d(smem_u32(Ws + (size_t)kb * w_bytes), &P.map_w, full(kb), kb * BK, n0);
      else for (int j = 0; j * 64 < BN; ++j)             // (K, N) weight: [64 k][64 n] boxes = MN-major B blocks
        tma_load_2d(smem_u32(Ws + (size_t)kb * w_bytes + j * 8192), &P.map_w, full(kb), n0 + 64 * j, kb * BK);
    }
  }
  pdl_wait();                                           // activations (A, residual) and the output buffer are safe from here

  if (warp == 0) {
    if (lane == 0) {                                    // ---- TMA producer ----
      for (int kb = 0; kb < P.nkb; ++kb) {
        const int s = kb % S;
        if (kb >= S) {
          mbar_wait(empty(s), ((kb / S) - 1) & 1);
          mbar_expect_tx(full(s), a_bytes + w_bytes);
          if (!P.w_kn) tma_load_2d(smem_u32(Ws + (size_t)s * w_bytes), &P.map_w, full(s), kb * BK, n0);
          else for (int j = 0; j * 64 < BN; ++j)
            tma_load_2d(smem_u32(Ws + (size_t)s * w_bytes + j * 8192), &P.map_w, full(s), n0 + 64 * j, kb * BK);
        }
        if (kb < P.nkb1) tma_load_2d(smem_u32(As + (size_t)s * a_bytes), &P.map_a, full(s), kb * BK, (int)m0);
        else             tma_load_2d(smem_u32(As + (size_t)s * a_bytes), &P.map_a2, full(s), (kb - P.nkb1) * BK, (int)m0);
      }
      trace_stamp(P.trace, 2);                          // all TMA issued
    }
  } else if (warp == 1) {
    if (lane == 0) {                                    // ---- MMA issuer ----
      const uint32_t idesc = make_idesc_bf16(BM, BN, 0, P.w_kn);
      const uint32_t wstep = P.w_kn ? (2048 >> 4) : 2;   // 16 contraction rows: 2 KB (MN-major) or 32 B (K-major) further
      for (int kb = 0; kb < P.nkb; ++kb) {
        const int s = kb % S;
        mbar_wait(full(s), (kb / S) & 1);
        if (kb == 0) trace_stamp(P.trace, 3);           // first operands landed
        tc_fence_after();
        const uint64_t ad = make_smem_desc(smem_u32(As + (size_t)s * a_bytes), 16, 1024, kLayoutSw128);
        const uint64_t wd = make_smem_desc(smem_u32(Ws + (size_t)s * w_bytes), P.w_kn ? 8192 : 16, 1024, kLayoutSw128);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) mma_ss(tmem_base, ad + 2 * k, wd + (uint64_t)wstep * k, idesc, (kb | k) != 0);
        tc_commit(empty(s));                            // smem slot reusable once these MMAs have read it
      }
      tc_commit(bar_acc);                               // accumulator complete
      trace_stamp(P.trace, 4);                          // all MMAs issued
    }
  } else {
    // ---- epilogue: warps 2..9.  TMEM lane quadrant q = warp % 4 (hardware rule); the two warps of a quadrant take the
    //      even / odd 32-column units.  Per unit: tcgen05.ld -> bias / GELU / DropPath scale in registers -> bf16 ->
    //      swizzled 2 KB staging tile -> row-contiguous read-back so the residual load and the store are coalesced 16 B.
    const int q = warp & 3;
    const int ch = (warp - 2) >> 2;
    uint8_t* stg = Epi + (warp - 2) * 2048;
    const uint32_t stg_u32 = smem_u32(stg);
    const int64_t mrow = m0 + q * 32 + lane;            // accumulator row held by this thread in phase 1
    const float sc = (P.sscale != nullptr && mrow < P.M) ? P.sscale[mrow / P.rps] : 1.0f;
    const int nunits = (BN + 31) >> 5;
    mbar_wait(bar_acc, 0);
    if (warp == 2 && lane == 0) trace_stamp(P.trace, 5);  // accumulator ready
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int u = ch; u < nunits; u += 2) {
      uint32_t v[32];
      tmem_ld32(trow + u * 32, v);
      tmem_wait_ld();
      const float4* b4 = reinterpret_cast<const float4*>(sBias + u * 32);
#pragma unroll
      for (int c = 0; c < 4; ++c) {                     // four 16-byte chunks of 8 columns
        float f[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 bb = b4[c * 2 + h];
          f[h * 4 + 0] = __uint_as_float(v[c * 8 + h * 4 + 0]) + bb.x;
          f[h * 4 + 1] = __uint_as_float(v[c * 8 + h * 4 + 1]) + bb.y;
          f[h * 4 + 2] = __uint_as_float(v[c * 8 + h * 4 + 2]) + bb.z;
          f[h * 4 + 3] = __uint_as_float(v[c * 8 + h * 4 + 3]) + bb.w;
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
#pragma unroll
      for (int pass = 0; pass < 4; ++pass) {            // 8 rows x 64 B per pass, 4 lanes per row
        const int r = pass * 8 + (lane >> 2), c = lane & 3;
        const int64_t m = m0 + q * 32 + r;
        const int n = n0 + u * 32 + c * 8;
        uint4 w;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w)
                     : "r"(stg_u32 + r * 64 + (((c ^ (r >> 1)) & 3) << 4)));
        if (m < P.M && n < P.N && u * 32 + c * 8 < BN) {
          __nv_bfloat16* dst = P.out + m * P.ldo + n;
          if (P.vec_ok && n + 8 <= P.N) {
            if (P.res != nullptr) {
              const uint4 rv = *reinterpret_cast<const uint4*>(P.res + m * P.ldr + n);
              w.x = pack_bf16x2(bf16_lo(w.x) + bf16_lo(rv.x), bf16_hi(w.x) + bf16_hi(rv.x));
              w.y = pack_bf16x2(bf16_lo(w.y) + bf16_lo(rv.y), bf16_hi(w.y) + bf16_hi(rv.y));
              w.z = pack_bf16x2(bf16_lo(w.z) + bf16_lo(rv.z), bf16_hi(w.z) + bf16_hi(rv.z));
              w.w = pack_bf16x2(bf16_lo(w.w) + bf16_lo(rv.w), bf16_hi(w.w) + bf16_hi(rv.w));
            }
            *reinterpret_cast<uint4*>(dst) = w;
          } else {
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              if (n + e < P.N) {
                float x = (e & 1) ? bf16_hi(ww[e >> 1]) : bf16_lo(ww[e >> 1]);
                if (P.res != nullptr) x += __bfloat162float(P.res[m * P.ldr + n + e]);
                dst[e] = __float2bfloat16_rn(x);
              }
            }
          }
        }
      }
      __syncwarp();
    }
    if (warp == 2 && lane == 0) trace_stamp(P.trace, 6);  // epilogue done
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) trace_stamp(P.trace, 7);                                   // exit
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

size_t smem_bytes(int bn, int stages) {
  const size_t ring = (size_t)stages * (BM * BK * 2 + (size_t)bn * BK * 2);     // >= 18 KB > the aliased 16 KB staging
  return 1024 + ring + 1024 /* bias */ + 256 /* barriers, TMEM slot */;
}
int tmem_cols_for(int bn) { return bn <= 64 ? 64 : bn <= 128 ? 128 : 256; }      // the epilogue reads whole 64-col groups

struct TileCfg { int bn, stages; };

// Tile-shape choice.  At cswin_tiny sizes a Linear is a handful of waves at most, so the launch is latency- not
// throughput-bound: the model below (constants fitted to L2-warm CUDA-event timings on B200, microseconds) trades the
// number of waves against per-tile latency = fixed setup + K-loop (faster with a deeper ring) + epilogue (per 64
// columns; GELU costs ~2x).  The ring depth is whatever fits once the CTAs that must share an SM are accounted for.
// CSWIN_GEMM_BN=<n> forces BN for experiments.
TileCfg pick_tile(int64_t M, int N, int nkb, int act, int sms, bool w_kn) {
  static const int forced = [] { const char* e = getenv("CSWIN_GEMM_BN"); return e ? atoi(e) : 0; }();
  const int n16 = w_kn ? ((N + 63) & ~63) : ((N + 15) & ~15);      // (K,N) weights are fetched in 64-column boxes
  const int64_t mt = (M + BM - 1) / BM;
  const int cands[] = {64, 96, 128, 192, 256};
  TileCfg best{n16 < 64 ? n16 : 64, 1};
  double best_t = 1e30;
  for (int bn : cands) {
    if (forced >= 16 && forced <= 256 && forced % 16 == 0) bn = forced;
    if (w_kn && bn % 64) continue;
    if (bn > n16) bn = n16 > 256 ? 256 : n16;
    if (N % bn != 0 && bn > 64 && bn != n16) continue;
    const int64_t tiles = mt * ((N + bn - 1) / bn);
    int resident = 512 / tmem_cols_for(bn);
    if (resident > 2) resident = 2;                              // 320 threads x ~80 registers: two CTAs per SM
    while (resident > 1 && smem_bytes(bn, 2) * resident > 220 * 1024) --resident;
    const int64_t waves = (tiles + (int64_t)sms * resident - 1) / ((int64_t)sms * resident);
    int64_t share = (tiles + sms - 1) / sms;                    // CTAs that will actually share an SM
    if (share > resident) share = resident;
    int st = 1;
    while (st < kMaxStages && st < nkb && smem_bytes(bn, st + 1) * share <= 220 * 1024) ++st;
    const double t_tile = 2.0 + nkb * (0.10 + 0.7 / st) + (bn / 64.0) * (act ? 3.5 : 1.6);
    const double t = waves * t_tile;
    if (t < best_t - 1e-9) { best_t = t; best = TileCfg{bn, st}; }
    if (forced) break;
  }
  return best;
}

}  // namespace

int linear_fwd_tc(const cswin_linear_args_t* a, cudaStream_t stream, bool* handled) {
  *handled = false;
  if (a->ln_gamma != nullptr) return CSWIN_OK;                               // LayerNorm prologue: SIMT kernel (host calls LN first on the bf16 path)
  const int K = a->K1 + a->K2;
  if (!aligned16(a->a) || !aligned16(a->w) || (a->lda * 2) % 16 || (a->ldw * 2) % 16) return CSWIN_OK;
  if (a->K1 % 8 || K % 8) return CSWIN_OK;
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
  const TileCfg cfg = pick_tile(a->M, a->N, P.nkb, a->act, sm_count(), a->w_layout != 0);
  P.BN = cfg.bn;
  P.stages = cfg.stages;
  P.tmem_cols = tmem_cols_for(P.BN);
  P.bias_vec = a->bias != nullptr && aligned16(a->bias);
  P.trace = g_trace.load(std::memory_order_relaxed);
  P.vec_ok = aligned16(a->out) && (a->ldo * 2) % 16 == 0 &&
             (a->residual == nullptr || (aligned16(a->residual) && (a->ldr * 2) % 16 == 0));

  {
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

  const size_t smem = smem_bytes(P.BN, P.stages);
  static std::atomic<size_t> configured{0};
  if (smem > configured.load(std::memory_order_relaxed)) {
    CSWIN_CUDA_OK(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured.store(227 * 1024, std::memory_order_relaxed);
  }
  dim3 grid((unsigned)((a->M + BM - 1) / BM), (unsigned)((a->N + P.BN - 1) / P.BN));
  CSWIN_CUDA_OK(launch_pdl(linear_tc_kernel, grid, dim3(kThreads), smem, stream, P));
  CSWIN_LAUNCH_CHECK();
  g_tc_launches.fetch_add(1, std::memory_order_relaxed);
  *handled = true;
  return CSWIN_OK;
}

}  // namespace cswin
