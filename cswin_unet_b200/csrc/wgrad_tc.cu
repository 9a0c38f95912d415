// wgrad_tc.cu — weight gradient of a Linear on tcgen05 / TMEM / TMA (bf16 operands, fp32 accumulate, fp32 atomics).
//
//     dW[n, k] += sum_m dZ[m, n] * A[m, k]          (autograd of nn.Linear: cswin_unet.py:169, :177, Mlp :22-26, ...)
//
// Both operands are stored row-major with the contraction index m as the SLOW dimension, so both enter the MMA
// "MN-major": a TMA box (64 columns, 64 rows of m) lands as 64 rows of 128 B (128-byte swizzle), which is exactly the
// canonical MN-major layout (8-row groups 1024 B apart = SBO; the next 64-column block 8 KB further = LBO).  No
// transposed copy of dZ or A is ever made.
//   tile: UMMA M = 128 output rows (n), UMMA N = BN output columns (k, 64..256), contraction in 64-row blocks of m;
//   the M range is split over gridDim.z CTAs (the output is tiny, the contraction long), each CTA accumulates its
//   slice in TMEM and adds it to the fp32 gradient with red.global.add.f32.
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace cswin {
namespace {

using namespace tc;

constexpr int TN = 128;           // n rows per tile (UMMA M)
constexpr int MB = 64;            // m rows per pipeline block
#ifndef CSWIN_WGRAD_STAGES
#define CSWIN_WGRAD_STAGES 4             // TMA ring depth (2 leaves room for a second kernel's CTA on the SM: see autograd._Fork)
#endif
constexpr int kWStages = CSWIN_WGRAD_STAGES;
constexpr int kWThreads = 192;    // warp 0 TMA, warp 1 MMA + TMEM, warps 2..5 epilogue

struct alignas(64) WgradParams {
  CUtensorMap map_z, map_a;
  float* dw; int64_t ldw;
  float* db;                      // bias gradient (column sums of dZ), accumulated by the k-tile-0 CTAs; may be NULL
  int N, K, BN, nblk, tmem_cols, vec4;
  int64_t M, mchunk;
};

__global__ void __launch_bounds__(kWThreads) linear_wgrad_tc_kernel(const __grid_constant__ WgradParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int BN = P.BN;
  const uint32_t z_bytes = MB * TN * 2;                 // two boxes of [64 m][64 n]
  const uint32_t a_bytes = (uint32_t)MB * BN * 2;       // BN/64 boxes of [64 m][64 k]
  uint8_t* Zs = smem;                                   // [stages][2][64][128 B]
  uint8_t* As = Zs + (size_t)kWStages * z_bytes;        // [stages][BN/64][64][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(As + (size_t)kWStages * a_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWStages + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  const int n0 = blockIdx.x * TN, k0 = blockIdx.y * BN;
  const int64_t mbeg = (int64_t)blockIdx.z * P.mchunk;
  const int64_t mend = mbeg + P.mchunk < P.M ? mbeg + P.mchunk : P.M;
  const int nmb = (int)((mend - mbeg + MB - 1) / MB);

  // bias gradient: the epilogue warps of the first k-tile's CTAs are idle during the main loop, so they sum the columns of
  // every dZ block as it lands in shared memory (each block is released by the MMA commit AND these four warps)
  const bool do_colsum = P.db != nullptr && blockIdx.y == 0;
  auto full = [&](int s) { return smem_u32(&bars[s]); };
  auto empty = [&](int s) { return smem_u32(&bars[kWStages + s]); };
  const uint32_t bar_acc = smem_u32(&bars[2 * kWStages]);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kWStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), do_colsum ? 5u : 1u); }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
    tma_prefetch_desc(&P.map_z); tma_prefetch_desc(&P.map_a);
  }
  if (warp == 1) { tmem_alloc(smem_u32(tmem_slot), (uint32_t)P.tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                           // dZ, A and the gradient buffer are safe from here

  if (warp == 0) {
    if (elect_one()) {
      int s = 0; uint32_t ph = 1;
      for (int mb = 0; mb < nmb; ++mb) {
        if (mb >= kWStages) mbar_wait(empty(s), ph);
        mbar_expect_tx(full(s), z_bytes + a_bytes);
        const int m = (int)(mbeg + (int64_t)mb * MB);
        for (int j = 0; j < 2; ++j)
          tma_load_2d(smem_u32(Zs + (size_t)s * z_bytes + j * 8192), &P.map_z, full(s), n0 + 64 * j, m);
        for (int j = 0; j < P.nblk; ++j)
          tma_load_2d(smem_u32(As + (size_t)s * a_bytes + j * 8192), &P.map_a, full(s), k0 + 64 * j, m);
        if (++s == kWStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(TN, BN, 1, 1);             // both operands MN-major
      int s = 0; uint32_t ph = 0;
      for (int mb = 0; mb < nmb; ++mb) {
        mbar_wait(full(s), ph);
        tc_fence_after();
        // MN-major, 128-byte swizzle: LBO = 8 KB (next 64-wide block along n / k), SBO = 1 KB (next 8 rows of m)
        const uint64_t zd = make_smem_desc(smem_u32(Zs + (size_t)s * z_bytes), 8192, 1024, kLayoutSw128);
        const uint64_t ad = make_smem_desc(smem_u32(As + (size_t)s * a_bytes), 8192, 1024, kLayoutSw128);
#pragma unroll
        for (int k = 0; k < MB / 16; ++k)                                // 16 rows of m per MMA = 2 KB further
          mma_ss(tmem_base, zd + (uint64_t)k * (2048 >> 4), ad + (uint64_t)k * (2048 >> 4), idesc, (mb | k) != 0);
        tc_commit(empty(s));
        if (++s == kWStages) { s = 0; ph ^= 1; }
      }
      tc_commit(bar_acc);
    }
  } else {
    const int q = warp & 3;
    if (do_colsum) {
      // thread <-> column c = (warp - 2) * 32 + lane of the 128-column dZ tile: box j = c / 64, 16-byte chunk (cc / 8) ^ (m % 8)
      const int c = (warp - 2) * 32 + lane, cc = c & 63;
      const uint32_t cbase = (uint32_t)(c >> 6) * 8192 + (uint32_t)(cc & 7) * 2;
      float acc = 0.f;
      int s = 0; uint32_t ph = 0;
      for (int mb = 0; mb < nmb; ++mb) {
        mbar_wait(full(s), ph);
        const uint32_t zb = smem_u32(Zs + (size_t)s * z_bytes) + cbase;
#pragma unroll 16
        for (int m = 0; m < MB; ++m) {
          uint16_t h;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(zb + m * 128 + ((((cc >> 3) ^ (m & 7))) << 4)));
          acc += __uint_as_float((uint32_t)h << 16);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty(s));
        if (++s == kWStages) { s = 0; ph ^= 1; }
      }
      if (n0 + c < P.N && nmb > 0) atomicAdd(P.db + n0 + c, acc);
    }
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const int n = n0 + q * 32 + lane;
    for (int c = 0; c * 32 < BN; ++c) {
      uint32_t v[32];
      tmem_ld32(trow + c * 32, v);
      tmem_wait_ld();
      if (n < P.N && nmb > 0) {
        float* dst = P.dw + (int64_t)n * P.ldw + k0 + c * 32;
        if (P.vec4 && k0 + c * 32 + 32 <= P.K) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)                // 16-byte vector reductions: 4x fewer L2 atomic operations
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])),
                         "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3])) : "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (k0 + c * 32 + j < P.K) atomicAdd(dst + j, __uint_as_float(v[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int linear_wgrad_tc(const void* dz, int64_t ldz, const void* a, int64_t lda, float* dw, int64_t ldw, float* db, int64_t M,
                    int N, int K, cudaStream_t stream, bool* handled) {
  *handled = false;
  if (!aligned16(dz) || !aligned16(a) || (ldz * 2) % 16 || (lda * 2) % 16 || M > 0x7fffffff) return CSWIN_OK;
  if (tc::encode_tiled_fn() == nullptr) return CSWIN_OK;
  WgradParams P;
  P.dw = dw; P.ldw = ldw; P.db = db; P.N = N; P.K = K; P.M = M;
  P.vec4 = aligned16(dw) && (ldw * 4) % 16 == 0;
  P.nblk = K >= 256 ? 4 : (K + 63) / 64;
  P.BN = P.nblk * 64;
  P.tmem_cols = P.BN <= 64 ? 64 : P.BN <= 128 ? 128 : 256;
  const int tiles = ((N + TN - 1) / TN) * ((K + P.BN - 1) / P.BN);
  // split the contraction so that ~1 CTA per SM exists (measured: profiles/r01_wgrad_split_sweep.log; 2 per SM costs 5 % more over
  // the 16 block shapes because of the extra fp32 atomics), but keep the number of fp32 atomics per launch bounded
  // CSWIN_WGRAD_CTAS_PER_SM / CSWIN_WGRAD_ATOMIC_CAP (elements of dW x split allowed per launch): tuning knobs
  static const int ctas_per_sm = [] { const char* e = getenv("CSWIN_WGRAD_CTAS_PER_SM"); return e ? std::max(1, atoi(e)) : 1; }();
  static const int64_t atomic_cap = [] { const char* e = getenv("CSWIN_WGRAD_ATOMIC_CAP"); return e ? std::max<int64_t>(1, atoll(e)) : (int64_t)(4 << 20); }();
  int64_t split = std::max<int64_t>(1, ((int64_t)ctas_per_sm * sm_count()) / tiles);
  const int64_t max_by_atomics = std::max<int64_t>(1, atomic_cap / ((int64_t)N * K));
  split = std::min(split, std::max<int64_t>(max_by_atomics, 1));
  split = std::min(split, ceil_div64(M, 2 * MB));
  if (split < 1) split = 1;
  P.mchunk = ceil_div64(ceil_div64(M, split), MB) * MB;
  split = ceil_div64(M, P.mchunk);
  {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    const uint64_t str[1] = {(uint64_t)ldz * 2};
    const uint32_t box[2] = {64, MB};
    if (!tc::make_tensor_map_bf16(&P.map_z, dz, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return CSWIN_ERR_CUDA;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t str[1] = {(uint64_t)lda * 2};
    const uint32_t box[2] = {64, MB};
    if (!tc::make_tensor_map_bf16(&P.map_a, a, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return CSWIN_ERR_CUDA;
  }
  const size_t smem = 1024 + (size_t)kWStages * (MB * TN * 2 + (size_t)MB * P.BN * 2) + 256;
  static std::atomic<bool> configured{false};
  if (!configured.load(std::memory_order_relaxed)) {
    CSWIN_CUDA_OK(cudaFuncSetAttribute(linear_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured.store(true, std::memory_order_relaxed);
  }
  dim3 grid((unsigned)((N + TN - 1) / TN), (unsigned)((K + P.BN - 1) / P.BN), (unsigned)split);
  CSWIN_CUDA_OK(launch_pdl(linear_wgrad_tc_kernel, grid, dim3(kWThreads), smem, stream, P));
  CSWIN_LAUNCH_CHECK();
  g_tc_launches.fetch_add(1, std::memory_order_relaxed);
  *handled = true;
  return CSWIN_OK;
}

}  // namespace cswin
