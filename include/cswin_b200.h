/*
 * cswin_b200.h — C ABI of libcswin_b200.so: the B200 (sm_100a) kernels behind the CSWin-UNet hot path.
 *
 * The reference (BoloniniD/CSWin-UNet) is pure PyTorch: its "FFI" for this path is the set of ATen calls
 * made by networks/cswin_unet.py.  Every entry point below replaces one fused group of those calls and cites
 * the reference lines it stands in for.  The drop-in nn.Modules in cswin_unet_b200/ (same constructor
 * arguments, attribute names and state_dict keys as the reference classes) bind these symbols with ctypes;
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *  - Plain pointers and sizes only.  Every pointer is a DEVICE pointer owned by the caller (torch allocates all
 *    inputs, outputs and workspaces); the library allocates no device memory and keeps no state except a
 *    mutex-guarded per-device attribute cache.
 *  - All work is enqueued on `stream` of the CURRENT device; nothing synchronises, allocates or calls
 *    cudaSetDevice, so every call is CUDA-graph capturable and re-entrant (autograd / DataParallel threads).
 *  - `dtype`: CSWIN_F32 (fp32 storage, fp32 SIMT arithmetic — the <=1e-4 parity path) or CSWIN_BF16 (bf16
 *    storage, tcgen05 tensor-core contractions with fp32 accumulation, fp32 softmax / LayerNorm / GELU).
 *    LayerNorm / conv-bias / LePE parameters are passed in the same dtype as the activations.
 *  - Strides and leading dimensions are in ELEMENTS.  Channel (innermost) stride is always 1.
 *  - Return value: 0 on success, otherwise a CSWIN_ERR_* code; cswin_last_error() returns a thread-local
 *    message.  Nothing exits, throws or asserts across the ABI.
 */
#ifndef CSWIN_B200_H_
#define CSWIN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSWIN_ABI_VERSION 8

typedef struct CUstream_st* cswin_stream_t; /* == cudaStream_t */

enum { CSWIN_F32 = 0, CSWIN_BF16 = 1 };

enum {
  CSWIN_OK = 0,
  CSWIN_ERR_INVALID = 1,     /* bad shape / stride / alignment / null pointer */
  CSWIN_ERR_UNSUPPORTED = 2, /* valid request outside the implemented envelope */
  CSWIN_ERR_CUDA = 3         /* a CUDA runtime / driver call failed */
};

int cswin_abi_version(void);
const char* cswin_last_error(void);
/* number of kernels this library has launched in the calling process (for bench.py's gpu_launches) */
uint64_t cswin_launch_count(void);
/* ... of which tcgen05 / TMEM / TMA kernels (attention_tc.cu, gemm_tc.cu); lets tests assert the tensor-core path ran */
uint64_t cswin_tc_launch_count(void);
/* bf16 calls that fell outside a tcgen05 kernel's envelope and ran on the general SIMT kernels instead (same results, much slower):
 * counted here; the first occurrence per op also prints one warning line to stderr (CSWIN_QUIET_FALLBACK=1 silences it). */
uint64_t cswin_simt_fallback_count(void);
/* process-wide tuning options (take effect for launches enqueued / graphs captured AFTER the call).
 *  CSWIN_OPT_GEMM_SMEM_CAP_KB: per-CTA shared-memory ceiling of the tcgen05 Linear's operand ring in KB (0 = none, default).
 *    With several independent forwards in flight (SliceEngine(inflight > 1)) a ceiling of ~100 KB lets CTAs of two launches share
 *    an SM: +6 % throughput at 3 forwards in flight, -4 % for a single forward (profiles/r02_concurrent_forwards.log). */
enum { CSWIN_OPT_GEMM_SMEM_CAP_KB = 1 };
int cswin_set_option(int32_t option, int32_t value);
/* debug / profiling aid: when non-NULL, the tcgen05 kernels write %globaltimer stamps of their phases for the first 1024
 * CTAs of every launch into this device buffer (1024 x 16 uint64); NULL (default) disables it. Not part of the data path. */
void cswin_debug_set_trace(void* device_buffer);

/* ------------------------------------------------------------------------------------------------
 * LePE cross-shaped-window attention.
 * Replaces LePEAttention.forward, networks/cswin_unet.py:82-109, including im2cswin :59-65, get_lepe :67-80,
 * get_v depthwise conv :55, img2windows :184-191, windows2img :194-202 and the branch concat of
 * CSWinBlock.forward :172-176 (each branch writes its channel slice of the (B,L,C) output directly).
 *
 * For every batch b, window (ih,iw), head g:   out = softmax(scale * q k^T) v + dwconv3x3_windowpad(v) + bias
 * q/k/v point at element (b=0, token=0, first channel of the branch); token t = y*reso + x.
 *
 * Kernel envelope (bf16): head_dim 32, 16-byte aligned rows and windows of N = H_sp*W_sp <= 128 tokens run on the tcgen05
 * kernels; 128 < N <= 256 on their wide variants (forward additionally needs 128 % W_sp == 0); anything else — and every fp32
 * call — on the general SIMT kernels (same results, slower).  Nothing is refused for its shape.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  const void* q; const void* k; const void* v;
  int64_t q_bs, q_ts;   /* batch stride, token stride of q (elements) */
  int64_t k_bs, k_ts;
  int64_t v_bs, v_ts;
  void* out;            /* element (b=0, token=0, first channel of the branch) of the (B,L,C) output */
  int64_t o_bs, o_ts;
  const void* conv_w;   /* get_v.weight, (C_b,1,3,3) contiguous */
  const void* conv_b;   /* get_v.bias, (C_b) */
  float* lse;           /* optional (B, L, heads) fp32 log-sum-exp of the scaled scores, or NULL */
  int32_t C_b;          /* channels of this branch */
  int32_t heads;        /* heads of this branch; head_dim = C_b / heads */
  int32_t H_sp, W_sp;   /* stripe window shape (cswin_unet.py:43-53) */
} cswin_lepe_branch_t;

int cswin_lepe_attention_fwd(const cswin_lepe_branch_t* branches, int32_t n_branches, int32_t B, int32_t reso,
                             float scale, int32_t dtype, cswin_stream_t stream);

/* Backward of the above (autograd of cswin_unet.py:82-109; formulas in SURVEY.md Appendix A).
 * dout has the layout of `out`; dq/dk/dv the layouts of q/k/v (may alias a (B,L,3C) buffer);
 * dconv_w (C_b,9) / dconv_b (C_b) are fp32 and ACCUMULATED into (caller zeroes them). */
typedef struct {
  cswin_lepe_branch_t fwd;      /* same description as the forward call (out/lse = forward results) */
  const void* dout; int64_t do_bs, do_ts;
  void* dq; void* dk; void* dv;
  int64_t dq_bs, dq_ts, dk_bs, dk_ts, dv_bs, dv_ts;
  float* dconv_w; float* dconv_b;
} cswin_lepe_branch_grad_t;

int cswin_lepe_attention_bwd(const cswin_lepe_branch_grad_t* branches, int32_t n_branches, int32_t B, int32_t reso,
                             float scale, int32_t dtype, cswin_stream_t stream);

/* The parameter-gradient half of the above on its own: dconv_w / dconv_b += autograd of get_v (cswin_unet.py:55, :76)
 * from `dout` and `fwd.v` only (q, k, lse, dq, dk, dv are not touched), so a caller can run it on a second stream next
 * to cswin_lepe_attention_bwd called with dconv_w = dconv_b = NULL (which then skips them).  *handled = 1 if the
 * kernel was launched, 0 if the configuration is outside its envelope (bf16, C_b % 32 == 0, 8-byte aligned rows) —
 * nothing was launched and the caller passes dconv_w / dconv_b to cswin_lepe_attention_bwd instead. */
int cswin_lepe_param_grad(const cswin_lepe_branch_grad_t* branches, int32_t n_branches, int32_t B, int32_t reso,
                          int32_t dtype, cswin_stream_t stream, int32_t* handled);

/* ------------------------------------------------------------------------------------------------
 * Slice resampling of the evaluation loop (test_single_volume, utils.py:61-90) — scipy.ndimage.zoom with scipy's defaults
 * (mode='constant', cval=0, prefilter=True, grid_mode=False), reproduced in float64 including its edge behaviour (a
 * coordinate that rounds above in-1, e.g. 223 * (511/223), yields cval: the last row / column of a 512 -> 224 zoom is 0).
 *   cswin_zoom_cubic_fwd  : utils.py:69  zoom(slice, (P/x, P/y), order=3).  in (n, H, W) float32 contiguous; work = n*H*W
 *                           doubles of caller-owned scratch (the spline coefficients); out element (s, c, oy, ox) at
 *                           s*out_slice_stride + c*out_channel_stride + oy*OW + ox for c < channel_copies (the 1 -> 3
 *                           channel repeat of vision_transformer.py:40-41 costs nothing here).
 *   cswin_zoom_nearest_u8 : utils.py:77  zoom(out, (x/P, y/P), order=0) on the uint8 label map, (n, H, W) -> (n, OH, OW).
 * ------------------------------------------------------------------------------------------------ */
int cswin_zoom_cubic_fwd(const float* in, int32_t n, int32_t H, int32_t W, double* work, float* out, int64_t out_slice_stride,
                         int64_t out_channel_stride, int32_t channel_copies, int32_t OH, int32_t OW, cswin_stream_t stream);
int cswin_zoom_nearest_u8(const uint8_t* in, int32_t n, int32_t H, int32_t W, uint8_t* out, int32_t OH, int32_t OW,
                          cswin_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm over the last dimension.  Replaces nn.LayerNorm calls: norm1/norm2 (cswin_unet.py:168,179),
 * Merge_Block.norm :218, stem LN :341, norm :497, norm_up :533.   y = (x-mean)/sqrt(var+eps)*gamma+beta
 * mean_out / rstd_out: optional fp32 (M) saved statistics for the backward.
 * ------------------------------------------------------------------------------------------------ */
int cswin_layernorm_fwd(const void* x, int64_t ldx, const void* gamma, const void* beta, void* y, int64_t ldy,
                        int64_t M, int32_t C, float eps, float* mean_out, float* rstd_out, int32_t dtype,
                        cswin_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Linear with fused prologue/epilogue.  Replaces nn.Linear + the ops fused around it:
 *   qkv :169 (LayerNorm prologue :168), proj + residual + DropPath :177-178, fc1 + GELU (Mlp :22-23, norm2 :179),
 *   fc2 + residual + DropPath :179, concat_linear{4,3,2} on cat([skip,x]) :509-510/:518-519/:526-527 (two A
 *   sources, no cat), and every 1x1 / im2col'ed conv (Merge_Block.conv :216, CARAFE down/encoder/out :240-241,:264,
 *   stem conv :339, output conv :542).
 *
 *   acc[m,n] = sum_k A[m,k] W[n,k]        A = [a | a2] along K (a2 optional), optionally LayerNorm'ed per row first
 *   t        = act(acc + bias[n])          act: 0 none, 1 GELU(erf)
 *   out[m,n] = residual[m,n] + sample_scale[m / rows_per_sample] * t        (residual / sample_scale optional)
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  const void* a;  int64_t lda;  int32_t K1;
  const void* a2; int64_t lda2; int32_t K2;       /* a2 == NULL -> K2 must be 0 */
  const void* w;  int64_t ldw;                     /* (N, K1+K2) row-major */
  const void* bias;                                /* (N) or NULL */
  const void* ln_gamma; const void* ln_beta; float ln_eps;   /* LayerNorm prologue over K1 (a2 must be NULL) or NULL */
  const void* residual; int64_t ldr;               /* (M,N) or NULL */
  const float* sample_scale; int32_t rows_per_sample;  /* DropPath m_b/(1-p), fp32 (M/rows_per_sample) or NULL */
  void* out; int64_t ldo;
  int64_t M; int32_t N;
  int32_t act;                                     /* 0 none, 1 GELU(erf), 2 see "training epilogues" below */
  int32_t w_layout;                                /* 0: w is (N, K) row-major (nn.Linear.weight);  1: w is (K, N) row-major, i.e.
                                                      out = a w — the data gradient dA = dZ W reads W in place, no transpose */
  /* --- LayerNorm folded into the Linear (bf16 / tcgen05 path only; all NULL / 0 by default) ---------------------------
   * LN(x) W^T + b  ==  rstd_m * (x (W o gamma)^T  -  mean_m * colsum_n)  +  (b + W beta)_n : the GEMM runs on the RAW rows
   * with the caller-derived weight w = bf16(W o gamma); ln_colsum[n] = sum_k float(w[n,k]); bias_f32 = b + W beta (fp32);
   * mean / rstd come from per-row partial sums (sum x, sum x^2) that the PRODUCER of x wrote: ln_stats is
   * (M, ln_stats_parts, 2) fp32.  stats_out: if non-NULL this launch writes the same partial sums of ITS output rows,
   * (M, gridDim.y = ceil(N / BN), 2) fp32, for the next folded Linear; cswin_linear_stats_parts() returns that count.
   * No separate LayerNorm kernel and no normalised copy of the activation exist on this path. */
  const float* ln_stats; int32_t ln_stats_parts; int32_t ln_C;
  const float* ln_colsum;
  const float* bias_f32;                           /* fp32 bias (used instead of `bias` when non-NULL) */
  float* stats_out;
  /* --- training epilogues (bf16 / tcgen05 path with 16-byte aligned rows and N % 8 == 0; otherwise CSWIN_ERR_UNSUPPORTED) --
   * aux_out: with act = 1 the pre-activation z = [a|a2] w^T + bias is ALSO written here ((M,N), row pitch ld_aux): fc1 of
   *          the Mlp keeps z for the backward and hands GELU(z) to fc2 without a separate activation pass (Mlp :22-23).
   * act = 2: out = (a w) * GELU'(residual): the data gradient of fc2 multiplied by GELU'(z) in the epilogue
   *          (`residual` carries z; it is not added). */
  void* aux_out; int64_t ld_aux;
} cswin_linear_args_t;

int cswin_linear_fwd(const cswin_linear_args_t* args, int32_t dtype, cswin_stream_t stream);
/* number of column tiles (= partial sums per row written to stats_out) the tcgen05 kernel will use for this problem */
int32_t cswin_linear_stats_parts(int64_t M, int32_t N, int32_t K, int32_t act);
/* ---- fused MLP half of a CSWinBlock (bf16 / tcgen05 only): out = x + GELU(LN(x) W1^T + b1) W2^T + b2 ------------------------
 * replaces cswin_unet.py:179 `x = x + drop_path(mlp(norm2(x)))` with Mlp.forward :22-26 (eval: DropPath = identity) in ONE
 * launch; the (M, 4C) hidden activation never reaches global memory.  LayerNorm is folded exactly as in
 * cswin_linear_args_t: w1 = bf16(W1 o gamma), ln_colsum[h] = sum_k float(w1[h,k]), b1 = fc1.bias + W1 beta (fp32),
 * ln_stats = (M, ln_stats_parts, 2) partial (sum, sum^2) of the rows of x written by the producer of x.
 * stats_out (optional): (M, cswin_mlp_stats_parts(C, hidden), 2) partial sums of the rows of out.
 * Supported: C in {64, 128, 256}, hidden = 4 C, 16-byte aligned rows; anything else returns CSWIN_ERR_UNSUPPORTED and the
 * caller composes two cswin_linear_fwd calls instead. */
typedef struct {
  const void* x; int64_t ldx;                      /* (M, C) bf16: GEMM operand AND residual */
  const void* w1; int64_t ldw1;                    /* (hidden, C) bf16, gamma-folded */
  const float* ln_colsum; const float* b1;         /* (hidden) fp32 */
  const void* w2; int64_t ldw2;                    /* (C, hidden) bf16 */
  const float* b2;                                 /* (C) fp32 */
  const float* ln_stats; int32_t ln_stats_parts; float ln_eps;
  void* out; int64_t ldo;                          /* (M, C) bf16 */
  float* stats_out;
  int64_t M; int32_t C; int32_t hidden;
} cswin_mlp_args_t;
int cswin_mlp_fwd(const cswin_mlp_args_t* args, int32_t dtype, cswin_stream_t stream);
int32_t cswin_mlp_stats_parts(int32_t C, int32_t hidden);   /* 0 if the shape is unsupported */

/* ---- [LayerNorm -> qkv Linear -> LePE attention of both branches] in one launch (bf16 / tcgen05 only) -----------------------
 * replaces, inside CSWinBlock.forward (networks/cswin_unet.py:160-181): norm1 :168, the qkv Linear and its (3,B,L,C) view
 * :169, both LePEAttention.forward calls :172-176 (:82-109 with im2cswin :59-65, get_lepe :67-80, img2windows :184-191,
 * windows2img :194-202) and the torch.cat of :174.  The (B, L, 3C) qkv tensor never exists in global memory: each CTA computes
 * the q | k | v columns of up to two heads for one 128-row window tile (gathered by TMA in window order) on the tensor cores,
 * converts them to bf16 tiles in shared memory and runs the attention of those heads from there.
 *   x        : (B, L = reso^2, C) bf16 activation, token t = y*reso + x; RAW rows when the LayerNorm is folded
 *   w        : (3C, C) bf16 rows [q | k | v] (qkv.weight, or qkv.weight o gamma when folded); bias_f32: (3C) fp32 or NULL
 *   ln_stats / ln_colsum / ln_eps: folded LayerNorm exactly as in cswin_linear_args_t (both NULL: x is used as it is)
 *   br[i]    : branch i owns channels [sum_{j<i} 32 heads_j, +32 heads_i) of q, of k, of v and of out
 *   out      : (B, L, C) bf16, both branches written into their concat position
 * Envelope (cswin_qkv_lepe_attention_supported): C in {64, 128, 192, 256}, head_dim 32, windows of <= 128 tokens, branches
 * with >= 2 heads use an even head count; otherwise CSWIN_ERR_UNSUPPORTED and the caller composes cswin_linear_fwd +
 * cswin_lepe_attention_fwd. */
typedef struct {
  const void* conv_w; const void* conv_b;          /* get_v.weight (C_b,1,3,3) / get_v.bias (C_b) of the branch, bf16 */
  int32_t heads, H_sp, W_sp, reserved;
} cswin_qkv_attn_branch_t;
typedef struct {
  const void* x; int64_t x_bs, x_ts;               /* batch / token stride of x (elements) */
  const void* w; int64_t ldw;
  const float* bias_f32;
  const float* ln_stats; const float* ln_colsum; int32_t ln_stats_parts; float ln_eps;
  void* out; int64_t o_bs, o_ts;
  int32_t B, reso, C, n_branches;
  cswin_qkv_attn_branch_t br[2];
  float scale; int32_t reserved;
} cswin_qkv_attn_args_t;
int cswin_qkv_lepe_attention_fwd(const cswin_qkv_attn_args_t* args, int32_t dtype, cswin_stream_t stream);
int32_t cswin_qkv_lepe_attention_supported(int32_t C, int32_t reso, int32_t n_branches, const int32_t* heads, const int32_t* H_sp,
                                           const int32_t* W_sp);

/* ---- the stem in one launch (bf16 / tcgen05 only): Conv2d(3, 64, 7, stride 4, padding 2) -> tokens -> LayerNorm(64) -------------
 * replaces `stage1_conv_embed` (networks/cswin_unet.py:338-342: conv :339, Rearrange :340, LayerNorm :341) — an implicit GEMM whose
 * A tile is gathered from the NCHW image inside the CTA (no column matrix), bias + LayerNorm + row statistics in the epilogue.
 *   x        : (B, 3, H, W) fp32 (x_is_f32 = 1) or bf16, contiguous
 *   w_packed : (64, 192) bf16 = conv.weight.reshape(64, 147) zero-padded to 192 columns (column (c*7 + ky)*7 + kx), 16-byte aligned
 *   bias, gamma, beta : (64) fp32 (conv.bias, LayerNorm weight / bias)
 *   out      : (B * Ho * Wo, 64) bf16 token-major, Ho = (H + 4 - 7) / 4 + 1;  stats: (B Ho Wo, 1, 2) fp32 (sum, sum^2) of the bf16 rows
 * *handled = 0 (nothing launched) when the staged input rows of a tile do not fit in shared memory (very narrow / very wide images): the caller composes
 * cswin_im2col_nchw + cswin_linear_fwd + cswin_layernorm_stats_fwd instead. */
int cswin_stem_fwd(const void* x, int32_t x_is_f32, const void* w_packed, const float* bias, const float* gamma, const float* beta,
                   float eps, void* out, float* stats, int32_t B, int32_t H, int32_t W, int32_t dtype, cswin_stream_t stream,
                   int32_t* handled);

/* ---- all CSWinBlocks of one stage in ONE persistent dataflow launch (bf16 / tcgen05 only, inference) ------------------------
 * replaces the stage loops `for blk in self.stageN: x = blk(x)` (networks/cswin_unet.py:462-478, :505-533) over
 * CSWinBlock.forward (:160-181): per block norm1 + qkv :168-169, both LePEAttention.forward calls + cat :172-176 (:82-109),
 * proj + residual :177-178, norm2 + Mlp + residual :179 (Mlp :22-26) — DropPath is the identity in eval mode.
 * The 128 x BN Linear tiles and 128-row attention tiles of cswin_linear_fwd / cswin_lepe_attention_fwd (same arithmetic) are
 * pulled from one in-order tile counter by resident CTAs and ordered by per-row-tile / per-image completion counters, so that
 * row tile m of an op starts when row tile m of its producer is complete: no launch gaps, no grid-wide barriers between the
 * 5 x n_blocks ops.  LayerNorms are folded exactly as in cswin_linear_args_t.
 *   x         : (B, L = reso^2, C) bf16 contiguous, IN: input of the first block, OUT: output of the last block
 *   stats_in  : (M = B L, stats_in_parts, 2) fp32 per-row partial (sum, sum^2) of x, written by the producer of x
 *   qkv, att, x1, hid : workspaces (M, 3C), (M, C), (M, C), (M, hidden) bf16 contiguous, 16-byte aligned
 *   stats_x, stats_x1 : (M, plan.parts_x, 2), (M, plan.parts_x1, 2) fp32; after the call stats_x holds the row statistics of x
 *   ctrl      : plan.ctrl_ints int32, ZEROED ONCE by the caller when it is allocated; the kernel leaves it zeroed.  One ctrl
 *               buffer must not be used by two launches that may run concurrently (different streams).
 *   blocks    : HOST array of n_blocks <= plan.max_blocks descriptors (device pointers inside)
 *   heads / H_sp / W_sp : the block's attention branches (branch i owns channels [32 sum_{j<i} heads_j, +32 heads_i))
 * Envelope (cswin_stage_plan returns CSWIN_ERR_UNSUPPORTED outside it): C and hidden multiples of 64, head_dim 32, stripe
 * windows of <= 128 tokens. */
typedef struct {
  const void* w_qkv; const float* cs_qkv; const float* b_qkv;     /* (3C, C) bf16 W o gamma1; (3C) fp32 column sums; (3C) fp32 b + W beta1 */
  const void* w_proj; const float* b_proj;                         /* (C, C) bf16; (C) fp32 */
  const void* w_fc1; const float* cs_fc1; const float* b_fc1;     /* (hidden, C) bf16 W o gamma2; (hidden) fp32; (hidden) fp32 */
  const void* w_fc2; const float* b_fc2;                           /* (C, hidden) bf16; (C) fp32 */
  const void* lepe_w[2]; const void* lepe_b[2];                    /* get_v.weight (C_b,1,3,3) / get_v.bias (C_b) per branch, bf16 */
  float eps1, eps2;                                                /* norm1.eps, norm2.eps */
} cswin_stage_block_t;
typedef struct {
  void* x; const float* stats_in; int32_t stats_in_parts; int32_t n_blocks;
  void* qkv; void* att; void* x1; void* hid;
  float* stats_x; float* stats_x1;
  int32_t* ctrl; int64_t ctrl_ints;
  const cswin_stage_block_t* blocks;
  int32_t B, reso, C, hidden, n_branches, reserved;
  int32_t heads[2], H_sp[2], W_sp[2];
  float scale; int32_t reserved2;
} cswin_stage_args_t;
typedef struct { int32_t parts_x, parts_x1, max_blocks, reserved; int64_t ctrl_ints; } cswin_stage_plan_t;
int cswin_stage_plan(int32_t B, int32_t reso, int32_t C, int32_t hidden, int32_t n_branches, const int32_t* heads,
                     const int32_t* H_sp, const int32_t* W_sp, cswin_stage_plan_t* plan);
int cswin_stage_fwd(const cswin_stage_args_t* args, int32_t dtype, cswin_stream_t stream);

/* ---- backward of cswin_carafe_head_fwd (the folded CARAFE4 + out + output head, up = 4; 2, 3, 4 or 9 classes) ---------------
 * dlogits: fp32 NCHW (B, C, 4H, 4W) contiguous.  Writes d enc (B*H*W, 144) and d z (B*H*W, zcols; columns >= C are zeroed),
 * accumulates d bias (C, fp32, pre-zeroed by the caller).  kws: workspace of B*H*W*144 floats (the softmaxed kernels,
 * tap-major).  The gradients of CARAFE4.out and of the `output` conv follow from d z / d bias by the chain rule of the fold
 * W_f = W_output W_out, b_f = W_output b_out (host side, two 9x64 matmuls). */
int cswin_carafe_head_bwd(const void* enc, int64_t ldenc, const void* z, int64_t ldz, const float* dlogits, void* denc,
                          int64_t lddenc, void* dz, int64_t lddz, int32_t zcols, float* dbias, float* kws, int32_t B, int32_t H,
                          int32_t W, int32_t C, int32_t up, int32_t dtype, cswin_stream_t stream);

/* ---- training loss of the baseline trainer: w_ce * CrossEntropy + w_dice * DiceLoss(softmax=True) --------------------------
 * replaces trainer.py:55-57 (`loss_ce`, `loss_dice`, 0.4 / 0.6 mix) with utils.py:9-45 (DiceLoss: per-class ratio of sums over the
 * whole LOCAL batch, smooth 1e-5) — two passes over the fp32 NCHW logits (B, C, HW) instead of ~25 element-wise kernels.
 * labels: (B, HW) uint8 / int32 / int64 (label_bytes = 1 / 4 / 8).  sums: 1 + 3C floats, ZEROED by the caller; after fwd:
 * sums[0] = sum of -log p[y], sums[1..C] = I_c, sums[1+C..2C] = Z_c, sums[1+2C..3C] = Y_c; the scalar loss is
 * w_ce * sums[0] / (B HW) + w_dice * mean_c (1 - (2 I_c + 1e-5) / (Z_c + Y_c + 1e-5)) (formed by the caller).
 * bwd: dlogits = *grad_out * d loss / d logits, grad_out a DEVICE scalar.  Supported C: 2, 3, 4, 9. */
int cswin_seg_loss_fwd(const float* logits, const void* labels, int32_t label_bytes, float* sums, int64_t B, int32_t C, int64_t HW,
                       cswin_stream_t stream);
int cswin_seg_loss_bwd(const float* logits, const void* labels, int32_t label_bytes, const float* sums, const float* grad_out,
                       float* dlogits, float w_ce, float w_dice, int64_t B, int32_t C, int64_t HW, cswin_stream_t stream);

/* ---- optimizer step of the data-parallel training path: torch.optim.SGD(momentum, weight_decay).step() -------------------
 * replaces `optimizer.step()` of trainer.py:42/:61 (SGD lr 0.05 poly-decayed, momentum 0.9, weight decay 1e-4) for ALL
 * parameters in one launch, and refreshes the bf16 copy of each weight that the next forward's tcgen05 kernels stream:
 *     g' = grad + weight_decay * p ;  m = momentum * m + g' ;  p -= lr * m ;  shadow = bf16(p)
 * (dampening 0, no Nesterov; a zero-initialised m reproduces torch's first step m = g').  `chunks` is a DEVICE array: one
 * entry per <= 65536-element piece of a parameter, pointers pre-offset; lr is read from DEVICE memory so a CUDA graph of
 * the step follows the caller's learning-rate schedule (trainer.py:63-66) without re-capture. */
typedef struct {
  float* param; const float* grad; float* momentum;
  void* shadow;                                    /* bf16 copy of param (NULL: none) */
  int64_t n;
} cswin_sgd_chunk_t;
int cswin_sgd_momentum_step(const cswin_sgd_chunk_t* chunks, int32_t n_chunks, const float* lr, float momentum,
                            float weight_decay, cswin_stream_t stream);

/* LayerNorm that also emits the (sum, sum^2) row statistics of its bf16 OUTPUT, (M, 1, 2) fp32, seeding the folded-LayerNorm
 * chain of the next CSWinBlock without a separate pass (bf16, C in {64,128,256,512}, 16-byte aligned rows; else UNSUPPORTED) */
int cswin_layernorm_stats_fwd(const void* x, int64_t ldx, const void* gamma, const void* beta, void* y, int64_t ldy,
                              int64_t M, int32_t C, float eps, float* row_stats_out, int32_t dtype, cswin_stream_t stream);

/* per-row (sum x, sum x^2) of a (M, C) activation as one part: stats (M, 1, 2) fp32 — for inputs no Linear produced */
int cswin_row_stats(const void* x, int64_t ldx, int64_t M, int32_t C, float* stats, int32_t dtype, cswin_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * im2col gathers feeding cswin_linear_fwd (convolutions as GEMMs).
 *  tokens variant: x is token-major (B, H*W, C) (what every block produces; the reference instead makes an NCHW copy,
 *                  cswin_unet.py:214,235) ; col[(b,oy,ox), (ky*KW+kx)*C + c], zero padded.  Used by Merge_Block.conv
 *                  (3x3 s2 p1, :216) and CARAFE.encoder (3x3 s1 p1, :241).
 *  nchw variant:   x is (B, C, H, W) (the network input) ; col[(b,oy,ox), (c*KH+ky)*KW + kx] padded with zeros to
 *                  ldcol columns.  Used by the stem conv (7x7 s4 p2, :339).
 * ------------------------------------------------------------------------------------------------ */
int cswin_im2col_tokens(const void* x, int64_t x_bs, int64_t x_ts, void* col, int64_t ldcol, int32_t B, int32_t H,
                        int32_t W, int32_t C, int32_t KH, int32_t KW, int32_t stride, int32_t pad, int32_t dtype,
                        cswin_stream_t stream);
int cswin_im2col_nchw(const void* x, int32_t x_is_f32, void* col, int64_t ldcol, int32_t B, int32_t C, int32_t H,
                      int32_t W, int32_t KH, int32_t KW, int32_t stride, int32_t pad, int32_t dtype,
                      cswin_stream_t stream);

/* ---- convolution over a token image as an IMPLICIT GEMM (bf16 / tcgen05 only): no column matrix --------------------------------
 * replaces the NCHW copy + nn.Conv2d of Merge_Block.conv (networks/cswin_unet.py:214-216: 3x3 stride 2 padding 1) and of
 * CARAFE.encoder (:240-241: 3x3 stride 1 padding 1) — and cswin_im2col_tokens + cswin_linear_fwd of this library:
 *   out[(b, oy, ox), n] = bias[n] + sum_{ky, kx, c} x[b, oy*stride + ky - pad, ox*stride + kx - pad, c] * w[n, (ky*KW + kx)*C + c]
 * The tcgen05 Linear kernel fetches its A operand as strided 4-D TMA boxes (64 channels, OW pixels every `stride`-th, bh rows every
 * `stride`-th, nb images) of the (B, H, W, C) image, one box per (tap, 64-channel block); padding is the TMA's out-of-bounds zero fill.
 *   x   : (B, H*W, C) token-major bf16, x_bs / x_ts = image / token strides in elements (16-byte multiples), C % 64 == 0
 *   w   : (N, KH*KW*C) bf16, row pitch ldw — conv.weight.permute(0, 2, 3, 1) flattened;  bias : (N) bf16 or NULL
 *   out : (B*OH*OW, N) bf16, row pitch ldo;  OH = (H + 2 pad - KH) / stride + 1
 * *handled = 0 (nothing launched) outside the envelope (C % 64 != 0, OW > 128, ...): the caller composes im2col + Linear. */
int cswin_conv_tokens_fwd(const void* x, int64_t x_bs, int64_t x_ts, const void* w, int64_t ldw, const void* bias, void* out,
                          int64_t ldo, int32_t B, int32_t H, int32_t W, int32_t C, int32_t N, int32_t KH, int32_t KW, int32_t stride,
                          int32_t pad, int32_t dtype, cswin_stream_t stream, int32_t* handled);

/* ------------------------------------------------------------------------------------------------
 * CARAFE content-aware reassembly.  Replaces cswin_unet.py:242-263 (pixel_shuffle, softmax over the 9 taps,
 * pad + unfold + matmul, pixel_shuffle) of CARAFE / CARAFE4 (:222-319), applied AFTER the 1x1 `out` conv (:264):
 * the conv is linear and per-pixel, so out(sum_t k_t X_t) + b = sum_t k_t out_nobias(X_t) + b; running it at low
 * resolution costs s^2 times fewer FLOPs and removes the (B, s^2 L, C) intermediate.
 *   enc  (B*H*W, 9 s^2): encoder logits, channel t*s^2 + a*s + e
 *   z    (B*H*W, C):     out-conv (no bias) of the low-resolution pixels
 *   y:   token-major (B, sH*sW, C) with leading dim ldy, or NCHW (B, C, sH, sW) when nchw_out != 0
 * ------------------------------------------------------------------------------------------------ */
int cswin_carafe_reassemble_fwd(const void* enc, int64_t ldenc, const void* z, int64_t ldz, const void* bias,
                                void* y, int64_t ldy, int32_t nchw_out, int32_t y_is_f32, int32_t B, int32_t H,
                                int32_t W, int32_t C, int32_t up, int32_t dtype, cswin_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Segmentation head.  Replaces CSWinTransformer.up_x4 (cswin_unet.py:536-544: CARAFE4 re-assembly + `out` conv + view /
 * permute + `output` conv, with the two 1x1 maps folded by the caller into z = x (W_output W_out)^T and
 * bias = W_output b_out) and, when `labels` is given, the argmax(softmax(.)) of utils.py:73-75 so that only a uint8
 * label map has to leave the GPU.
 *   enc (B*H*W, 9 up^2), z (B*H*W, C), bias (C), C <= 16
 *   logits: (B, C, up H, up W) NCHW, fp32 or the compute dtype, or NULL;  labels: (B, up H, up W) uint8, or NULL
 * ------------------------------------------------------------------------------------------------ */
int cswin_carafe_head_fwd(const void* enc, int64_t ldenc, const void* z, int64_t ldz, const void* bias, void* logits,
                          int32_t logits_is_f32, uint8_t* labels, int32_t B, int32_t H, int32_t W, int32_t C, int32_t up,
                          int32_t dtype, cswin_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Backward kernels of the training step.  The reference has no backward source: these are the autograd derivatives of
 * cswin_unet.py:160-181 (CSWinBlock: Linear / GELU / LayerNorm / residual / DropPath), :211-220 (Merge_Block conv as
 * im2col) and :232-319 (CARAFE re-assembly), validated against torch.autograd on the oracle.  Parameter gradients are
 * fp32 and ACCUMULATED into (the caller zeroes them); activation gradients use the compute dtype.
 *   Linear  y = act(a W^T + b):   dZ = cswin_act_bwd(dY, Z)   (GELU' and / or DropPath scale)
 *                                 dA = cswin_linear_fwd(dZ, W^T)          (data gradient = a forward Linear)
 *                                 dW, db = cswin_linear_wgrad(dZ, a)
 * ------------------------------------------------------------------------------------------------ */
int cswin_act_fwd(const void* z, int64_t ldz, void* out, int64_t ldo, int64_t M, int32_t N, int32_t act, int32_t dtype,
                  cswin_stream_t stream);
int cswin_act_bwd(const void* dout, int64_t ldd, const void* z, int64_t ldz, const float* sample_scale,
                  int32_t rows_per_sample, void* dz, int64_t ldo, int64_t M, int32_t N, int32_t act, int32_t dtype,
                  cswin_stream_t stream);
int cswin_linear_wgrad(const void* dz, int64_t ldz, const void* a, int64_t lda, float* dw, int64_t ldw, float* db, int64_t M,
                       int32_t N, int32_t K, int32_t dtype, cswin_stream_t stream);
/* dx = LayerNorm'(dy) [+ dx_add]; dx_add (optional, layout of dx with row pitch ld_add) is the gradient that by-passes the
 * LayerNorm through the residual connection (x + f(LN(x)), cswin_unet.py:178-179): autograd's separate accumulation add
 * becomes part of this kernel's store.  dgamma / dbeta (C) fp32 are accumulated into. */
int cswin_layernorm_bwd(const void* x, int64_t ldx, const void* dy, int64_t ldy, const void* gamma, const float* mean,
                        const float* rstd, void* dx, int64_t ldo, const void* dx_add, int64_t ld_add, float* dgamma,
                        float* dbeta, int64_t M, int32_t C, int32_t dtype, cswin_stream_t stream);
/* adjoint of cswin_im2col_tokens: dx (B, H*W, C) from dcol (B*Ho*Wo, KH*KW*C) */
int cswin_col2im_tokens(const void* dcol, int64_t ldcol, void* dx, int64_t x_bs, int64_t x_ts, int32_t B, int32_t H, int32_t W,
                        int32_t C, int32_t KH, int32_t KW, int32_t stride, int32_t pad, int32_t dtype, cswin_stream_t stream);
/* backward of cswin_carafe_reassemble_fwd / cswin_carafe_head_fwd.  dy element (b, oy, ox, c) is read at
 * dy + b*dy_sb + oy*dy_sy + ox*dy_sx + c*dy_sc (token-major or NCHW).  kappa_ws: fp32 workspace (B*H*W * up^2 * 9). */
int cswin_carafe_reassemble_bwd(const void* enc, int64_t ldenc, const void* z, int64_t ldz, const void* dy, int32_t dy_is_f32,
                                int64_t dy_sb, int64_t dy_sy, int64_t dy_sx, int64_t dy_sc, void* denc, int64_t lddenc,
                                void* dz, int64_t lddz, float* dbias, float* kappa_ws, int32_t B, int32_t H, int32_t W,
                                int32_t C, int32_t up, int32_t dtype, cswin_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CSWIN_B200_H_ */
