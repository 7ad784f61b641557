/* mh_b200.h -- C ABI of the B200-native MelHuBERT hot path (libmh_b200.so).
 *
 * Every entry point takes plain device pointers + sizes and a cudaStream_t (passed as void*);
 * PyTorch (or any other host) owns all memory.  Return value: 0 on success, non-zero on error
 * with a thread-local message available from mh_last_error().  There is no CPU fallback.
 *
 * Each function names the reference code it replaces (paths relative to the reference root,
 * dlion168/Speech-SSL-Compression).  Activations are bf16, row-major, one row per frame in
 * (batch, time) order; master weights / gradients / statistics are fp32.
 */
#ifndef MH_B200_H
#define MH_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* mh_last_error(void);
int mh_version(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
long long mh_launch_count(void);

/* Dropout streams (reference: nn.Dropout / FairseqDropout at module.py:121-131, forward_multihead_attention.py:64-66).
 * A stream is cut into words of 32 consecutive elements; element i of word w is kept iff its 12-bit number
 *   R = bit-sliced Philox4x32-7(key = seed, counter = (2 w + call, *offset, site))
 * is >= round(p * 4096); survivors are scaled by 4096 / (4096 - thr) (p is quantised to 1/4096).  `offset` is an
 * optional device-resident step counter (NULL disables it) so that CUDA-graph replays draw fresh masks: bump it once
 * per step with mh_counter_add.  Forward and backward of one step must see the same counter value.  (The seed is the
 * Philox key: its round keys are derived on the host and travel as kernel parameters.)  Kernels with a dropout
 * argument need their column count to be a multiple of 32. */
int mh_set_dropout_offset_ptr(const unsigned long long* device_counter);
int mh_counter_add(unsigned long long* device_counter, unsigned long long v, void* stream);

/* ---------------------------------------------------------------------------------------
 * GEMM on tcgen05/TMEM fed by TMA:  D[M,N] = A[M,K] * B[N,K]^T  (+ epilogue)
 * Replaces torch._C._nn.linear at pytorch_code/forward_multihead_attention.py:71-76,110,233
 * (q/k/v/out projections), module.py:126-129 (fc1/fc2), model.py:109-110,148 (pre/final
 * projections) and their autograd dgrad / wgrad.
 *   a_mn = 0: A stored [M][K] (K contiguous, leading dim lda);  a_mn = 1: stored [K][M].
 *   b_mn = 0: B stored [N][K];                                   b_mn = 1: stored [K][N].
 * All leading dimensions are in elements and must be multiples of 8; pointers 16-byte aligned.
 * ------------------------------------------------------------------------------------- */
enum {
  MH_EPI_BF16 = 0,  /* D(bf16) = acc [+ bias[n]]                                              */
  MH_EPI_GELU = 1,  /* pre = bf16(acc + bias); aux_out = pre (if given); D = dropout(gelu(pre)) */
  MH_EPI_RES = 2,   /* D(bf16) = dropout(acc + bias) + aux_in[m,n]                             */
  MH_EPI_F32 = 3,   /* D(f32) += acc * (mask ? mask[m,n] : 1)   (atomic; split-K capable)      */
  MH_EPI_DGELU = 4, /* D(bf16) = acc * dropout_keep_scale * gelu'(aux_in[m,n])                 */
  MH_EPI_ADD = 5,   /* D(bf16) = acc + aux_in[m,n]                                             */
  MH_EPI_DELTA = 6  /* D(bf16) = acc; delta[m / T, n / 64, m % T] = sum over the 64 columns of a head of
                       bf16(acc)[m,n] * aux_in[m,n] -- the rowsum(dO * O) term of the attention backward
                       (pytorch_code/forward_multihead_attention.py:62, softmax backward), fused into the
                       out_proj dgrad GEMM that produces dO.  K-major A, MN-major B, 256-wide tiles only. */
};

typedef struct {
  int M, N, K;
  const void* A; long long lda; int a_mn;
  const void* B; long long ldb; int b_mn;
  void* D; long long ldd;
  int epilogue;
  const float* bias;         /* [N] or NULL */
  const void* aux_in;        /* bf16 [M, ld_aux] */
  void* aux_out;             /* bf16 [M, ld_aux] */
  long long ld_aux;
  const uint8_t* mask;       /* MH_EPI_F32: 0/1 bytes [M, ldd] or NULL */
  float p_drop; uint64_t seed; uint32_t site;   /* dropout stream (see mh_common.cuh) */
  int block_n;               /* 0 = auto; 128 / 256 = single-CTA 128 x block_n tiles; -256 = CTA-pair (cta_group::2) 256 x 256 tiles */
  int split_k;               /* 0 = auto (MH_EPI_F32 only), else number of K splits */
  float* delta;              /* MH_EPI_DELTA: f32 [M / delta_T, N / 64, delta_T], else ignored */
  int delta_T;               /* MH_EPI_DELTA: rows per batch element (frames); M %% delta_T == 0 */
} mh_gemm_args;

int mh_gemm(const mh_gemm_args* args, void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused multi-head attention (flash style, no T x T tensor in HBM).
 * Replaces pytorch_code/forward_multihead_attention.py:39-69 (_scaled_dot_product_attention)
 * and the reshapes at :199-201, :231.
 *   qkv : bf16 [B*T, 3*E] rows = frames, columns = [q | k | v], E = 64 * heads
 *   kv_len[b] : number of valid (non padded) keys of batch element b (suffix padding)
 *   out : bf16 [B*T, E]      lse : f32 [B, heads, T] (base-2 log-sum-exp of the scaled scores)
 *   dropout on the probabilities regenerated from (seed, site) in the backward.
 * ------------------------------------------------------------------------------------- */
int mh_attn_fwd(const void* qkv, const int* kv_len, void* out, float* lse, uint8_t* keep_bits, int B, int T, int heads,
                int causal, float p_drop, uint64_t seed, uint32_t site, void* stream);
/* keep_bits : 16 * ceil(T / 128) bytes per (batch, head, query) = u32 [B, heads, 4 * ceil(T / 128), T] -- the dropout
 *   decisions of the forward (word w of a query row covers keys 32 w .. 32 w + 31, key 8 g + 2 j + h of them in bit
 *   15 + 16 h - j - 4 g: the order the forward's predicate-free compare produces them in; query-minor so that 32
 *   consecutive query rows share a 128-byte line; opaque to callers), written by mh_attn_fwd when non-NULL and p_drop > 0 and REQUIRED by
 *   mh_attn_bwd when p_drop > 0 (1 bit per score instead of re-running Philox in the instruction-bound backward).
 * dqkv : bf16 [B*T, 3*E];  delta : f32 scratch [B, heads, T];  dq_acc : f32 scratch [B*T, E] */
int mh_attn_bwd(const void* qkv, const int* kv_len, const void* out, const void* dout, const float* lse,
                const uint8_t* keep_bits, float* delta, float* dq_acc, void* dqkv, int B, int T, int heads, int causal,
                float p_drop, uint64_t seed, uint32_t site, void* stream);
/* mh_attn_bwd_ex: flags & 1 = dq_acc has already been zeroed by the caller (e.g. on a side stream, overlapped with the
 * GEMMs in front of the call); flags & 2 = delta already holds rowsum(dO * O) per (batch, head, query) -- written by the
 * MH_EPI_DELTA epilogue of the out_proj dgrad GEMM that produced dout. */
/* flags & 4 = leave dQ in the fp32 workspace; the caller finishes with mh_dq_finish_colsum, which writes
 * dqkv[:, 0:E] = bf16(dq_acc) AND accumulates colsum[c] += sum over rows of dqkv[:, c] for all 3E columns (the q / k / v
 * bias gradients, fairseq_code/multihead_attention.py:151) in one pass. */
int mh_dq_finish_colsum(const float* dq_acc, void* dqkv, float* colsum, int rows, int E, void* stream);
int mh_attn_bwd_ex(const void* qkv, const int* kv_len, const void* out, const void* dout, const float* lse,
                   const uint8_t* keep_bits, float* delta, float* dq_acc, void* dqkv, int B, int T, int heads, int causal,
                   float p_drop, uint64_t seed, uint32_t site, int flags, void* stream);
/* mh_attn_bwd_bias: mh_attn_bwd_ex (flags 1 and 2) that also produces the stacked q / k / v bias gradients
 * (fairseq_code/multihead_attention.py:151: the three in_proj biases), bias_grad[0:3E] += column sums of dqkv.  The k / v
 * parts are reduced inside the attention backward from its fp32 dK / dV accumulators (a 31-shuffle transposing butterfly per
 * warp and key block), the q part in the pass that converts the fp32 dQ workspace to bf16: no kernel re-reads dK / dV. */
int mh_attn_bwd_bias(const void* qkv, const int* kv_len, const void* out, const void* dout, const float* lse,
                     const uint8_t* keep_bits, float* delta, float* dq_acc, void* dqkv, float* bias_grad, int B, int T,
                     int heads, int causal, float p_drop, uint64_t seed, uint32_t site, int flags, void* stream);

/* ---------------------------------------------------------------------------------------
 * LayerNorm family (module.py:121-123,129-131,232-236: dropout -> +residual -> LayerNorm is
 * split as GEMM epilogue (dropout, +residual) + this kernel).
 *   y = LN(x) * gamma + beta, eps; stats (mean, rstd) saved for the backward.
 *   optional dropout on the output (encoder-level F.dropout, module.py:236).
 * ------------------------------------------------------------------------------------- */
int mh_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                     int rows, int cols, float eps, float p_drop, uint64_t seed, uint32_t site, void* stream);
/* dy_eff = dy * keep-mask(p_in, seed_in, site_in)  (the dropout that was applied to the LN
 * output in the forward, if any);  dx (bf16) = LN backward of dy_eff;  dx_drop (optional)
 * = dx * keep-mask(p_out, ...) -- the gradient w.r.t. the GEMM output that fed this LN through
 * dropout + residual, ready to be the A operand of the dgrad / wgrad GEMMs.
 * dgamma / dbeta (f32 [cols]) are accumulated (atomicAdd). */
int mh_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                     void* dx, void* dx_drop, float* dgamma, float* dbeta, int rows, int cols, float p_in,
                     uint64_t seed_in, uint32_t site_in, float p_out, uint64_t seed_out, uint32_t site_out,
                     void* stream);
/* same, and dcol[c] += sum over rows of the output (dx_drop when given, else dx): the bias gradient of the linear layer
 * whose (dropped) output fed this LayerNorm's input -- saves the separate column-sum pass over that tensor */
int mh_layernorm_bwd_colsum(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                     void* dx, void* dx_drop, float* dgamma, float* dbeta, float* dcol, int rows, int cols, float p_in,
                     uint64_t seed_in, uint32_t site_in, float p_out, uint64_t seed_out, uint32_t site_out,
                     void* stream);

/* y = x * keep-mask / (1 - p), same element indexing as the GEMM epilogues (row * cols + col);
 * used where a dropout gradient is needed without an adjacent LayerNorm (pre-LN blocks). */
int mh_dropout_apply(const void* x, void* y, int rows, int cols, float p_drop, uint64_t seed, uint32_t site,
                     void* stream);

/* column sums: out[n] += sum_m x[m, n]  (bias gradients).  x bf16 [rows, ld] */
int mh_colsum(const void* x, long long ld, float* out, int rows, int cols, void* stream);

/* ---------------------------------------------------------------------------------------
 * Positional convolution (module.py:175-188 make pos_conv, :229-231 x = x + pos_conv(x)):
 * weight-normed grouped Conv1d(C, C, k = 128, padding = 64, groups = C / 48) -> SamePad -> GELU,
 * added to its input.  Implicit GEMM on tcgen05 with a sliding-window A operand (no im2col).
 *   x, y, z, dz, dy, dx : bf16 [B*T, C] (frames x channels);  v : f32 [C, 48, 128], g : f32 [128]
 *   w_fwd / w_bwd : bf16 [C/48][128][6][48][8] operand layouts written by mh_posconv_weight_prep
 *   (w_bwd = taps flipped, in/out swapped, for the input gradient);  norm : f32 [128] = ||v[:, :, k]||
 *   normsq_ws / dot_ws : f32 [128] scratch.
 * forward : z = conv(x) + bias (saved),  y = x + gelu(z)
 * backward: dz = dy * gelu'(z) (mh_gelu_bwd_mul);  dx = dy + conv^T(dz) (mh_posconv_dgrad);
 *           dw[C,48,128] (f32) += dz^T * window(x) (mh_posconv_wgrad);  dv += ..., dg += ... through the
 *           weight norm (mh_posconv_weight_bwd);  the bias gradient is mh_colsum(dz).
 * ------------------------------------------------------------------------------------- */
int mh_posconv_weight_prep(const float* v, const float* g, void* w_fwd, void* w_bwd, float* normsq_ws, float* norm,
                           int C, int groups, int ktaps, void* stream);
int mh_posconv_fwd(const void* x, const void* w_fwd, const float* bias, void* z, void* y, int B, int T, int C,
                   int groups, int ktaps, void* stream);
int mh_posconv_dgrad(const void* dz, const void* w_bwd, const void* dy, void* dx, int B, int T, int C, int groups,
                     int ktaps, void* stream);
int mh_gelu_bwd_mul(const void* dy, const void* z, void* dz, long long n, void* stream);
int mh_posconv_wgrad(const void* dz, const void* x, float* dw, int B, int T, int C, int groups, int ktaps,
                     void* stream);
int mh_posconv_weight_bwd(const float* dw, const float* v, const float* g, const float* norm, float* dot_ws, float* dv,
                          float* dg, int C, int groups, int ktaps, void* stream);

/* ---------------------------------------------------------------------------------------
 * Weight preparation W0: fp32 master (+ optional bool mask) -> bf16 operand, optionally also
 * its transpose (for dgrad).  Replaces the per-forward prune hook
 * pytorch_code/prune.py:24-38,64-85 (weight_orig.masked_fill(~mask, 0)) x 144 tensors and the
 * autocast casts.
 *   src f32 [rows, cols]; mask u8 [rows, cols] or NULL; dst bf16 [rows, ld_dst];
 *   dst_t bf16 [cols, ld_dst_t] or NULL.
 * ------------------------------------------------------------------------------------- */
int mh_weight_prep(const float* src, const uint8_t* mask, void* dst, long long ld_dst, void* dst_t,
                   long long ld_dst_t, int rows, int cols, void* stream);
/* f32 vector (+mask) -> f32 vector (bias prep: masked_fill) */
int mh_bias_prep(const float* src, const uint8_t* mask, float* dst, int n, void* stream);
/* generic casts */
int mh_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream);
int mh_cast_bf16_to_f32(const void* src, float* dst, long long n, void* stream);

/* ---------------------------------------------------------------------------------------
 * Input masking + padding (model.py:80 x[mask_indices] = 0; module.py:226-227 x[pad] = 0):
 *   dst(bf16)[r, :] = zero[r] ? 0 : src(f32)[r, :]
 * ------------------------------------------------------------------------------------- */
int mh_mask_rows_f32_to_bf16(const float* src, const uint8_t* zero_row, void* dst, int rows, int cols,
                             void* stream);
int mh_zero_rows_bf16(void* x, const uint8_t* zero_row, int rows, int cols, void* stream);

/* ---------------------------------------------------------------------------------------
 * Masked-frame selection (model.py:147-150): index list of rows with sel[r] != 0 in row-major
 * (b, t) order, count written to *count (device).  idx has room for `rows` entries; unused
 * tail entries are set to -1.
 * ------------------------------------------------------------------------------------- */
int mh_select_rows(const uint8_t* sel, int* idx, int* count, int rows, void* stream);
/* dst[i, :] = src[idx[i], :] for i < n_idx (idx[i] < 0 -> zeros) */
int mh_gather_rows(const void* src, const int* idx, void* dst, int n_idx, int cols, void* stream);
/* dst[idx[i], :] += src[i, :]  (backward of the gather; rows are unique) */
int mh_scatter_rows_add(const void* src, const int* idx, void* dst, int n_idx, int cols, void* stream);
int mh_gather_labels(const long long* label, const int* idx, long long* dst, int n_idx, void* stream);

/* ---------------------------------------------------------------------------------------
 * Criteria.
 * CE (upstream/melhubert/pretrain_expert.py:25,116-117): CrossEntropyLoss(ignore_index=-100,
 * mean) fused forward + backward over bf16 logits [n_rows, n_class].
 *   acc[0] += sum of row losses, acc[1] += number of non-ignored rows;
 *   dlogits (bf16) = (softmax - onehot) * grad_scale[0]   (grad_scale lives on the device so
 *   that the global-mean normaliser can come from an all-reduce without a host sync).
 *   Rows >= *n_valid (device, optional) are ignored.
 * ------------------------------------------------------------------------------------- */
int mh_ce_fwd(const void* logits, const long long* labels, const int* n_valid, float* row_loss, float* acc,
              int n_rows, int n_class, void* stream);
int mh_ce_bwd(const void* logits, const long long* labels, const int* n_valid, const float* grad_scale,
              void* dlogits, int n_rows, int n_class, void* stream);
/* loss[0] = weight * acc[0] / acc[1];  grad_scale[0] = weight / acc[1]  (device side) */
int mh_ce_finalize(const float* acc, float weight, float* loss, float* grad_scale, void* stream);
/* KD (distillation/pretrain_expert.py:83-92): hard = CE(student), soft = KL(softmax(t/T) ||
 * softmax(s/T)) summed over rows; acc[0] += sum CE_s, acc[1] += count, acc[2] += sum KL,
 * acc[3] += sum CE_t, acc[4] += rows.  Backward: dlogits = w_hard[0]*(p_s - onehot) +
 * w_soft[0]*(p_sT - p_tT)/T with device-side weights. */
int mh_kd_fwd(const void* s_logits, const void* t_logits, const long long* labels, const int* n_valid, float T,
              float* acc, int n_rows, int n_class, void* stream);
int mh_kd_bwd(const void* s_logits, const void* t_logits, const long long* labels, const int* n_valid, float T,
              const float* w_hard, const float* w_soft, void* dlogits, int n_rows, int n_class, void* stream);
/* out[0] = total, out[1] = hard CE, out[2] = soft KL(batchmean), out[3] = teacher CE;
 * w_hard[0] = (1-alpha)/count, w_soft[0] = alpha/rows */
int mh_kd_finalize(const float* acc, float alpha, float* out, float* w_hard, float* w_soft, void* stream);
/* L1 + cosine per-frame criterion (north_star; DistilHuBERT form):
 *   acc[0] += sum |p - t|, acc[1] += sum -logsigmoid(cos(p, t));  rows of `cols` bf16.
 *   backward: dpred = w_l1[0]*sign(p - t) + w_cos[0]*d(-logsigmoid(cos))/dp */
int mh_l1cos_fwd(const void* pred, const void* target, float* acc, int rows, int cols, void* stream);
int mh_l1cos_bwd(const void* pred, const void* target, const float* w_l1, const float* w_cos, void* dpred,
                 int rows, int cols, void* stream);

/* ---------------------------------------------------------------------------------------
 * Pruning objects.
 * Global magnitude threshold (pytorch_code/prune.py:553-573 topk(k, largest=False) over the
 * concatenation of all prunable tensors): exact k-th smallest |w| by 4-pass radix select over
 * the fp32 bit patterns.  ptrs/sizes describe the tensors (device array of device pointers).
 *   result[0] = threshold bits, result[1] = #elements strictly below, result[2] = #equal.
 * mh_apply_threshold_mask then clears mask bytes: |w| < thr always; |w| == thr for the first
 * `n_ties` elements in flat order ("lowest flat index wins").
 * ------------------------------------------------------------------------------------- */
int mh_abs_kth_smallest(const float* const* ptrs, const long long* sizes, int n_tensors, long long k,
                        unsigned long long* workspace /* >= 65536+8 u64 */, unsigned long long* result,
                        void* stream);
int mh_apply_threshold_mask(const float* w, uint8_t* mask, long long n, const unsigned long long* result,
                            long long k, unsigned long long* tie_counter, void* stream);
/* L1 scores: out[r] = sum_c |w[r, c]| for r < rows (fp64 accumulate), w f32 [rows, ld] */
int mh_row_abs_sums(const float* w, long long ld, double* out, int rows, int cols, void* stream);
int mh_col_abs_sums(const float* w, long long ld, double* out, int rows, int cols, void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused Adam step (runner.py:411-427: grad /= n; clip_grad_norm_; Adam; zero_grad) over a
 * flat fp32 parameter buffer.  sumsq: device scalar holding the global grad sum of squares
 * (from mh_sumsq); clip applied as min(1, max_norm / (sqrt(sumsq)*grad_scale + 1e-6)).
 * bf16_shadow (optional, same indexing as param): bf16 copy of the updated parameters, written in the same
 * pass -- the next step's GEMM operands, replacing the per-step casts of runner.py:363's autocast.
 * ------------------------------------------------------------------------------------- */
int mh_sumsq(const float* x, long long n, float* out, void* stream);
/* Weight-pruning mode (pytorch_code/prune.py:24-38: weight = weight_orig.masked_fill(~mask, 0) before every forward,
 * gradients of pruned elements zeroed by its backward): `mask` is one byte per element of the flat buffer (0 = pruned,
 * 4-byte aligned).  The masked variants ignore pruned gradients (norm and update) and write the EFFECTIVE parameters
 * param * mask as the bf16 operand shadow and, optionally, as an fp32 copy (bias operands) -- replacing the 144
 * per-forward mask applications.  mh_flat_effective builds the same two copies outside the optimizer. */
int mh_sumsq_masked(const float* x, const uint8_t* mask, long long n, float* out, void* stream);
int mh_adam_step_masked(float* param, float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                        float beta2, float eps, float weight_decay, const unsigned long long* step, float grad_scale,
                        float max_norm, const float* sumsq, int zero_grad, void* bf16_shadow, const uint8_t* mask,
                        float* effective, void* stream);
int mh_flat_effective(const float* param, const uint8_t* mask, void* bf16_shadow, float* effective, long long n, void* stream);
int mh_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                 float beta2, float eps, float weight_decay, const unsigned long long* step /* device, 1-based */, float grad_scale,
                 float max_norm, const float* sumsq, int zero_grad, void* bf16_shadow, void* stream);

/* ---------------------------------------------------------------------------------------
 * Data-parallel gradient exchange over NVLink peer memory -- replaces the reduce_add_coalesced / broadcast_coalesced
 * that nn.DataParallel runs per step (upstream/melhubert/pretrain_expert.py:28-30, runner.py:372-373).
 *   grads[p], flags[p] (p < world): device pointers, valid on THIS device, to rank p's flat fp32 gradient buffer and to
 *   its flag array (8 x u64, zero-initialised); entry `rank` is the local one.  state: local 2 x u64, zero-initialised.
 *   A bucket [start, start + count) (multiples of 4 elements) is cut into `world` shards:
 *     mh_peer_reduce_scatter : this rank's shard <- sum over ranks 0..world-1 (fixed order), pulled over NVLink
 *     mh_peer_all_gather     : every other shard <- the owner's reduced copy
 *     mh_peer_barrier_sum    : barrier across ranks (call before the optimizer reuses the buffer); with n > 0 also
 *                              vals[0..n) <- sum over ranks (n <= 16; mailboxes[p]: rank p's 2 x 8 x 16 float mailbox)
 *   All ranks must issue the same sequence of these three calls.  ctas <= 0 picks a default; the kernels are sized to
 *   co-reside with the persistent GEMM / attention CTAs (128 threads, <= 56 registers, no shared memory).
 * ------------------------------------------------------------------------------------- */
int mh_peer_reduce_scatter(void* const* grads, void* const* flags, void* state, long long start, long long count,
                           int rank, int world, int ctas, void* stream);
int mh_peer_all_gather(void* const* grads, void* const* flags, void* state, long long start, long long count,
                       int rank, int world, int ctas, void* stream);
/* Copy-engine variants (default transport): the same exchange with the NVLink pulls issued as cudaMemcpyAsync on the
 * peer-mapped pointers (DMA, no SMs) and one light local kernel that adds the staged shards in rank order.
 * staging: local scratch of >= (world - 1) * (shard length rounded up to 32) floats. */
int mh_peer_reduce_scatter_ce(void* const* grads, void* const* flags, void* state, float* staging, long long staging_elems,
                              long long start, long long count, int rank, int world, void* stream);
int mh_peer_all_gather_ce(void* const* grads, void* const* flags, void* state, long long start, long long count, int rank,
                          int world, void* stream);
int mh_peer_barrier_sum(void* const* grads, void* const* flags, void* const* mailboxes, void* state, float* vals, int n,
                        int rank, int world, void* stream);

/* ---------------------------------------------------------------------------------------
 * Front-end: batched Kaldi-compatible log-mel filterbank, replacing the host-side
 * torchaudio.compliance.kaldi.fbank call of extract_feature.py:32-53 and s3prl_upstream/expert.py:23-43
 * (num_mel_bins=40, sample_frequency=16000, window_type='hamming', frame_length=25, frame_shift=10 on the
 * waveform x 2^15; torchaudio defaults otherwise: dither 0, preemphasis 0.97, remove_dc_offset, snip_edges,
 * 512-point FFT, power spectrum, log(max(e, FLT_EPSILON)), no energy) and the (y - mean) / std normalisation
 * that follows it (extract_feature.py:42-44).
 *   wave        : f32 [batch, ld_wave] zero-padded waveforms in [-1, 1);  n_samples : i32 [batch]
 *   mel_weights : f32 [n_mel, 257] (get_mel_banks(...) padded with a zero column, built on the host)
 *   mean, inv_std : f32 [n_mel] or both NULL;  window_type : 0 hamming, 1 hanning, 2 povey, 3 rectangular
 *   out         : f32 [batch, max_frames, n_mel]; frame f of utterance b exists for
 *                 f < 1 + (n_samples[b] - frame_len) / frame_shift, the rest is written as 0.
 * ------------------------------------------------------------------------------------- */
int mh_fbank(const float* wave, long long ld_wave, const int* n_samples, int batch, const float* mel_weights,
             const float* mean, const float* inv_std, float* out, int max_frames, int n_mel, int frame_len,
             int frame_shift, float scale, float preemph, int window_type, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MH_B200_H */
