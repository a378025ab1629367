/*
 * pvqa.h — C-ABI of libpvqa_sm100.so, the B200 (sm_100a) hot path of PhonoVQA.
 *
 * The reference (hieunghia-pat/phoneme-VQA) is pure Python/PyTorch and has no
 * FFI of its own; every entry point below replaces a stretch of reference
 * Python that today runs as a chain of ATen kernels.  Each declaration cites
 * the reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain C: raw device pointers + int64 sizes, no torch types.
 *   - every function returns 0 on success, a pvqa_status otherwise;
 *     pvqa_last_error() gives the thread-local message.  No exceptions, no
 *     silent fallback: unsupported shapes/dtypes are errors.
 *   - all buffers are caller-allocated, contiguous row-major unless a stride
 *     argument says otherwise; kernels never allocate or retain pointers.
 *   - work is enqueued on the caller's stream (a cudaStream_t passed as void*).
 *   - index tensors are int64 exactly as the reference datasets produce them
 *     (core/data/PhonemeLaTrDataset.py:51-58).
 */
#ifndef PVQA_H_
#define PVQA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVQA_ABI_VERSION 11

typedef enum {
  PVQA_OK = 0,
  PVQA_ERR_SHAPE = 1,     /* unsupported or inconsistent dimensions          */
  PVQA_ERR_DTYPE = 2,     /* dtype enum not supported by this entry point    */
  PVQA_ERR_ALIGN = 3,     /* pointer / row not 16-byte aligned               */
  PVQA_ERR_NULL = 4,      /* required pointer is NULL                        */
  PVQA_ERR_CUDA = 5,      /* CUDA launch / runtime error (message has code)  */
  PVQA_ERR_UNSUPPORTED = 6
} pvqa_status;

typedef enum { PVQA_F32 = 0, PVQA_BF16 = 1 } pvqa_dtype;

int pvqa_abi_version(void);
const char* pvqa_last_error(void);
/* number of kernels this library has launched in this process (all threads);
 * bench.py reports the delta over the timed region as "gpu_launches". */
int64_t pvqa_launch_count(void);
/* Optional: register a device-resident uint64 step counter.  Every dropout kernel adds its value to the
 * Philox offset it was launched with, so a CUDA graph capturing a whole training step draws fresh masks on
 * each replay (the caller increments the counter between replays).  NULL (default) disables it. */
int pvqa_set_rng_step_counter(const uint64_t* device_counter);
/* Data-parallel runs: cut the persistent attention backward into k CTAs per SM (1..16, default
 * 1 = one static item range per SM).  The NCCL all-reduces that overlap the backward hold some SMs for part of a
 * launch; with k = 1 the CTAs that were meant for those SMs queue behind the collective with a whole range of work
 * (measured at 8 ranks: 235 -> 327 us per launch), with k = 4 the block scheduler spreads the finer ranges over the
 * SMs that are free.  Affects subsequent launches (and therefore CUDA graphs captured afterwards). */
int pvqa_set_attn_bwd_waves(int k);

/* ------------------------------------------------------------------------
 * K1  fused multimodal embedding
 * replaces  core/model/PhonemeLaTr.py:33-44   SpatialModule.forward (6 gathers + 5 adds)
 *           core/model/PhonemeLaTr.py:221-229 shared(ocr)+spatial, shared(q), cat, mask cat
 *           (identical copies: core/model/LaTr.py:29-39,85-97; PreSTU family with L_ocr = 0)
 *
 * out[b, 0:S_img]                 = img_feat[b]                      (already projected ViT tokens)
 * out[b, S_img + l]               = shared[ocr_ids[b,l]] + sum_t layout[t][coords[b,l,t]]
 * out[b, S_img + L_ocr + q]       = shared[q_ids[b,q]]
 * out_mask[b]                     = [1]*S_img ++ ocr_mask[b] ++ q_mask[b]      (float32)
 *
 * layout table order t = 0..5 follows coordinate columns x0,y0,x1,y1,w,h, i.e.
 * top_left_x, top_left_y, bottom_right_x, bottom_right_y, width_emb, height_emb.
 * tab_dtype: dtype of shared/layout tables; act_dtype: dtype of img_feat and out.
 * Accumulation is fp32 in the reference's association order
 * ((((((x0+y0)+x1)+y1)+w)+h) then ocr + that).
 * err_flag (optional, device int32): set to 1 if any index is out of range
 * (the reference would hit a device assert); such rows are written as zeros.
 * ------------------------------------------------------------------------ */
int pvqa_embed_mm_fwd(const void* img_feat, const int64_t* coords,
                      const int64_t* ocr_ids, const int64_t* q_ids,
                      const float* ocr_mask, const float* q_mask,
                      const void* shared_tab, const void* const* layout_tabs /* host array[6] of device ptrs */,
                      void* out, float* out_mask,
                      int64_t B, int64_t S_img, int64_t L_ocr, int64_t L_q,
                      int64_t d, int64_t V, int64_t n_pos,
                      int tab_dtype, int act_dtype, int32_t* err_flag, void* stream);

/* backward of K1: scatter-add of d_out rows into fp32 gradient tables.
 * replaces the autograd of the same reference lines (6x embedding_dense_backward
 * + 2x for `shared`, + add/cat backward).  Gradient tables are fp32, must be
 * zero-initialised (or hold a running sum) by the caller; nn.Embedding has no
 * padding_idx in the reference so pad rows DO receive gradient.
 * d_img is not produced: it is the view d_out[:, 0:S_img]. */
int pvqa_embed_mm_bwd(const void* d_out, const int64_t* coords,
                      const int64_t* ocr_ids, const int64_t* q_ids,
                      float* d_shared, float* const* d_layout_tabs /* host array[6] */,
                      int64_t B, int64_t S_img, int64_t L_ocr, int64_t L_q,
                      int64_t d, int64_t V, int64_t n_pos,
                      int act_dtype, void* stream);

/* ------------------------------------------------------------------------
 * K1' fused multi-token phoneme (target) embedding + sinusoidal PE
 * replaces  core/model/modules/phoneme_utils.py:10-21 (intended 3-table form
 *           PhonoLaTr/modules.py:40-63: onset/rhyme/tone gathers + concat)
 *           core/model/modules/transformer_utils.py:23-25 (+ pos_embedding[:, :T], dropout)
 *
 * out[b,t,:] = concat(onset[l0], rhyme[l1], tone[l2]) + pe[t]   then inverted dropout(p)
 * labels (B,T,3) int64; onset (V_o,on_dim), rhyme (V_r,rt_dim), tone (V_t,rt_dim);
 * on_dim + 2*rt_dim == d.  pe (>=T, d) fp32 (the persistent pos_embedding buffer).
 * dropout uses Philox4x32-10 keyed by (seed, offset) per element index; p = 0 disables.
 * ------------------------------------------------------------------------ */
int pvqa_embed_tgt_fwd(const int64_t* labels, const void* onset_tab, const void* rhyme_tab,
                       const void* tone_tab, const float* pe, void* out,
                       int64_t B, int64_t T, int64_t d, int64_t on_dim, int64_t rt_dim,
                       int64_t V_o, int64_t V_r, int64_t V_t,
                       int tab_dtype, int act_dtype,
                       float dropout_p, uint64_t seed, uint64_t offset,
                       int32_t* err_flag, void* stream);

int pvqa_embed_tgt_bwd(const void* d_out, const int64_t* labels,
                       float* d_onset, float* d_rhyme, float* d_tone,
                       int64_t B, int64_t T, int64_t d, int64_t on_dim, int64_t rt_dim,
                       int64_t V_o, int64_t V_r, int64_t V_t,
                       int act_dtype, float dropout_p, uint64_t seed, uint64_t offset,
                       void* stream);

/* ------------------------------------------------------------------------
 * K4  fused multi-token phoneme output head + 3x cross-entropy
 * replaces  core/model/PhonemeLaTr.py:124-130 (column split + onset/rhyme/tone Linear heads)
 *           core/executor/PhonemeLaTr_Executor.py:181-190 (3x CrossEntropyLoss(ignore_index=pad), summed)
 *
 * h (N, d) is the output of shared_lm_head (N = B*T rows).  For sub-head k with
 * column slice [off_k, off_k + w_k) of h, weight W_k (V_k, w_k), bias b_k (V_k):
 *   logits_k = h[:, slice_k] @ W_k^T + b_k ;  loss_k = mean_{n: tgt[n,k] != ignore} -log_softmax(logits_k)[tgt[n,k]]
 * loss = loss_0 + loss_1 + loss_2.  Logits never reach HBM unless logits_out != NULL
 * (the reference-compatible forward() needs them; the fused-loss fast path does not).
 *
 * fwd writes: loss_sum[3] (fp32, sum of NLL per head), count[3] (int32 non-ignored targets),
 *             lse (N,3) fp32 saved for backward.
 * bwd writes: dlogits_k (N,V_k) act dtype = grad_loss/count_k * (softmax(logits_k) - onehot(tgt))
 *             (zero on ignored rows), recomputing the logits from h; the three small
 *             GEMMs d_h = dlogits_k @ W_k, dW_k = dlogits_k^T @ h_k, db_k = sum dlogits_k are
 *             left to the caller's GEMM library (9 MB of dlogits at B=64; the logits
 *             themselves are never written).
 * ------------------------------------------------------------------------ */
int pvqa_phoneme_head_ce_fwd(const void* h, const int64_t* targets /* (N,3) strided */,
                             int64_t tgt_row_stride,
                             const void* W_onset, const void* b_onset,
                             const void* W_rhyme, const void* b_rhyme,
                             const void* W_tone, const void* b_tone,
                             float* loss_sum /*3*/, int32_t* count /*3*/, float* lse /*N,3*/,
                             void* logits_onset, void* logits_rhyme, void* logits_tone /* optional */,
                             int64_t N, int64_t d, int64_t on_dim, int64_t rt_dim,
                             int64_t V_o, int64_t V_r, int64_t V_t, int64_t ignore_index,
                             int w_dtype, int act_dtype, void* stream);

/* K4 on the Blackwell tensor path: shared_lm_head + column split + the three heads + 3x cross-entropy in ONE tcgen05
 * kernel (phoneme-vqa_b200/csrc/head_tc.cu).  replaces core/model/PhonemeLaTr.py:121-130 and
 * core/executor/PhonemeLaTr_Executor.py:181-190 on the fused-loss path.
 *   x (N,768) bf16 = decoder output;  W_shared (768,768) bf16, b_shared (768) fp32;  W_k (V_k,256) bf16, b_k (V_k) fp32
 *   writes h_out (N,768) bf16 = shared_lm_head(x) (what pvqa_phoneme_head_ce_bwd recomputes from), loss_sum[3],
 *   count[3], lse (N,3) — same meaning as pvqa_phoneme_head_ce_fwd — and, when the dl_* pointers are given, the
 *   UNSCALED logit gradients softmax - onehot (zero rows for ignored targets; rows padded to 16 columns, pad = 0):
 *   the backward needs no recompute kernel, it scales by g / count_k inside its GEMMs.  Specialised for d = 768 (slices 256|256|256) and
 *   sub-vocabularies of at most 192 entries; other shapes: PVQA_ERR_SHAPE (use pvqa_phoneme_head_ce_fwd on the output
 *   of a library GEMM). */
int pvqa_phoneme_head_fused_fwd(const void* x, const void* W_shared, const float* b_shared,
                                const int64_t* targets /* (N,3) strided */, int64_t tgt_row_stride,
                                const void* W_onset, const float* b_onset,
                                const void* W_rhyme, const float* b_rhyme,
                                const void* W_tone, const float* b_tone,
                                void* h_out, float* loss_sum /*3*/, int32_t* count /*3*/, float* lse /*N,3*/,
                                void* dl_onset, void* dl_rhyme, void* dl_tone /* all three or none: (N, round16(V_k)) bf16 */,
                                int64_t N, int64_t d, int64_t on_dim, int64_t rt_dim,
                                int64_t V_o, int64_t V_r, int64_t V_t, int64_t ignore_index, void* stream);

int pvqa_phoneme_head_ce_bwd(const void* h, const int64_t* targets, int64_t tgt_row_stride,
                             const void* W_onset, const void* b_onset,
                             const void* W_rhyme, const void* b_rhyme,
                             const void* W_tone, const void* b_tone,
                             const float* lse, const int32_t* count, const float* grad_loss /* device scalar */,
                             void* dlogits_onset, void* dlogits_rhyme, void* dlogits_tone,
                             int64_t N, int64_t d, int64_t on_dim, int64_t rt_dim,
                             int64_t V_o, int64_t V_r, int64_t V_t, int64_t ignore_index,
                             int w_dtype, int act_dtype, void* stream);

/* ------------------------------------------------------------------------
 * K4 large-vocabulary variant (LaTr): softmax + cross-entropy + gradient over a CHUNK of logits rows
 * replaces  core/model/LaTr.py:83 (lm_head logits) + core/executor/base_executor.py:169 /
 *           core/executor/LaTr_Executor.py:160-163 (CrossEntropyLoss(ignore_index=pad))
 * logits (n,V) fp32 for one chunk of rows (the host produces them chunk by chunk with its GEMM library, so
 * the full (B*T, 36096) logits / log-softmax / gradient tensors of the reference never exist);
 * writes dlogits (n,V) bf16 = (softmax - onehot(target)) * (*inv_count) (zeros on ignored rows) and adds
 * sum_n (lse_n - logit_n[target]) into *loss_sum.  inv_count = 1 / #non-ignored targets of the WHOLE batch.
 * ------------------------------------------------------------------------ */
int pvqa_vocab_ce_grad(const float* logits, const int64_t* targets, int64_t tgt_stride,
                       const float* inv_count, float* loss_sum, void* dlogits_bf16,
                       int64_t n, int64_t V, int64_t ignore_index, void* stream);

/* ------------------------------------------------------------------------
 * K2 / K3  flash attention (tcgen05 + TMEM + TMA), bf16 in, fp32 accumulate
 * K2 replaces HF T5Attention.forward as called from core/model/PhonemeLaTr.py:111-114
 *    (transformers/models/t5/modeling_t5.py:253-345: unscaled QK^T + shared
 *    bucketed relative bias + key mask, fp32 softmax, P@V)
 * K3 replaces nn.MultiheadAttention inside nn.TransformerDecoder as called from
 *    core/model/modules/transformer_utils.py:47-64 / core/model/PhonemeLaTr.py:134-144
 *    (scaled QK^T + causal -inf + FLOAT additive key_padding masks, D14 in SURVEY)
 *
 * Q (B,Sq,H,D) / K,V (B,Sk,H,D) bf16 with explicit element strides so packed
 * projections (q|k|v in one buffer) are consumed in place; D == 64.
 *   s[b,h,i,j] = scale * q_i . k_j + rel_bias[h][j - i + Sq - 1] + key_add[b][j] (+ -inf if causal and j > i)
 *   o = softmax_j(s) @ v
 * rel_bias (H, Sq+Sk-1) fp32 or NULL: T5 bucketed bias expanded over relative offsets
 *   (bucket(j-i) is a function of j-i only; expansion done on the host side of the ABI).
 * key_add (B,Sk) fp32 or NULL: additive per-key term (0 / -inf for T5 masks; +1.0 / 0.0
 *   float masks of the reference decoder).
 * lse (B,H,Sq) fp32 out: log-sum-exp per row, saved for backward.
 * scp_bucket (B,L,L) uint8 + scp_table (H,32) fp32, or NULL: the SaL family's spatial (SCP) bias
 *   (core/model/modules/SaL_utils.py:131-195,208-223): s += scp_table[h][scp_bucket[b][i-q0][j-q0]] for
 *   q0 <= i,j < q0+L (the OCR x OCR block), computed in-kernel instead of materialising (B,H,S,S) fp32.
 *   q0 and L must be multiples of 16.  Backward accumulates d_scp_table (H,32).
 * dropout_p > 0: inverted dropout on the softmax probabilities (HF modeling_t5.py:332 /
 *   nn.MultiheadAttention dropout) with an in-kernel Philox4x32-10 stream keyed by
 *   (seed, offset, b, h, i, j); p is quantised to 1/256.  Backward must get the same triple.
 * ------------------------------------------------------------------------ */
int pvqa_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse,
                  const float* rel_bias, const float* key_add,
                  int64_t B, int64_t H, int64_t Sq, int64_t Sk, int64_t D,
                  int64_t q_stride_b, int64_t q_stride_s, int64_t q_stride_h,
                  int64_t k_stride_b, int64_t k_stride_s, int64_t k_stride_h,
                  int64_t v_stride_b, int64_t v_stride_s, int64_t v_stride_h,
                  int64_t o_stride_b, int64_t o_stride_s, int64_t o_stride_h,
                  float scale, int causal,
                  float dropout_p, uint64_t seed, uint64_t offset,
                  const uint8_t* scp_bucket, const float* scp_table, int64_t scp_q0, int64_t scp_L,
                  void* stream);

/* backward.  dk, dv: bf16 with explicit strides (may point into a packed d(qkv) buffer).
 * dq_accum: fp32 (B,Sq,H,64) contiguous, ZERO-INITIALISED by the caller — every 128-key tile
 * adds its partial dQ with fp32 reductions; the caller converts/copies it to bf16.
 * d_rel_bias (H, Sq+Sk-1) fp32, accumulated (caller zero-initialises) = sum_{b,i,j: j-i fixed} dS, or NULL.
 * rel_far: 0, or a promise about rel_bias that lets the kernel skip the per-diagonal reduction far from the diagonal:
 *   rel_bias[h][r] is the same for all r with (r - (Sq-1)) >= rel_far, and the same for all r with
 *   (r - (Sq-1)) <= -rel_far (T5's bucketed bias: every offset beyond the last bucket boundary shares one table entry,
 *   HF modeling_t5.py:_relative_position_bucket).  The gradient of such a tail is then returned on ONE offset of the
 *   tail per 32x32 block — exact for any caller that sums d_rel_bias over offsets sharing a bias value, which is what
 *   the bucket scatter of the T5 table does.
 * o, d_o: forward output and its gradient (bf16, strided). */
int pvqa_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                  const float* lse, const float* rel_bias, const float* key_add,
                  float* dq_accum, void* dk, void* dv, float* d_rel_bias,
                  float* delta_ws /* (B,H,Sq) fp32 workspace: rowsum(dO*O) */,
                  int64_t B, int64_t H, int64_t Sq, int64_t Sk, int64_t D,
                  int64_t q_stride_b, int64_t q_stride_s, int64_t q_stride_h,
                  int64_t k_stride_b, int64_t k_stride_s, int64_t k_stride_h,
                  int64_t v_stride_b, int64_t v_stride_s, int64_t v_stride_h,
                  int64_t o_stride_b, int64_t o_stride_s, int64_t o_stride_h,
                  int64_t do_stride_b, int64_t do_stride_s, int64_t do_stride_h,
                  int64_t dk_stride_b, int64_t dk_stride_s, int64_t dk_stride_h,
                  int64_t dv_stride_b, int64_t dv_stride_s, int64_t dv_stride_h,
                  float scale, int causal,
                  float dropout_p, uint64_t seed, uint64_t offset,
                  const uint8_t* scp_bucket, const float* scp_table, float* d_scp_table,
                  int64_t scp_q0, int64_t scp_L, int64_t rel_far, void* stream);

/* ------------------------------------------------------------------------
 * fp32 parity-mode attention (CUDA cores, no tensor-core rounding): same score definition as
 * pvqa_attn_fwd/bwd with fp32 operands and outputs, explicit strides everywhere.
 * Used when compute_dtype = float32 so logits stay within 1e-5 of the reference.
 * delta_ws: (B,H,Sq) fp32 workspace (rowsum(dO*O), written by the dQ pass, read by the dK/dV pass).
 * ------------------------------------------------------------------------ */
int pvqa_attn_f32_fwd(const float* q, const float* k, const float* v, float* o, float* lse,
                      const float* rel_bias, const float* key_add,
                      int64_t B, int64_t H, int64_t Sq, int64_t Sk, int64_t D,
                      int64_t q_stride_b, int64_t q_stride_s, int64_t q_stride_h,
                      int64_t k_stride_b, int64_t k_stride_s, int64_t k_stride_h,
                      int64_t v_stride_b, int64_t v_stride_s, int64_t v_stride_h,
                      int64_t o_stride_b, int64_t o_stride_s, int64_t o_stride_h,
                      float scale, int causal,
                      float dropout_p, uint64_t seed, uint64_t offset,
                      const uint8_t* scp_bucket, const float* scp_table, int64_t scp_q0, int64_t scp_L,
                      void* stream);

int pvqa_attn_f32_bwd(const float* q, const float* k, const float* v, const float* o, const float* d_o,
                      const float* lse, const float* rel_bias, const float* key_add,
                      float* dq, float* dk, float* dv, float* d_rel_bias, float* delta_ws,
                      int64_t B, int64_t H, int64_t Sq, int64_t Sk, int64_t D,
                      int64_t q_stride_b, int64_t q_stride_s, int64_t q_stride_h,
                      int64_t k_stride_b, int64_t k_stride_s, int64_t k_stride_h,
                      int64_t v_stride_b, int64_t v_stride_s, int64_t v_stride_h,
                      int64_t o_stride_b, int64_t o_stride_s, int64_t o_stride_h,
                      int64_t do_stride_b, int64_t do_stride_s, int64_t do_stride_h,
                      int64_t dq_stride_b, int64_t dq_stride_s, int64_t dq_stride_h,
                      int64_t dk_stride_b, int64_t dk_stride_s, int64_t dk_stride_h,
                      int64_t dv_stride_b, int64_t dv_stride_s, int64_t dv_stride_h,
                      float scale, int causal,
                      float dropout_p, uint64_t seed, uint64_t offset,
                      const uint8_t* scp_bucket, const float* scp_table, float* d_scp_table,
                      int64_t scp_q0, int64_t scp_L, void* stream);

/* ------------------------------------------------------------------------
 * Fused glue of the transformer blocks (HBM-bound, one launch each)
 * rms_norm      replaces HF T5LayerNorm.forward (transformers modeling_t5.py:46-70), called from every
 *               T5 block reached through core/model/PhonemeLaTr.py:111-114: y = w * x * rsqrt(mean(x^2)+eps),
 *               statistics in fp32; rstd (N) saved for backward.  bwd accumulates dw (fp32, caller zeroes).
 * residual_dropout_add  replaces `hidden + dropout(sublayer)` (T5LayerSelfAttention/T5LayerFF.forward;
 *               torch TransformerDecoderLayer x + dropoutN(...)): out fp32 = hidden fp32 + dropout(update).
 *               bwd: d_update = d_out * mask / keep (d_hidden is d_out itself).
 * relu_dropout  replaces dropout(relu(wi(x))) of T5DenseActDense / TransformerDecoderLayer._ff_block;
 *               backward needs only y (y != 0 <=> x > 0 and kept).
 * Dropout: Philox4x32-10 keyed by (seed, offset + element/8), p quantised to 1/65536.
 * ------------------------------------------------------------------------ */
int pvqa_rms_norm_fwd(const void* x, const float* w, void* y, float* rstd, int64_t N, int64_t d, float eps,
                      int x_dtype, int y_dtype, void* stream);
int pvqa_rms_norm_bwd(const void* dy, const void* x, const float* w, const float* rstd,
                      const void* d_residual /* optional (x dtype): gradient of the residual path, added to dx */,
                      void* dx, float* dw, int64_t N, int64_t d, int x_dtype, int y_dtype, void* stream);
int pvqa_residual_dropout_add(const float* hidden, const void* update, float* out, int64_t n, int upd_dtype,
                              float dropout_p, uint64_t seed, uint64_t offset, void* stream);
int pvqa_residual_dropout_bwd(const float* d_out, void* d_update, int64_t n, int upd_dtype, float dropout_p,
                              uint64_t seed, uint64_t offset, void* stream);
int pvqa_relu_dropout_fwd(const void* x, void* y, int64_t n, int dtype, float dropout_p, uint64_t seed,
                          uint64_t offset, void* stream);
int pvqa_relu_dropout_bwd(const void* dy, const void* y, void* dx, int64_t n, int dtype, float dropout_p,
                          void* stream);

/* Post-norm tail of the target decoder layer, one launch per sublayer:
 *     y = LayerNorm(hidden + dropout(update))        (biased variance, eps inside the rsqrt)
 * replaces the `x = self.normN(x + self.dropoutN(block(x)))` lines of torch's nn.TransformerDecoderLayer
 * (norm_first=False), which the reference instantiates in core/model/modules/transformer_utils.py:47-64 and runs
 * from core/model/PhonemeLaTr.py:134-144.  hidden/y/z are fp32 (N, d); update is bf16 or fp32; y_lp (optional)
 * is the bf16 copy of y the next GEMM consumes; z (optional) is the pre-norm sum saved for the backward;
 * mean/rstd are (N).  The dropout mask is the one pvqa_residual_dropout_add draws for the same (seed, offset).
 * With a bf16 hidden stream (and y == NULL) the same kernel is the `x = x + sublayer; y = layernorm(x)` step of
 * the frozen ViT tower (HF ViTLayer.forward, called from core/model/PhonemeLaTr.py:220).
 * Backward: dy and/or dy_lp are the gradients of y / y_lp; d_hidden (fp32) is the residual-path gradient,
 * d_update = mask * d_hidden / keep in the update's dtype; dgamma/dbeta (d) are ACCUMULATED. */
int pvqa_add_dropout_ln_fwd(const void* hidden, int hidden_dtype /* fp32, or bf16 with a bf16 update */,
                            const void* update /* may be NULL: plain LayerNorm */, int upd_dtype, const float* gamma,
                            const float* beta, void* z /* hidden dtype */, float* y, void* y_lp, int lp_dtype,
                            float* mean, float* rstd, int64_t N, int64_t d, float eps, float dropout_p, uint64_t seed,
                            uint64_t offset, void* stream);
int pvqa_add_dropout_ln_bwd(const float* dy, const void* dy_lp, int lp_dtype, const float* z, const float* gamma,
                            const float* mean, const float* rstd, float* d_hidden, void* d_update /* may be NULL */,
                            int upd_dtype, float* dgamma, float* dbeta, int64_t N, int64_t d, float dropout_p,
                            uint64_t seed, uint64_t offset, void* stream);

/* Pre-norm step between two sublayers of a T5 block, one launch:
 *     hidden_out = hidden + dropout(update);   y = T5LayerNorm(hidden_out) * weight
 * i.e. the tail of T5LayerSelfAttention/T5LayerFF.forward (`hidden_states + self.dropout(...)`) fused with the
 * `self.layer_norm(hidden_states)` that opens the next sublayer (transformers modeling_t5.py:46-70 and the
 * T5Layer* classes; reached from core/model/PhonemeLaTr.py:111-114).  hidden/hidden_out fp32, y bf16 or fp32.
 * Backward: dy = gradient of y, d_residual (optional) = gradient reaching hidden_out past the norm;
 * d_hidden = d_residual + d(norm); d_update = mask * d_hidden / keep; dweight (d) is ACCUMULATED. */
int pvqa_add_dropout_rms_fwd(const float* hidden, const void* update, int upd_dtype, const float* weight,
                             float* hidden_out, void* y, int y_dtype, float* rstd, int64_t N, int64_t d, float eps,
                             float dropout_p, uint64_t seed, uint64_t offset, void* stream);
int pvqa_add_dropout_rms_bwd(const void* dy, int y_dtype, const float* d_residual, const float* hidden_out,
                             const float* weight, const float* rstd, float* d_hidden, void* d_update, int upd_dtype,
                             float* dweight, int64_t N, int64_t d, float dropout_p, uint64_t seed, uint64_t offset,
                             void* stream);

/* out[c] (+)= sum_r x[r][c] in fp32: the bias gradient of every biased Linear on the decoder path
 * (autograd's `grad_output.sum(0)` for nn.Linear / nn.MultiheadAttention in_proj/out_proj biases).
 * x is (N, d) bf16 or fp32 row-major, d % 8 == 0; accumulate == 0 zeroes `out` first. */
int pvqa_col_sum(const void* x, float* out, int64_t N, int64_t d, int dtype, int accumulate, void* stream);

/* dst[r][0:d] = src[r][0:d] with src fp32 contiguous (N, d) and dst bf16/fp32 rows `dst_row_stride` elements
 * apart: drops the fp32 dQ accumulator of pvqa_attn_bwd into the q slot of the packed (B,S,3,H,D) gradient that
 * the fused q/k/v projection backward consumes (the reference gets three separate .grad tensors from autograd:
 * transformers modeling_t5.py:312-336). */
int pvqa_cast_rows(const float* src, void* dst, int64_t N, int64_t d, int64_t dst_row_stride, int dst_dtype,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PVQA_H_ */
