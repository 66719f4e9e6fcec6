/*
 * deepsc_b200.h - C ABI of libdeepsc_b200.so, the sm_100a implementation of the
 * DeepSC-GAN transmit path (encode -> channel(+attack) -> decode -> BLEU).
 *
 * The reference (jiang99999/DeepSC-GAN) is TensorFlow/Keras Python with no FFI of
 * its own; its seam is the Keras layer call.  Each entry point below replaces the
 * library ops behind one reference call site (cited as DeepSC-GAN/<file>:<line>);
 * INTEGRATION.md shows the ctypes stub a maintainer would bind them with.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer (caller-owned; the library never allocates,
 *    frees or retains pointers past the call); fp32 row-major unless stated;
 *  - `ld*` are leading dimensions in elements; rows of 128 floats must be 16-byte aligned;
 *  - `stream` is a cudaStream_t; every call is asynchronous on it and re-entrant;
 *  - return 0 on success, a negative dsc_status otherwise; dsc_last_error() gives the
 *    thread-local message.  Nothing throws across the ABI.  There is no CPU fallback.
 *  - the "star tile" layout is [sentences][32][128]: rows 0..30 are the satellite nodes
 *    (tokens, sequence length is fixed at 31, dataset/dataloader.py:11) and row 31 is the
 *    relay node s.
 */
#ifndef DEEPSC_B200_H
#define DEEPSC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSC_D_MODEL 128
#define DSC_HEADS 8
#define DSC_SEQ 31        /* satellites per sentence */
#define DSC_TILE_ROWS 32  /* satellites + relay */
#define DSC_SYMS 16       /* channel uses per token */

typedef enum {
  DSC_OK = 0,
  DSC_ERR_BAD_ARG = -1,    /* null pointer, bad size, misaligned leading dimension */
  DSC_ERR_CUDA = -2,       /* launch or runtime error; message holds cudaGetErrorString */
  DSC_ERR_UNSUPPORTED = -3 /* shape outside what the kernels are written for */
} dsc_status;

/* library identity ------------------------------------------------------------------ */
int dsc_version(void);               /* 10000*major + 100*minor + patch */
const char* dsc_last_error(void);    /* thread-local, never NULL */
int dsc_device_arch(void);           /* 10*major + minor of the current device, or <0 */

/* K1: embedding gather * sqrt(128) + positional row.
 * replaces Embedding + scale + pos_encoding, models/modules.py:497-502, 542-544, 661-666, 706-708.
 * ids are read at ids[s*ids_stride + i], i < len; token i of sentence s gets position pos0+i;
 * out row (s*len + i) has leading dimension ld_out. */
int dsc_embed(const int32_t* ids, int64_t ids_stride, const float* table, int vocab,
              const float* pos_table, float* out, int64_t ld_out,
              int n_sent, int len, int pos0, void* stream);

/* K2/K7/K10/K11/K12: y = act(x @ w + bias), the tf.keras.layers.Dense call
 * (models/modules.py:35-39,110-121,536; models/transceiver.py:89-90,103-105; models/gan.py:7-8).
 * w is Keras layout [K, N] with leading dimension ldw (ldw % 4 == 0, columns [N,ldw) readable).
 * act: 0 none, 1 relu.  If row_mod > 0, rows with (row % row_mod) == row_skip are not stored.
 * prec: 0 = fp32 FFMA kernel; 1 = tcgen05 bf16x3 split (fp32-class accuracy); 2 = tcgen05 bf16. */
int dsc_linear(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias,
               float* y, int64_t ldy, int M, int K, int N, int act,
               int row_mod, int row_skip, int prec, void* stream);

/* tcgen05 Dense path.  dsc_pack_weight splits a Keras-layout fp32 weight [K, N] into bf16 hi/lo planes laid
 * out as the K-major, 128-byte-swizzled shared-memory image the UMMA descriptors expect
 * (dsc_packed_weight_bytes(K, N) bytes, 128-byte aligned, K % 64 == 0); the caller owns and caches the blob.
 * dsc_linear_tc is dsc_linear on those planes: prec 1 = three bf16 passes (x_hi*W_hi + x_lo*W_hi + x_hi*W_lo,
 * fp32 accumulation in TMEM), prec 2 = one bf16 pass.  K % 128 == 0. */
int64_t dsc_packed_weight_bytes(int K, int N);
int dsc_pack_weight(const float* w, int64_t ldw, int K, int N, void* blob, void* stream);
int dsc_linear_tc(const float* x, int64_t ldx, const void* packed_w, const float* bias,
                  float* y, int64_t ldy, int M, int K, int N, int act,
                  int row_mod, int row_skip, int prec, void* stream);

/* K6: out = LN_b(2 * LN_a(x + res)) (gamma_b != NULL) or LN_a(x + res); eps = 1e-6, biased variance.
 * replaces LayerNormalization call sites models/modules.py:310-314, 354, 382-386, 425-429, 458-467;
 * models/transceiver.py:112.  Rows are addressed in groups: logical row r -> group r / group_rows,
 * member r % group_rows, element offset group*<x|res|out>_group_stride + member*128. */
int dsc_add_layernorm(const float* x, int64_t x_group_stride, const float* res, int64_t res_group_stride,
                      const float* gamma_a, const float* beta_a, const float* gamma_b, const float* beta_b,
                      float* out, int64_t out_group_stride, int n_rows, int group_rows, void* stream);

/* star tile helpers: pack [n_sent,31,128] into the tile layout and set row 31 = mean over the
 * 31 rows (s = reduce_mean(h, axis=1), models/modules.py:286,359). */
int dsc_star_pack(const float* src, float* tile, int n_sent, void* stream);

/* K3: satellite attention of one star cycle (models/modules.py:289-299, 361-372), deduplicated:
 * qkv  [n_sent*32, 384] = tile @ [wq|wk|wv] of multi_att_satellite (row 31 carries k_s, v_s);
 * kv_e [n_sent*32, 256] = e-tile @ [wk|wv] (constant over cycles);
 * att  [n_sent*32, 128]: rows 0..30 = sum_j softmax_j(q.k_j/4) v_j over the five keys
 * {h[i+1], h[i], h[i-1], e[i], s} (cyclic neighbours, no mask); row 31 is zero-filled. */
int dsc_star_satellite_attn(const float* qkv, const float* kv_e, float* att, int n_sent, void* stream);

/* The fused tcgen05 star-cycle kernel (dsc_star_cycles_tc below) reads per-row data in the "interleaved tile" layout
 * [tile][k/4][row][4 floats] (a tile = 4 sentences = 128 rows = 128 TMEM lanes), so that a warp's 32 rows read 512
 * contiguous bytes; dsc_star_interleave converts a row-major [n_groups*group_rows, width] tensor (group_rows 128, or
 * 32 for the h2 key cache) into it.  n_sent % 4 == 0. */
int dsc_star_interleave(const float* src, int64_t src_group_stride, float* dst, int n_groups, int group_rows,
                        int width, void* stream);
/* vals [n_sent,256] (k|v of one h2 row per sentence) -> row `row_index` of kv2 [n_sent][64][32][4]. */
int dsc_star_kv2_put(const float* vals, float* kv2, int row_index, int n_sent, void* stream);

/* K2+K3+K4, all cycles in one launch: the loop of models/modules.py:287-306 / 360-378 with the tile state (X, ATT, s, q)
 * resident in tensor / shared memory and the weights streamed from L2 through a shared-memory ring.
 * x_tile0 [n_sent/4][32][128][4] = the e tile (cycle-0 node states, interleaved); s0 [n_sent,128] = mean row;
 * q0 [n_sent,128] = s0 @ wq_relay; kv_e [n_sent/4][64][128][4] = k|v of the e rows under the satellite weights (interleaved); kv2 [n_sent][64][32][4] = the cached
 * k|v of the h2 rows under the relay weights, of which the first n2 rows are attended (decoder only; n2 = 0, kv2 NULL otherwise); the five packed weights are
 * dsc_pack_weight images of the grouped [128,384] satellite projection, wo_satellite, [wk|wv]_relay, wo_relay, wq_relay
 * (pass the satellite matrices again for the layers that drive the relay with the satellite weights, :175, :243).
 * x_rowmajor [n_sent][32][128] receives the tile after n_cycles cycles (rows 0..30 = h, row 31 = s).
 * prec | DSC_STAR_FIRST_SAT_DONE: the satellite half of the FIRST cycle (:287-300) - which depends on the e tile only,
 * not on kv2 - was computed before (rows 0..30 of a 1-cycle call's output, re-interleaved): x_tile0 holds that X' and
 * the first cycle runs its relay half only.  A greedy decoder pays the satellite half once per batch, not per step.
 * prec | DSC_STAR_NO_FINAL_RELAY: the caller reads the satellite rows only (a greedy decoder: utlis/eval.py:112 takes
 * predictions[:, -1:]), so the relay half of the LAST cycle is not run: row 31 of x_rowmajor is the relay node before its last update. */
#define DSC_STAR_FIRST_SAT_DONE 0x100
#define DSC_STAR_NO_FINAL_RELAY 0x200
/* Kernel form.  The library has one form, the one-tile kernel (a CTA works on one 4-sentence tile at a time);
 * DSC_STAR_FORM_ONE_TILE names it explicitly.  DSC_STAR_FORM_TWO_TILE is honoured by the debug-tools library only
 * (libdeepsc_b200_debug.so: an experimental kernel that keeps two tiles per CTA half a cycle apart, bit-identical
 * results); the product library rejects it with DSC_ERR_BAD_ARG. */
#define DSC_STAR_FORM_ONE_TILE 0x400
#define DSC_STAR_FORM_TWO_TILE 0x800
int dsc_star_cycles_tc(const float* x_tile0, const float* s0, const float* q0, const float* kv_e,
                       const float* kv2, int n2,
                       const void* packed_wqkv_grouped, const void* packed_wo, const void* packed_wkv_relay,
                       const void* packed_wo_relay, const void* packed_wq_relay,
                       const float* bias_o, const float* bias_o_relay,
                       float* x_rowmajor, int n_sent, int n_cycles, int prec, void* stream);

/* Tail of the causal target branch of a star decoder step (models/modules.py:352-354 and the relay k|v projection of
 * :375-377), fused: h2 = LayerNorm(resid + attn @ wo + bias_o; gamma, beta), k|v = h2 @ [wk|wv]_relay, written to row
 * `row_index` of the interleaved key cache kv2 [M][64][32][4] (NULL: skipped) and/or row-major to kv_rows [M,256]
 * (NULL: skipped); h2_out [M,128] optional.  attn = attention output before the dense layer.  packed_* =
 * dsc_pack_weight images of the [128,128] / [128,256] kernels. */
int dsc_target_tail_tc(const float* attn, int64_t ld_attn, const float* resid, int64_t ld_resid,
                       const void* packed_wo, const float* bias_o, const float* gamma, const float* beta,
                       const void* packed_wkv_relay, float* kv2, int row_index,
                       float* kv_rows, int64_t ld_kv, float* h2_out, int64_t ld_h2,
                       int M, int prec, void* stream);

/* K4: relay attention of one star cycle (models/modules.py:303-306, 375-378).
 * qkv_r [n_sent*32, 384] = updated tile @ [wq|wk|wv] of the relay weights; the query is row 31,
 * keys/values are row 31 (s), rows 0..30 (h) and, for the decoder, the first n2 rows of
 * kv2 [n_sent, kv2_rows, 256] (k|v of h2 under the relay weights).  out [n_sent, 128]. */
int dsc_star_relay_attn(const float* qkv_r, const float* kv2, int kv2_rows, int n2,
                        float* out, int n_sent, void* stream);

/* K5: softmax(q k^T / 4 + mask * -1e9) v for 8 heads of depth 16 (sublayer1.scale_dot_product_attention,
 * models/modules.py:41-76).  q [n, lq, 128] (ldq between rows, q_batch_stride between sentences),
 * k,v [n, lk, 128] likewise.  Masking is the sum of: a dense 0/1 float mask read at
 * mask[b*mask_b_stride + i*mask_q_stride + j] (NULL = none), key_ids (NULL = none) masking
 * keys whose id is 0 (create_padding_mask :757), and causal != 0 masking j > q_off + i
 * (create_look_ahead_mask :761).  lq <= 64, lk <= 64. */
int dsc_mha_attention(const float* q, int64_t ldq, int64_t q_batch_stride,
                      const float* k, const float* v, int64_t ldkv, int64_t kv_batch_stride,
                      float* out, int64_t ldo, int64_t o_batch_stride,
                      const float* mask, int64_t mask_b_stride, int64_t mask_q_stride,
                      const int32_t* key_ids, int64_t key_ids_stride, int causal, int q_off,
                      int n, int lq, int lk, void* stream);

/* K8: per-unit sum of squares; unit u covers elems_per_unit consecutive floats
 * (the whole-batch reduce_mean(x^2) of models/transceiver.py:91 and models/gan.py:9). */
int dsc_unit_sumsq(const float* x, float* sumsq, int n_units, int64_t elems_per_unit, void* stream);

/* K8: out = x / sqrt(factor * sumsq[u] / elems_per_unit): the power-norm Lambda of
 * models/transceiver.py:91 (factor 1) and models/gan.py:9 (factor 2). */
int dsc_power_normalize(const float* x, const float* sumsq, float factor, float* out,
                        int n_units, int64_t elems_per_unit, void* stream);

/* K9: the fused channel (models/transceiver.py:25-33 awgn, :35-83 fading; utlis/eval.py:90-93).
 * For element e of unit u (elems_per_unit floats, adjacent pairs are I/Q):
 *   xs = x / sqrt(x_factor * x_sumsq[u] / elems_per_unit)        (x_sumsq NULL: xs = x)
 *   z  = noise[e] (unit normal) or Philox4x32-10 + Box-Muller(seed, e)  (noise NULL)
 *   ps = p[e] * p_scale[u] / sqrt(p_factor * p_sumsq[u] / elems_per_unit)   (p NULL: 0; p_sumsq NULL: no divide)
 *   AWGN   (h NULL): y = xs + n_std[u]*z + ps
 *   fading (h[u] = (re, im)): y = xs*h + n_std[u]*z  (complex); detector 1 = LS  y*conj(h)/|h|^2,
 *                             2 = MMSE y*conj(h)/(|h|^2 + 2 n_std^2), 0 = return y (reference :74-75)
 * xs is also written to x_norm when x_norm != NULL (Channel_Encoder's return value). */
int dsc_channel(const float* x, const float* x_sumsq, float x_factor,
                const float* noise, uint64_t seed, uint64_t offset,
                const float* p, const float* p_sumsq, float p_factor, const float* p_scale,
                const float* h, const float* n_std, int detector,
                float* y, float* x_norm, int n_units, int64_t elems_per_unit, void* stream);

/* K12+K13: ids[r*ids_stride] = argmax_j (x[r] . w[:, j] + bias[j]), j < N, smallest index among ties
 * (tf.argmax, utlis/eval.py:112-113).  Logits are not materialised unless logits != NULL
 * (then logits[r*ld_logits + j] is stored as well, the `predictions` tensor of Transeiver*.call). */
int dsc_vocab_argmax(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias,
                     int32_t* ids, int64_t ids_stride, float* logits, int64_t ld_logits,
                     float* workspace, int64_t workspace_floats,
                     int M, int N, int prec, void* stream);
int64_t dsc_vocab_argmax_workspace(int M, int N);   /* floats */

/* K12+K13 on tcgen05: same result as dsc_vocab_argmax with the projection on the tensor cores (prec 1 = bf16x3,
 * 2 = bf16) and the logits never leaving the SM: every CTA keeps a running (max, first index) per row over its
 * range of 64-column vocabulary tiles; the last CTA of a 256-row block folds the per-range partials in ascending
 * column order (first maximum wins, like tf.argmax).  packed_w = dsc_pack_weight of the [128, N] kernel.
 * workspace: dsc_vocab_argmax_tc_workspace(M, N) BYTES, 16-byte aligned, zero-filled by the caller before the FIRST
 * call (its arrival counters reset themselves, so it can be reused by later calls of the same shape). */
int64_t dsc_vocab_argmax_tc_workspace(int M, int N);   /* bytes */
int dsc_vocab_argmax_tc(const float* x, int64_t ldx, const void* packed_w, const float* bias,
                        int32_t* ids, int64_t ids_stride, void* workspace, int64_t workspace_bytes,
                        int M, int N, int prec, void* stream);

/* K13 standalone: row-wise argmax of materialised logits. */
int dsc_argmax_rows(const float* logits, int64_t ld, int32_t* ids, int64_t ids_stride, int M, int N, void* stream);

/* K14: masked sparse CE (loss_function, models/modules.py:738-755): per-row
 * (logsumexp(logits[r]) - logits[r][target[r]]) * (target[r] != 0) into row_loss[r]. */
int dsc_masked_ce_rows(const float* logits, int64_t ld, const int32_t* target, float* row_loss,
                       int M, int N, void* stream);

/* K16: BLEU n-gram counts (utlis/tools.py:15-24,37-43 restated on ids).
 * ref [n, ref_len], hyp [n, hyp_len] int32; counts [n,10] = match_1..4, total_1..4, hyp_len, ref_len. */
int dsc_bleu_counts(const int32_t* ref, int ref_len, const int32_t* hyp, int hyp_len,
                    int32_t* counts, int n, void* stream);

/* K15: FGM normalisation (utlis/eval.py:215-224): p = r/||r||_F with r_b = eps*g_b/||g_b||_2 per
 * sample of a unit; g [n_units, samples_per_unit, elems_per_sample]. */
int dsc_fgm_normalize(const float* g, float* p, float epsilon, int n_units, int samples_per_unit,
                      int elems_per_sample, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * K17: backward kernels.  The reference differentiates with tf.GradientTape (utlis/eval.py:25-33, 197-213;
 * utlis/trainer.py:17-25, 37-62; utlis/gan_train.py:15-44); each entry point below is the gradient of one forward
 * entry point above, fp32, recomputing softmax weights / LayerNorm statistics instead of storing them.
 * ------------------------------------------------------------------------------------------------------------ */

/* C[M,N] (+)= op(A)[M,K] @ op(B)[K,N], fp32.  trans_a: A is stored [K, M] (lda >= M); trans_b: B is stored [N, K].
 * Dense backward: dx = dz @ W^T (trans_b = 1 on the Keras kernel), dW = x^T @ dz (trans_a = 1).  accumulate != 0 adds
 * into C.  Large-K, small-output products are split over K with atomic accumulation. */
int dsc_gemm(const float* A, int64_t lda, int trans_a, const float* B, int64_t ldb, int trans_b,
             float* C, int64_t ldc, int M, int N, int K, int accumulate, void* stream);

/* C[M,N] (+)= A[M,K] @ B[N,K]^T on the tensor cores (bf16x3, fp32 accumulate; both operands fp32 row-major, contiguous
 * along K, 8-byte aligned with even K and leading dimensions; split-K when M*N is small).  dsc_gemm routes its large
 * trans_a = 0, trans_b = 1 products here (dX = dY @ W^T); the tape calls it directly for dW = X^T @ dY after
 * dsc_transpose of both operands.  accumulate != 0: C += (atomicAdd). */
int dsc_gemm_nt_tc(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                   int M, int N, int K, int accumulate, void* stream);

/* dst[c*ld_dst + r] = src[r*ld_src + c] for r < rows, c < cols (fp32, tiled through shared memory). */
int dsc_transpose(const float* src, int64_t ld_src, float* dst, int64_t ld_dst, int rows, int cols, void* stream);

/* dz = dy * (y > 0) for act = 1 (relu; dz may alias dy), and dbias[c] = sum_r dz[r][c] (dbias NULL: skipped;
 * act = 0: dz is not written, only dbias is produced).  y is the forward output of dsc_linear. */
int dsc_bias_act_backward(const float* dy, int64_t ld_dy, const float* y, int64_t ld_y, int act,
                          float* dz, int64_t ld_dz, float* dbias, int M, int N, void* stream);

/* Gradient of dsc_add_layernorm (same row addressing).  dv = d(x + res) (the gradient of both x and res);
 * dgamma_x / dbeta_x [128] are ACCUMULATED (caller zero-fills); the _b pair is NULL for the single-LN form. */
int dsc_add_layernorm_backward(const float* x, int64_t x_group_stride, const float* res, int64_t res_group_stride,
                               const float* gamma_a, const float* beta_a, const float* gamma_b, const float* beta_b,
                               const float* dout, int64_t dout_group_stride, float* dv, int64_t dv_group_stride,
                               float* dgamma_a, float* dbeta_a, float* dgamma_b, float* dbeta_b,
                               int n_rows, int group_rows, void* stream);

/* Gradient of dsc_mha_attention: dq [n, lq, 128] (ld_dq, dq_batch_stride), dk / dv [n, lk, 128] (shared strides). */
int dsc_mha_attention_backward(const float* q, int64_t ldq, int64_t q_batch_stride,
                               const float* k, const float* v, int64_t ldkv, int64_t kv_batch_stride,
                               const float* dout, int64_t ldo, int64_t o_batch_stride,
                               const float* mask, int64_t mask_b_stride, int64_t mask_q_stride,
                               const int32_t* key_ids, int64_t key_ids_stride, int causal, int q_off,
                               float* dq, int64_t ld_dq, int64_t dq_batch_stride,
                               float* dk, float* dv, int64_t ld_dkv, int64_t dkv_batch_stride,
                               int n, int lq, int lk, void* stream);

/* Gradient of dsc_star_satellite_attn: dqkv [n_sent*32, 384] and dkv_e [n_sent*32, 256], fully written. */
int dsc_star_satellite_attn_backward(const float* qkv, const float* kv_e, const float* datt,
                                     float* dqkv, float* dkv_e, int n_sent, void* stream);

/* Gradient of dsc_star_relay_attn: dqkv_r [n_sent*32, 384] fully written; dkv2 [n_sent, kv2_rows, 256] fully written
 * (rows >= n2 are zero).  kv2_rows <= 32. */
int dsc_star_relay_attn_backward(const float* qkv_r, const float* kv2, int kv2_rows, int n2, const float* dout,
                                 float* dqkv_r, float* dkv2, int n_sent, void* stream);

/* Gradient of dsc_embed with respect to the table: dtable[ids] += sqrt(128) * dout (atomic; caller zero-fills). */
int dsc_embed_backward(const int32_t* ids, int64_t ids_stride, const float* dout, int64_t ld_dout,
                       float* dtable, int vocab, int n_sent, int len, void* stream);

/* Gradient of dsc_star_pack: dsrc[i] = dtile[i] + dtile[31] / 31. */
int dsc_star_pack_backward(const float* dtile, float* dsrc, int n_sent, void* stream);

/* Gradient of dsc_masked_ce_rows: dlogits[r] = grad_rows[r] * (softmax(logits[r]) - onehot(target[r])) * (target[r] != 0). */
int dsc_masked_ce_backward(const float* logits, int64_t ld, const int32_t* target, const float* grad_rows,
                           float* dlogits, int64_t ld_d, int M, int N, void* stream);

/* out[u] = sum over unit u of a * b (the reduction the power-norm gradient needs). */
int dsc_unit_dot(const float* a, const float* b, float* out, int n_units, int64_t elems_per_unit, void* stream);

/* Gradient of dsc_power_normalize: dx = r*dy - x * r^3 * (factor/n) * dot[u], r = (factor*sumsq[u]/n)^(-1/2),
 * dot = dsc_unit_dot(x, dy). */
int dsc_power_normalize_backward(const float* x, const float* sumsq, const float* dot, float factor,
                                 const float* dy, float* dx, int n_units, int64_t elems_per_unit, void* stream);

/* Gradient of dsc_channel with respect to its (already normalised) symbol input and perturbation input:
 * AWGN dx = dy, dp = p_scale[u] * dy; fading dx = conj(h) * dy (after undoing the detector), dp = 0.  dx or dp may be NULL. */
int dsc_channel_backward(const float* dy, const float* h, const float* n_std, int detector, const float* p_scale,
                         float* dx, float* dp, int n_units, int64_t elems_per_unit, void* stream);

/* tf.keras.layers.Dropout in training mode (models/modules.py:179-183, 220, 246-250, 424-428, 458-466, 505, 546):
 * out = keep ? x / (1 - rate) : 0 with a Philox4x32-10 keep mask keyed by (seed, offset, element); calling it on dy
 * with the same (seed, offset) is the backward pass.  n % 4 == 0.  step_dev (NULL = none): a device-side step counter
 * mixed into the offset, so that a CUDA-graph replay of a training step draws fresh masks. */
int dsc_dropout(const float* x, float* out, float rate, uint64_t seed, uint64_t offset, const int64_t* step_dev,
                int64_t n, void* stream);

/* One tf.keras.optimizers.Adam update on a flat fp32 buffer: g = grad * grad_scale (+ grad2 * grad2_scale when
 * grad2 != NULL, the lambda-mix of utlis/gan_train.py:22); m, v moments; step >= 1 is the optimizer's iteration count.
 * step_dev (NULL = none): device-side count of completed training steps; the iteration count used for the bias
 * correction is then step + *step_dev * steps_per_iter (CUDA-graph replay of a step with steps_per_iter applies). */
int dsc_adam_step(float* param, const float* grad, const float* grad2, float* m, float* v, float lr, float beta1,
                  float beta2, float eps, int step, const int64_t* step_dev, int steps_per_iter,
                  float grad_scale, float grad2_scale, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPSC_B200_H */
