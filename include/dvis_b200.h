/*
 * dvis_b200.h -- C ABI of the B200-native DVIS++ hot path (libdvis_b200.so, sm_100a).
 *
 * Plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in `_host`.
 * Every entry point launches asynchronously on `stream` (a cudaStream_t passed as void*), never
 * synchronises, never allocates, never mutates its inputs, and returns DVIS_OK or a DVIS_ERR_* code
 * (message via dvis_last_error()).  There is no CPU fallback anywhere behind this interface.
 *
 * Citations: OPS = DVIS_Plus/mask2former/modeling/pixel_decoder/ops, P = DVIS_Plus (reference @ c0eb2495).
 */
#ifndef DVIS_B200_H_
#define DVIS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVIS_B200_ABI_VERSION 1

enum {
  DVIS_OK = 0,
  DVIS_ERR_INVALID = 1,      /* bad argument (null pointer, non-positive size, misalignment) */
  DVIS_ERR_UNSUPPORTED = 2,  /* dtype / shape combination this build has no kernel for */
  DVIS_ERR_CUDA = 3          /* launch or driver error */
};

/* element types */
enum { DVIS_F32 = 0, DVIS_F64 = 1, DVIS_BF16 = 2 };

int dvis_abi_version(void);
/* thread-local, valid until the next failing call on this thread */
const char *dvis_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Multi-scale deformable attention, forward.
 * Replaces MSDA.ms_deform_attn_forward (OPS/src/vision.cpp:19 -> OPS/src/ms_deform_attn.h:25-44 ->
 * OPS/src/cuda/ms_deform_attn_cuda.cu:25-85 -> ms_deformable_im2col_cuda, ms_deform_im2col_cuda.cuh:928-959).
 * The seven ints are the reference launcher's (cuh:936-942).
 *   value            (batch, spatial_size, num_heads, channels)        dtype
 *   spatial_shapes   (num_levels, 2) int64 (H_l, W_l)                  -- on device, like the reference (cu:72)
 *   level_start      (num_levels,)   int64                             -- on device (cu:73)
 *   sampling_loc     (batch, num_query, num_heads, num_levels, num_point, 2)  dtype, (x, y) in [0,1]
 *   attn_weight      (batch, num_query, num_heads, num_levels, num_point)     dtype
 *   out              (batch, num_query, num_heads*channels)            dtype; fully overwritten (no memset needed)
 *   item_order       optional (num_query*num_heads,) int32 permutation of q*num_heads+m giving the order in
 *                    which (query, head) items are walked -- a locality schedule only; may be NULL.
 * dtype: DVIS_F32 or DVIS_F64 (the reference dispatches float/double only, cu:69).
 * Unlike the reference there is no im2col_step chunking: one launch covers the whole batch.
 */
int dvis_msda_forward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start,
                      const void *sampling_loc, const void *attn_weight, int batch, int spatial_size,
                      int num_heads, int channels, int num_levels, int num_query, int num_point, int dtype,
                      const int32_t *item_order, void *out, void *stream);

/* Multi-scale deformable attention, backward.
 * Replaces MSDA.ms_deform_attn_backward (OPS/src/vision.cpp:20 -> ms_deform_attn_cuda.cu:88-158 ->
 * ms_deformable_col2im_cuda, cuh:961-1331).  grad_value / grad_sampling_loc / grad_attn_weight must be
 * zero-filled by the caller (the reference does at::zeros_like, cu:126-128). */
int dvis_msda_backward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start,
                       const void *sampling_loc, const void *attn_weight, const void *grad_out, int batch,
                       int spatial_size, int num_heads, int channels, int num_levels, int num_query,
                       int num_point, int dtype, void *grad_value, void *grad_sampling_loc,
                       void *grad_attn_weight, void *stream);

/* Fused variant used by the module-level drop-in (MSDeformAttn.forward, OPS/modules/ms_deform_attn.py:98-118):
 * takes the raw outputs of the sampling_offsets / attention_weights linears and the reference points and does
 * softmax over L*P (py:104), location = ref + offset / (W_l, H_l) (py:106-109, 2-d reference points) or the
 * box form (py:110-112, 4-d), bilinear gather and the weighted reduction in one pass.
 *   offsets   (batch, num_query, num_heads, num_levels, num_point, 2), row stride `offsets_stride` elements
 *   logits    (batch, num_query, num_heads, num_levels*num_point),    row stride `logits_stride` elements
 *             both of `param_dtype` (DVIS_F32 or DVIS_BF16); the strides let both live in one fused linear
 *             output of width M*L*P*3
 *   ref       (batch, num_query, num_levels, ref_dim) f32, ref_dim = 2 or 4
 *   value     f32 or bf16 (value_dtype); out f32 or bf16 (out_dtype)
 */
int dvis_msda_fused_forward(const void *value, int value_dtype, const int64_t *spatial_shapes,
                            const int64_t *level_start, const void *offsets, int64_t offsets_stride,
                            const void *logits, int64_t logits_stride, int param_dtype, const float *ref, int ref_dim,
                            int batch, int spatial_size, int num_heads, int channels, int num_levels,
                            int num_query, int num_point, const int32_t *item_order, void *out, int out_dtype,
                            void *stream);

/* The fused forward on a HEAD-MAJOR value tensor (batch, num_heads, spatial_size, 32) bf16 -- the layout the value projection of
 * dvis_linear_tc writes -- with bf16 output: the two x-adjacent corners of a sampling point are 128 contiguous bytes, so a point
 * costs ~3 L1 lines instead of 4 (OPS/src/cuda/ms_deform_im2col_cuda.cuh:242-304 reads (batch, spatial, heads, channels)).
 * channels must be 32; value_hm 16-byte aligned (128-byte for the one-line pairs); everything else as dvis_msda_fused_forward.
 */
int dvis_msda_fused_forward_hm(const void *value_hm, const int64_t *spatial_shapes, const int64_t *level_start,
                               const void *offsets, int64_t offsets_stride, const void *logits, int64_t logits_stride,
                               int param_dtype, const float *ref, int ref_dim, int batch, int spatial_size, int num_heads,
                               int channels, int num_levels, int num_query, int num_point, const int32_t *item_order,
                               void *out, void *stream);

/* Pair-packed bf16 variant of the fused forward (same arithmetic as dvis_msda_fused_forward with bf16 value / output).
 * dvis_msda_pack_pairs re-lays `value` (batch, S, M, 32) bf16 out as pairs (batch, S+1, M, 2, 32) bf16 with
 * pairs[n, e, m] = [value[n, e-1, m], value[n, e, m]] (zeros outside [0, S)), so that the two x-adjacent bilinear
 * corners share one 128-byte line; dvis_msda_pair_forward then gathers 2 lines per sampling point instead of 4.
 * channels must be 32; num_levels * num_point <= 32; out (batch, num_query, M*32) bf16.
 */
int dvis_msda_pack_pairs(const void *value, int batch, int spatial_size, int num_heads, int channels, void *pairs,
                         void *stream);
int dvis_msda_pair_forward(const void *pairs, const int64_t *spatial_shapes, const int64_t *level_start,
                           const void *offsets, int64_t offsets_stride, const void *logits, int64_t logits_stride,
                           int param_dtype, const float *ref, int ref_dim, int batch, int spatial_size, int num_heads,
                           int channels, int num_levels, int num_query, int num_point, const int32_t *item_order,
                           void *out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Mask logits:  out[b, q, p] = sum_c emb[b, q, c] * feat[b, p, c]      (tcgen05 / TMEM GEMM, TMA-fed)
 * Replaces torch.einsum("bqc,bchw->bqhw") of the mask head (P/dvis_Plus/video_mask2former_transformer_decoder.py:363)
 * and "lbtqc,btchw->lbqthw" of tracker / refiner (P/dvis_Plus/tracker.py:379, refiner.py:185-189).
 *   emb   (B, Q, C)    bf16, row-major
 *   feat  (B, HW, C)   bf16, channels-last pixel features (an NCHW tensor in torch.channels_last memory format)
 *   out   (B, Q, HW)   out_dtype (DVIS_F32 or DVIS_BF16), row stride HW; fully overwritten
 * Constraints: C % 64 == 0, C <= 512, Q <= 256 (split larger query sets), 16-byte aligned emb / feat.
 */
int dvis_mask_logits(const void *emb, const void *feat, int B, int Q, int C, int64_t HW, void *out,
                     int out_dtype, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-head attention core on the warp tensor cores (mma.sync bf16, fp32 accumulate, online softmax):
 *   out[b, i, h, :] = softmax_j(scale * q[b,i,h,:] . k[b,j,h,:]) @ v[b,j,h,:]
 * what nn.MultiheadAttention does between its projections; the kernel the temporal stage runs (tracker / refiner self- and cross-attention, P/dvis_Plus/tracker.py:8-92,
 * P/dvis_Plus/refiner.py:105-137) and, with `mask_bits`, the segmenter decoder's masked cross-attention
 * (P/dvis_Plus/video_mask2former_transformer_decoder.py:295-315: `memory_mask=attn_mask`).
 * q (B, Lq, H, Dh), k / v (B, Lk, H, Dh), out (B, Lq, H*Dh) bf16; *_row / *_batch / *_head = elements between consecutive
 * sequence positions / batch items / heads.  k / v rows 16-byte aligned, q / out 4-byte aligned.  Dh in {32, 64}; any Lk.
 * mask_bits (optional, NULL = none): B x Lq rows of bits, bit (j % 8) of byte (j / 8) set = key j may NOT be attended by
 * that query (all heads); rows are mask_row_bytes apart (multiple of 8, >= ceil(Lk / 64) * 8), batches mask_batch_bytes.
 * A row with every key masked yields zeros (torch would give NaN); the reference never produces one (:370-371).
 */
int dvis_flash_attn(const void *q, int64_t q_row, int64_t q_batch, int64_t q_head, const void *k, int64_t k_row,
                    int64_t k_batch, int64_t k_head, const void *v, int64_t v_row, int64_t v_batch, int64_t v_head,
                    void *out, int64_t o_row, int64_t o_batch, const void *mask_bits, int64_t mask_row_bytes,
                    int64_t mask_batch_bytes, int B, int Lq, int Lk, int H, int Dh, float scale, void *stream);

/* ------------------------------------------------------------------------------------------------
 * One dependent step of a tracker / refiner block as one kernel (warp tensor cores, bf16 operands, fp32 accumulate):
 *     Y = act(A @ W^T + bias) [+ residual]
 * W (N, K) bf16 row-major = nn.Linear.weight (nn.MultiheadAttention.in_proj_weight / out_proj.weight, FFNLayer.linear1/2,
 * MLP.layers[i], P/mask2former_video/modeling/transformer_decoder/video_mask2former_transformer_decoder.py:18-205;
 * Conv1d weights of P/dvis_Plus/refiner.py:44-52 re-laid as (C_out, k*C_in)); `batch` independent problems (x_batch /
 * w_batch / bias_batch / y_batch elements apart).
 * A is either
 *   x != NULL: a bf16 matrix (M, K / taps), rows ldx apart; taps > 1 = Conv1d over time with replicate padding as a GEMM:
 *              row r = t * tap_period + q reads, for tap d, row clamp(t + d - tap_pad, 0, tap_len - 1) * tap_period + q;
 *   src0 != NULL: built in the kernel's prologue from the fp32 residual stream (post-norm blocks, :40-50,98-108,160-164;
 *              P/dvis_Plus/tracker.py:37-50):  A = LN1(LN0(src0) + src1), src0 (M, K) f32, src1 (M, K) f32|bf16, each LN /
 *              src1 optional (NULL); K <= 512, K % 128 == 0, batch == 1.  side0 / side1 (optional, (M, K) f32) receive
 *              LN0(src0) and the final A rows (the next residual).
 * residual (optional): f32 (M, N), rows ldr apart, added after the activation.  Outputs: y_f32 and / or y_bf16 (M, N), rows
 * ldy apart.  K % 64 == 0, N % 8 == 0.
 * splitk_workspace / splitk_counters (optional, NULL = never split): 4*M*N floats of scratch and ceil(M/32)*ceil(N/32) ints
 * that are ZERO before the first call (the kernel leaves them zero): lets long reductions (K >= 1024) over a narrow output run
 * as a 4-way split-K whose partial tiles are added in a fixed order by the last CTA to arrive (deterministic, no atomics on
 * the data).
 * dvis_set_pdl(1) launches these kernels (and dvis_flash_attn) with programmatic dependent launch: weight tiles are
 * prefetched while the previous kernel of the stream drains.
 */
int dvis_linear_small(const void *x, int64_t ldx, int64_t x_batch, int taps, int tap_pad, int tap_period, int tap_len,
                      const float *src0, const float *ln0_gamma, const float *ln0_beta, const void *src1, int src1_dtype,
                      const float *ln1_gamma, const float *ln1_beta, float eps, float *side0, float *side1, const void *w,
                      int64_t w_batch, const float *bias, int64_t bias_batch, const float *residual, int64_t ldr, int relu,
                      float *y_f32, void *y_bf16, int64_t ldy, int64_t y_batch, int batch, int M, int N, int K,
                      float *splitk_workspace, int *splitk_counters, void *stream);
/* The same step with the post-norm block's LayerNorm in the producer's EPILOGUE (SelfAttentionLayer / FFNLayer.forward_post,
 * P/mask2former_video/modeling/transformer_decoder/video_mask2former_transformer_decoder.py:40-50,160-164):
 *     y = act(x @ W^T + bias) + residual;   e1 = LN(y; ln_gamma, ln_beta);   e2 = LN(e1 + src1; ln2_gamma, ln2_beta)  (optional)
 * x bf16 (M, K) rows ldx apart, W (N, K) bf16, residual f32 (M, N) rows ldr apart, src1 (M, N) f32|bf16, M <= 512,
 * N <= 512 with N % 128 == 0.  y_f32: (M, N) staging / pre-norm output.  The last CTA of a 32-row block to finish normalises
 * the block (rowblock_counters: ceil(M/32) ints, zero before the first call and left zero).  Outputs ln_f32 / ln_bf16 and
 * ln2_f32 / ln2_bf16, (M, N) contiguous, each optional.  splitk_* as in dvis_linear_small (K >= 1024). */
int dvis_linear_small_ln(const void *x, int64_t ldx, const void *w, const float *bias, const float *residual, int64_t ldr, int relu,
                         int M, int N, int K, float *y_f32, const float *ln_gamma, const float *ln_beta, float eps, const void *src1,
                         int src1_dtype, const float *ln2_gamma, const float *ln2_beta, float *ln_f32, void *ln_bf16, float *ln2_f32,
                         void *ln2_bf16, float *splitk_workspace, int *splitk_counters, int *rowblock_counters, void *stream);
int dvis_set_pdl(int enabled);
/* debug aid of tests/perf (DVIS_LS_PROF=1): clock64 stamps (8 values) of CTA 0 of the last dvis_linear_small launch */
int dvis_debug_linear_small_stamps(long long *host_out);

/* ------------------------------------------------------------------------------------------------
 * Token-wise linear layers of the MSDeformAttn encoder on tcgen05 / TMEM, TMA-fed (csrc/linear_tc.cu):
 *   y[r, :] = x[r, :] . W^T + bias,   x (rows, K) bf16 with row stride ldx, W (N, K) bf16 (nn.Linear.weight), bias (N,) f32 or NULL
 *   N a multiple of 32, <= 256; K a multiple of 64, <= 512.
 * dvis_linear_tc: y (rows, N) bf16, row stride ldy, optional ReLU -- nn.Linear.forward.
 * dvis_linear_tc_heads: MSDeformAttn.value_proj (P/.../ops/modules/ms_deform_attn.py:98-101) with the output written
 *   head-major, value_hm (batch, N/32, S, 32) bf16 -- what dvis_msda_fused_forward_hm reads -- rows = batch * S tokens,
 *   row_mask (batch * S) bytes or NULL: 1 zeroes the token (input_padding_mask, py:99-100).
 * dvis_linear_tc_add_ln: MSDeformAttn.output_proj followed by the encoder layer's post-norm residual block
 *   (ms_deform_attn.py:118, msdeformattn.py:118-119): LayerNorm(residual + y) * gamma + beta over the N columns;
 *   residual (rows, N) f32; outputs as dvis_add_layernorm: out_f32 (rows, N) f32, out_lp bf16, out_lp_pos bf16 = result +
 *   pos[row % pos_rows] (pos (pos_rows, N) f32); each output optional (NULL).
 */
int dvis_linear_tc(const void *x, int64_t ldx, const void *w, const float *bias, int relu, int rows, int N, int K, void *y,
                   int64_t ldy, void *stream);
int dvis_linear_tc_heads(const void *x, int64_t ldx, const void *w, const float *bias, int batch, int S, int N, int K,
                         const uint8_t *row_mask, void *value_hm, void *stream);
int dvis_linear_tc_add_ln(const void *x, int64_t ldx, const void *w, const float *bias, const float *residual,
                          const float *gamma, const float *beta, float eps, int rows, int N, int K, const float *pos,
                          int pos_rows, float *out_f32, void *out_lp, void *out_lp_pos, void *stream);

/* ------------------------------------------------------------------------------------------------
 * y = LayerNorm(x + residual) * gamma + beta over the last dim C, one pass.
 * The post-norm residual blocks of the path: MSDeformAttnTransformerEncoderLayer.forward
 * (P/mask2former/modeling/pixel_decoder/msdeformattn.py:118-119,125-126) and SelfAttentionLayer / CrossAttentionLayer /
 * FFNLayer.forward_post (P/mask2former_video/modeling/transformer_decoder/video_mask2former_transformer_decoder.py:48-49,
 * 109-110,165-166; P/dvis_Plus/tracker.py:50-51).
 *   x (rows, C) x_dtype; residual (rows, C) residual_dtype or NULL; gamma, beta (C,) f32
 *   outputs, each optional (NULL to skip): out_f32 (rows, C) f32; out_lp (rows, C) lp_dtype -- the copy the next
 *   GEMM reads; out_lp_pos (rows, C) lp_dtype = y + pos[row % pos_rows] (with_pos_embed, msdeformattn.py:112-114),
 *   pos (pos_rows, C) f32.
 * dtypes: DVIS_F32 or DVIS_BF16.  C must be a multiple of 128 (built: 128..1024, 2048).
 */
int dvis_add_layernorm(const void *x, int x_dtype, const void *residual, int residual_dtype, const float *gamma,
                       const float *beta, const float *pos, int64_t pos_rows, int64_t rows, int C, float eps,
                       float *out_f32, void *out_lp, void *out_lp_pos, int lp_dtype, void *stream);

/* ------------------------------------------------------------------------------------------------
 * GroupNorm over a channels-last (N, HW, C) map with fused epilogue:
 *   y = GroupNorm_G(x) * gamma + beta  [+ bilinear_upsample(up) to (H, W), align_corners=False]  [ReLU]
 * Replaces nn.GroupNorm(32, conv_dim) after the pixel decoder's 1x1 / 3x3 convs together with the FPN top-down add and
 * the ReLU (P/mask2former/modeling/pixel_decoder/msdeformattn.py:213-226,262-286,321,346-351) and the flatten /
 * transpose / cat / with_pos_embed that build the encoder inputs (:70-80,112-114).
 *   x (N, HW, C) x_dtype, batch stride x_batch_stride elements; gamma, beta (C,) f32
 *   sums_workspace: 2*N*G*(1 + ceil(HW / 256)) doubles of scratch (per-chunk partial sums, reduced in a fixed order)
 *   up: optional (N, up_h, up_w, C) f32 with batch stride up_batch_stride (needs H*W == HW), else NULL
 *   outputs, each optional: out_f32; out_lp (lp_dtype); out_lp_pos = y + pos[pixel] (pos (HW, C) f32); all laid out as
 *   (N, HW, C) with batch stride out_batch_stride elements (so a level can land in its slice of a token buffer).
 * Constraints: C % 4 == 0, (C/G) % 4 == 0, C <= 1024, G <= 64; dtypes DVIS_F32 or DVIS_BF16.
 */
int dvis_groupnorm_nhwc(const void *x, int x_dtype, int64_t x_batch_stride, int N, int HW, int C, int G,
                        const float *gamma, const float *beta, float eps, int relu, double *sums_workspace,
                        const float *up, int64_t up_batch_stride, int up_h, int up_w, int H, int W, const float *pos,
                        float *out_f32, void *out_lp, void *out_lp_pos, int lp_dtype, int64_t out_batch_stride,
                        void *stream);

/* ------------------------------------------------------------------------------------------------
 * Batched linear assignment + index chain for the tracker's frame-to-frame query matching.
 * Replaces, for a whole window at once, the per-frame `C.cpu()` + scipy.optimize.linear_sum_assignment(C.T)[1] of
 * Noiser.match_embds (P/dvis_Plus/noiser.py:43-56; called per frame at P/dvis_Plus/tracker.py:224,285).
 *   cost   (T, n, n) f32: cost[t][r][c] = 1 - cos(reference item r of frame t, current item c of frame t), where the
 *          reference items of frame t are frame t-1's items in their ORIGINAL order (frame 0: the window's reference);
 *          NaN entries count as 0 (noiser.py:52)
 *   sigma  (T, n) i64 out: per-frame optimal assignment, row r -> column sigma[t][r]
 *   idx    (T, n) i64 out: the tracker's indices, idx[t] = sigma[t] o idx[t-1], idx[-1] = idx_init (NULL = identity)
 * n <= 1024.  One CTA per frame; exact (double potentials), equal to SciPy's result whenever the optimum is unique.
 */
int dvis_lap_chain(const float *cost, int T, int n, const int64_t *idx_init, int64_t *sigma, int64_t *idx, void *stream);

/* Rectangular linear assignment, batched: cost (B, rows, cols) f32 row-major -> row_to_col (B, rows) i64, the column
 * assigned to each row, -1 for a row left unmatched (only when rows > cols); min(rows, cols) pairs of minimum total cost,
 * i.e. scipy.optimize.linear_sum_assignment(cost) as a dense map.  Replaces the `C.cpu()` + SciPy call of DVIS-DAQ's
 * VideoInstanceCutter.match_with_embeds (D/dvis_daq/track_module.py:749-759: track queries x segmenter queries).
 * rows, cols <= 1024; NaN entries count as 0. */
int dvis_lap_rect(const float *cost, int B, int rows, int cols, int64_t *row_to_col, void *stream);

/* Strided form for query slices: emb (B, Q, C) with `emb_batch_stride` elements between batch items (a multiple of 8) and
 * out (B, Q, HW) with `out_batch_stride` elements between batch items, so a slice [q0, q1) of a larger query set can be
 * computed in place (used to split Q > 256, e.g. the DAQ stress size Q = 300, into two launches). */
int dvis_mask_logits_strided(const void *emb, int64_t emb_batch_stride, const void *feat, int B, int Q, int C, int64_t HW,
                             void *out, int64_t out_batch_stride, int out_dtype, void *stream);

/* dvis_mask_logits for the T frames of a clip, written query-major: out (Q, T, HW) -- the "b q t h w" layout (b = 1) of
 * P/dvis_Plus/refiner.py:185-189 / tracker.py:379 -- instead of (T, Q, HW); emb (T, Q, C) bf16, feat (T, HW, C) bf16, Q <= 256. */
int dvis_mask_logits_clip(const void *emb, const void *feat, int T, int Q, int C, int64_t HW, void *out, int out_dtype,
                          void *stream);

/* The same two operators with fp32 operands multiplied as TF32 (tcgen05 kind::tf32) and fp32 results: the mask head of the
 * reference when it runs without autocast (decoder.py:363), inside 1e-3 of the output scale.  emb (B, Q, C) f32, feat (B, HW, C)
 * f32 (channels-last), out / bias (B, Q, HW) f32; any Q (query slices of 128 inside); row_open_workspace B*Q ints.
 */
int dvis_mask_logits_tf32(const void *emb, const void *feat, int B, int Q, int C, int64_t HW, void *out, void *stream);
int dvis_mask_attn_bias_tf32(const void *emb, const void *feat, int B, int Q, int C, int64_t HW, void *bias,
                             int *row_open_workspace, void *stream);

/* Same GEMM with the masked-attention decoder's threshold fused into the epilogue
 * (P/dvis_Plus/video_mask2former_transformer_decoder.py:370-371 and :297): writes, instead of the logits, the additive
 * attention bias (B, Q, HW) in bias_dtype (DVIS_F32 | DVIS_BF16): -inf where sigmoid(logit) < 0.5, 0 elsewhere, rows that
 * would be -inf everywhere reset to 0.  row_open_workspace: B*Q ints of scratch.  `feat` is mask_features already
 * resized to the attention level (interpolate(E @ F) == E @ interpolate(F)). */
int dvis_mask_attn_bias(const void *emb, const void *feat, int B, int Q, int C, int64_t HW, void *bias, int bias_dtype,
                        int *row_open_workspace, void *stream);

/* The same threshold as ONE BIT per (query, pixel) -- the form dvis_flash_attn's `mask_bits` consumes: bit (p % 8) of byte
 * (p / 8) of row (b*Q + q) is set where sigmoid(logit) < 0.5, i.e. the query may NOT attend pixel p
 * (P/dvis_Plus/video_mask2former_transformer_decoder.py:370-371); rows that would be fully masked are cleared (:297).
 * bits: B*Q rows of bits_row_bytes bytes (multiple of 8, >= ceil(HW / 64) * 8; bytes past ceil(HW / 8) are unspecified).
 * row_open_workspace: B*Q ints. */
int dvis_mask_attn_bits(const void *emb, const void *feat, int B, int Q, int C, int64_t HW, void *bits, int64_t bits_row_bytes,
                        int *row_open_workspace, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Helpers of the masked-attention decoder's mask head (P/dvis_Plus/video_mask2former_transformer_decoder.py:358-374).
 * dvis_resize_bilinear_nhwc: F.interpolate(mode="bilinear", align_corners=False) (py:367) of a channels-last bf16 map
 *   in (N, h, w, C) -> out (N, H, W, C); C % 4 == 0.  Used once per attention level on mask_features:
 *   interpolate(E @ F) == E @ interpolate(F).
 * dvis_attn_bias_from_logits: logits (rows, hw) f32 -> additive attention bias (rows, hw) f32|bf16: -inf where
 *   sigmoid(logit) < 0.5 (py:370-371), 0 elsewhere; a row that would be -inf everywhere becomes all 0 (py:297).
 */
int dvis_resize_bilinear_nhwc(const void *in, int N, int h, int w, int C, void *out, int H, int W, void *stream);
/* The decoder's per-level memory in one pass (py:270-279: src = input_proj(x) + level_embed; keys see src + pos):
 *   tok[b,p,:] = x[b,p,:] + level_embed[:]   and (key != NULL)   key[b,p,:] = tok[b,p,:] + pos[p,:],  both bf16 (B, HW, C).
 * x: f32|bf16, pixel stride C, batch stride x_batch_stride elements (a slice of the encoder's token buffer); pos (HW, C) f32. */
int dvis_level_tokens(const void *x, int x_dtype, int64_t x_batch_stride, const float *level_embed, const float *pos, int B, int HW,
                      int C, void *tok, void *key, void *stream);
int dvis_attn_bias_from_logits(const float *logits, int64_t rows, int hw, void *bias, int bias_dtype, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Video post-processing: what DVIS_Plus_online.inference_video_vis / _vps / _vss do after the final mask GEMM
 * (P/dvis_Plus/meta_architecture.py:818-868, 870-956, 958-979; same code in D/dvis_daq/meta_architecture.py:704-).
 * The resize chain  logits (h, w) --bilinear--> (H1, W1) = padded network input --crop--> (Hc, Wc) = image without
 * padding --[sigmoid]--> --bilinear--> (Ho, Wo)  (F.interpolate(mode="bilinear", align_corners=False) twice,
 * py:838-844) is evaluated per OUTPUT pixel from the stride-4 logits; no full-resolution float tensor is ever written.
 * `logits` is a (queries, frames, h, w) view of f32 or bf16 mask logits with contiguous (h, w) planes: element
 * strides `q_stride` between queries and `t_stride` between frames (the frame-major (T, Q, HW) output of
 * dvis_mask_logits has q_stride = HW, t_stride = Q*HW).
 */

/* scores[q, c] = softmax(pred_cls[q, :])[c]; for c < K1-1 max'ed with softmax(aux_cls[q, :])[c] when aux_cls != NULL
 * (py:823-827, 873-876, 962-965).  pred_cls, aux_cls, scores: (Q, K1) f32, K1 = classes + 1 (no-object last). */
int dvis_class_scores(const float *pred_cls, const float *aux_cls, int Q, int K1, float *scores, void *stream);

/* VIS instance selection (py:823-835): class scores as above, then the max_num largest of the Q*(K1-1) object scores.
 * Output r (score descending; ties: lower flat index first -- torch.topk(sorted=False) leaves the order open):
 * out_scores[r], out_labels[r] = flat % (K1-1), out_query[r] = flat / (K1-1).  scores_workspace: Q*K1 floats
 * (receives the scores, selected entries overwritten with -inf).  Fails like torch.topk if max_num > Q*(K1-1). */
int dvis_vis_topk(const float *pred_cls, const float *aux_cls, int Q, int K1, int max_num, float *scores_workspace,
                  float *out_scores, int64_t *out_labels, int64_t *out_query, void *stream);

/* VIS masks (py:836-846): out[n, t] = resize_chain(logits[sel[n], t]) > 0 as bytes 0/1 (torch.bool layout),
 * out (n_sel, frames, Ho, Wo).  sel: n_sel int64 query indices on the device, NULL = queries 0..n_sel-1.
 * n_sel*frames <= 65535 and Ho <= 65535 per call. */
int dvis_vis_masks(const void *logits, int logits_dtype, int64_t q_stride, int64_t t_stride, const int64_t *sel, int n_sel,
                   int frames, int h, int w, int H1, int W1, int Hc, int Wc, int Ho, int Wo, uint8_t *out, void *stream);

/* Same masks as ONE BIT per pixel: out (n_sel, frames, Ho, ceil(Wo/8)) bytes, bit i of byte b = pixel 8*b + i of the row
 * (numpy.unpackbits(..., bitorder="little")), bits past Wo are 0.  8x less HBM write and device->host traffic for results
 * that are consumed on the host (the reference moves its bool masks to the CPU, py:851). */
int dvis_vis_masks_packed(const void *logits, int logits_dtype, int64_t q_stride, int64_t t_stride, const int64_t *sel,
                          int n_sel, int frames, int h, int w, int H1, int W1, int Hc, int Wc, int Ho, int Wo, uint8_t *out,
                          void *stream);

/* VPS (py:889-926): per output pixel the arg-max over the kept queries k of keep_score[k] * sigmoid-resized mask
 * (cur_prob_masks.argmax(0), first maximum), plus the three pixel counts the segment filter needs, so the host loop
 * (py:919-949) runs on 3*n_keep integers instead of n_keep full-resolution reductions with a sync each.
 *   keep_idx (n_keep) int64 query indices, keep_score (n_keep) f32 -- on the device
 *   win (frames, Ho, Wo) int32 out: k if the winner's own probability is >= 0.5 (the pixel belongs to `mask`, py:925),
 *       ~k (negative) if it is below
 *   areas (3, n_keep) uint64 out: [0] pixels won by k (py:923), [1] pixels with probability_k >= 0.5 (py:924),
 *       [2] pixels won by k with probability_k >= 0.5 (py:925-926); zeroed by the call
 * dvis_vps_paint then writes panoptic[i] = win[i] >= 0 ? seg_of_k[win[i]] : 0 (py:934,939; seg_of_k[k] = 0 drops k). */
int dvis_vps_argmax(const void *logits, int logits_dtype, int64_t q_stride, int64_t t_stride, const int64_t *keep_idx,
                    const float *keep_score, int n_keep, int frames, int h, int w, int H1, int W1, int Hc, int Wc, int Ho,
                    int Wo, int32_t *win, unsigned long long *areas, void *stream);
int dvis_vps_paint(const int32_t *win, const int32_t *seg_of_k, int64_t total, int32_t *panoptic, void *stream);

/* VSS (py:968-975): out[t, y, x] = argmax_c sum_q mask_cls[q, c] * sigmoid-resized mask_q (einsum "qc,qthw->cthw" +
 * max(0), first maximum).  mask_cls (Q, K) f32 with row stride cls_stride (>= K; the scores of dvis_class_scores
 * without their last column: cls_stride = K + 1).  out (frames, Ho, Wo) int64.  Q <= 400. */
int dvis_vss_argmax(const void *logits, int logits_dtype, int64_t q_stride, int64_t t_stride, const float *mask_cls,
                    int64_t cls_stride, int Q, int K, int frames, int h, int w, int H1, int W1, int Hc, int Wc, int Ho,
                    int Wo, int64_t *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DVIS_B200_H_ */
