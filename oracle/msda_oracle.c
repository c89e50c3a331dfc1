/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * Plain-C, single-threaded CPU restatement of the reference's multi-scale deformable attention
 * forward / backward arithmetic and of the mask-logit contraction.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; dvis_plus_b200/ never does.
 *
 * Parity pin: checked against outputs of the reference's own `ms_deform_attn_core_pytorch`
 * (OPS/functions/ms_deform_attn_func.py:52-72) on the OPS/test.py:24-39 inputs and larger seeded
 * cases -- fixtures under tests/golden/ made by tests/golden/make_golden.py (tests/test_oracle.py).
 *
 * OPS = /root/reference/DVIS_Plus/mask2former/modeling/pixel_decoder/ops
 *
 * What it follows:
 *   forward  : OPS/src/cuda/ms_deform_im2col_cuda.cuh:242-304 (per-output-element loop over L x P points,
 *              pixel coordinate = loc * size - 0.5, point kept iff -1 < h < H and -1 < w < W)
 *              and :38-89 (4-corner bilinear sample, zero outside the map, value layout (S, M, D)).
 *   backward : OPS/src/cuda/ms_deform_im2col_cuda.cuh:92-164 (corner scatter into grad_value, gradients of
 *              the sampling location scaled by W / H, gradient of the attention weight = top_grad * sample),
 *              reduced over channels as the shm_reduce kernels do (:306-408).
 *   mask     : einsum "bqc,bchw->bqhw", DVIS_Plus/dvis_Plus/video_mask2former_transformer_decoder.py:363.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/build.py).  -ffp-contract=off keeps the
 * multiply/add sequence un-fused like the reference's scalar code path.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define DEFINE_MSDA(NAME, T)                                                                           \
  /* one bilinear tap set; mirrors cuh:38-89 */                                                        \
  static T NAME##_sample(const T *lvl, int H, int W, int M, int D, T h, T w, int m, int c) {           \
    const int h0 = (int)floor((double)h), w0 = (int)floor((double)w);                                  \
    const int h1 = h0 + 1, w1 = w0 + 1;                                                                \
    const T lh = h - (T)h0, lw = w - (T)w0, hh = (T)1 - lh, hw = (T)1 - lw;                            \
    const long row = (long)M * D, base = (long)m * D + c;                                              \
    T v00 = 0, v01 = 0, v10 = 0, v11 = 0;                                                              \
    if (h0 >= 0 && w0 >= 0) v00 = lvl[((long)h0 * W + w0) * row + base];                               \
    if (h0 >= 0 && w1 <= W - 1) v01 = lvl[((long)h0 * W + w1) * row + base];                           \
    if (h1 <= H - 1 && w0 >= 0) v10 = lvl[((long)h1 * W + w0) * row + base];                           \
    if (h1 <= H - 1 && w1 <= W - 1) v11 = lvl[((long)h1 * W + w1) * row + base];                       \
    return hh * hw * v00 + hh * lw * v01 + lh * hw * v10 + lh * lw * v11;                              \
  }                                                                                                    \
  /* forward; mirrors cuh:242-304.  value (N,S,M,D) loc (N,Lq,M,L,P,2)=(x,y) attn (N,Lq,M,L,P)     */  \
  void NAME##_forward(const T *value, const int64_t *shapes, const int64_t *lsi, const T *loc,         \
                      const T *attn, int N, int S, int M, int D, int L, int Lq, int P, T *out) {       \
    for (int n = 0; n < N; ++n)                                                                        \
      for (int q = 0; q < Lq; ++q)                                                                     \
        for (int m = 0; m < M; ++m) {                                                                  \
          const long qm = ((long)n * Lq + q) * M + m;                                                  \
          for (int c = 0; c < D; ++c) {                                                                \
            T acc = 0;                                                                                 \
            for (int l = 0; l < L; ++l) {                                                              \
              const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                            \
              const T *lvl = value + ((long)n * S + lsi[l]) * M * D;                                   \
              for (int p = 0; p < P; ++p) {                                                            \
                const long i = (qm * L + l) * P + p;                                                   \
                const T x = loc[2 * i], y = loc[2 * i + 1], a = attn[i];                               \
                const T h = (T)((double)(y * (T)H) - 0.5), w = (T)((double)(x * (T)W) - 0.5);          \
                if (h > -1 && w > -1 && h < H && w < W)                                                \
                  acc += NAME##_sample(lvl, H, W, M, D, h, w, m, c) * a;                               \
              }                                                                                        \
            }                                                                                          \
            out[qm * D + c] = acc;                                                                     \
          }                                                                                            \
        }                                                                                              \
  }                                                                                                    \
  /* backward; mirrors cuh:92-164 with the per-(q,m,l,p) channel reduction of cuh:306-408.          */ \
  /* grad_value / grad_loc / grad_attn must be zero-filled by the caller (ms_deform_attn_cuda.cu:126-128) */ \
  void NAME##_backward(const T *value, const int64_t *shapes, const int64_t *lsi, const T *loc,        \
                       const T *attn, const T *gout, int N, int S, int M, int D, int L, int Lq, int P, \
                       T *gvalue, T *gloc, T *gattn) {                                                 \
    for (int n = 0; n < N; ++n)                                                                        \
      for (int q = 0; q < Lq; ++q)                                                                     \
        for (int m = 0; m < M; ++m) {                                                                  \
          const long qm = ((long)n * Lq + q) * M + m;                                                  \
          for (int l = 0; l < L; ++l) {                                                                \
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                              \
            const long lvl_off = ((long)n * S + lsi[l]) * M * D;                                       \
            for (int p = 0; p < P; ++p) {                                                              \
              const long i = (qm * L + l) * P + p;                                                     \
              const T x = loc[2 * i], y = loc[2 * i + 1], a = attn[i];                                 \
              const T h = (T)((double)(y * (T)H) - 0.5), w = (T)((double)(x * (T)W) - 0.5);            \
              if (!(h > -1 && w > -1 && h < H && w < W)) continue;                                     \
              const int h0 = (int)floor((double)h), w0 = (int)floor((double)w);                        \
              const int h1 = h0 + 1, w1 = w0 + 1;                                                      \
              const T lh = h - (T)h0, lw = w - (T)w0, hh = (T)1 - lh, hw = (T)1 - lw;                  \
              const long row = (long)M * D;                                                            \
              T gx = 0, gy = 0, ga = 0;                                                                \
              for (int c = 0; c < D; ++c) {                                                            \
                const long base = (long)m * D + c;                                                     \
                const T tg = gout[qm * D + c], tgv = tg * a;                                           \
                T gh = 0, gw = 0, v00 = 0, v01 = 0, v10 = 0, v11 = 0;                                  \
                if (h0 >= 0 && w0 >= 0) {                                                              \
                  const long o = lvl_off + ((long)h0 * W + w0) * row + base;                           \
                  v00 = value[o]; gh -= hw * v00; gw -= hh * v00; gvalue[o] += hh * hw * tgv;          \
                }                                                                                      \
                if (h0 >= 0 && w1 <= W - 1) {                                                          \
                  const long o = lvl_off + ((long)h0 * W + w1) * row + base;                           \
                  v01 = value[o]; gh -= lw * v01; gw += hh * v01; gvalue[o] += hh * lw * tgv;          \
                }                                                                                      \
                if (h1 <= H - 1 && w0 >= 0) {                                                          \
                  const long o = lvl_off + ((long)h1 * W + w0) * row + base;                           \
                  v10 = value[o]; gh += hw * v10; gw -= lh * v10; gvalue[o] += lh * hw * tgv;          \
                }                                                                                      \
                if (h1 <= H - 1 && w1 <= W - 1) {                                                      \
                  const long o = lvl_off + ((long)h1 * W + w1) * row + base;                           \
                  v11 = value[o]; gh += lw * v11; gw += lh * v11; gvalue[o] += lh * lw * tgv;          \
                }                                                                                      \
                ga += tg * (hh * hw * v00 + hh * lw * v01 + lh * hw * v10 + lh * lw * v11);            \
                gx += (T)W * gw * tgv;                                                                 \
                gy += (T)H * gh * tgv;                                                                 \
              }                                                                                        \
              gattn[i] = ga; gloc[2 * i] = gx; gloc[2 * i + 1] = gy;                                   \
            }                                                                                          \
          }                                                                                            \
        }                                                                                              \
  }

DEFINE_MSDA(msda_oracle_f32, float)
DEFINE_MSDA(msda_oracle_f64, double)

/* mask logits: out[b,q,hw] = sum_c emb[b,q,c] * feat[b,c,hw]   (decoder.py:363, NCHW features).
 * Accumulates in double so that it is a tight checker for both the fp32 and the bf16 GPU paths. */
void mask_oracle_f32(const float *emb, const float *feat, int B, int Q, int C, long HW, float *out) {
  for (int b = 0; b < B; ++b)
    for (int q = 0; q < Q; ++q) {
      const float *e = emb + ((long)b * Q + q) * C;
      float *o = out + ((long)b * Q + q) * HW;
      for (long p = 0; p < HW; ++p) {
        double acc = 0.0;
        const float *f = feat + (long)b * C * HW + p;
        for (int c = 0; c < C; ++c) acc += (double)e[c] * (double)f[(long)c * HW];
        o[p] = (float)acc;
      }
    }
}
