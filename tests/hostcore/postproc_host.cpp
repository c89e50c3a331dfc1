// TEST INFRASTRUCTURE ONLY -- host build of the per-pixel arithmetic the post-processing kernels execute.
//
// dvis_plus_b200/csrc/resize_core.cuh is written as host/device inline code; this file compiles it with plain g++ and
// drives it with the same work decomposition as the kernels in dvis_plus_b200/csrc/postproc.cu (planes x row bands x
// strips, one call per output pixel), so the CPU test-suite can check that arithmetic against the oracle
// (oracle/postprocess_port.py) and the reference's golden vectors without a GPU.  Built by tests/hostcore/hostcore_build.py into
// tests/hostcore/_build/ (git-ignored); never linked into, loaded by or shipped with libdvis_b200.so.
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "resize_core.cuh"

using namespace dvis::rc;

namespace {

template <typename R>
void vis_masks(const R *logits, int64_t q_stride, int64_t t_stride, const int64_t *sel, int n_sel, int frames, const Geom &g,
               uint8_t *out) {
  constexpr int PX = 8, rows_per_warp = 16;   // postproc.cu: kStripPx, rows_per_warp
  for (int plane = 0; plane < n_sel * frames; ++plane) {
    const int n = plane / frames, t = plane % frames;
    const int64_t q = sel ? sel[n] : n;
    Plane<R> pl;
    pl.p = logits + q * q_stride + t * t_stride;
    pl.w = g.w;
    uint8_t *o = out + (int64_t)plane * g.Ho * g.Wo;
    if (g.Ho == g.Hc && g.Wo == g.Wc) {                      // vis_masks_strip_kernel
      for (int oy_begin = 0; oy_begin < g.Ho; oy_begin += rows_per_warp)
        for (int ox0 = 0; ox0 < g.Wo; ox0 += PX) {
          Strip<PX, R> strip;
          strip.init(g, ox0);
          for (int oy = oy_begin; oy < std::min(oy_begin + rows_per_warp, g.Ho); ++oy) {
            const uint32_t bits = strip.row(pl, g, oy);
            for (int i = 0; i < PX; ++i)
              if (ox0 + i < g.Wo) o[(int64_t)oy * g.Wo + ox0 + i] = uint8_t((bits >> i) & 1u);
          }
        }
    } else {                                                  // vis_masks_two_stage_kernel
      constexpr int PX2 = 4, rows2 = 24;                      // postproc.cu: kStrip2Px, rows_per_warp
      for (int oy_begin = 0; oy_begin < g.Ho; oy_begin += rows2)
        for (int ox0 = 0; ox0 < g.Wo; ox0 += PX2) {
          Strip2<PX2, R> strip;
          strip.init(g, ox0);
          for (int oy = oy_begin; oy < std::min(oy_begin + rows2, g.Ho); ++oy) {
            const uint32_t bits = strip.row(pl, g, oy);
            for (int i = 0; i < PX2; ++i)
              if (ox0 + i < g.Wo) o[(int64_t)oy * g.Wo + ox0 + i] = uint8_t((bits >> i) & 1u);
          }
        }
    }
  }
}

template <typename R>
void vps_argmax(const R *logits, int64_t q_stride, int64_t t_stride, const int64_t *keep_idx, const float *keep_score, int n_keep,
                int frames, const Geom &g, int32_t *win, unsigned long long *areas) {
  std::fill(areas, areas + 3 * n_keep, 0ull);
  for (int t = 0; t < frames; ++t)
    for (int oy = 0; oy < g.Ho; ++oy) {
      const Tap t2y = make_tap(oy, g.s2y, g.Hc);
      for (int ox = 0; ox < g.Wo; ++ox) {
        bool solid = false;
        const int best = vps_pixel(logits + t * t_stride, q_stride, keep_idx, keep_score, n_keep, g, t2y,
                                   make_tap(ox, g.s2x, g.Wc), &solid,
                                   [&](int k, float v) { areas[n_keep + k] += v >= 0.5f; });
        areas[best] += 1;
        areas[2 * n_keep + best] += solid;
        win[((int64_t)t * g.Ho + oy) * g.Wo + ox] = solid ? best : ~best;
      }
    }
}

template <typename R>
void vss_argmax(const R *logits, int64_t q_stride, int64_t t_stride, const float *mask_cls, int64_t cls_stride, int Q, int K,
                int frames, const Geom &g, int64_t *out) {
  constexpr int kClasses = 8;                                 // postproc.cu: kVssClasses
  std::vector<float> probs(Q);
  for (int t = 0; t < frames; ++t)
    for (int oy = 0; oy < g.Ho; ++oy) {
      const Tap t2y = make_tap(oy, g.s2y, g.Hc);
      for (int ox = 0; ox < g.Wo; ++ox) {
        const Tap t2x = make_tap(ox, g.s2x, g.Wc);
        for (int q = 0; q < Q; ++q) {
          Plane<R> pl;
          pl.p = logits + q * q_stride + t * t_stride;
          pl.w = g.w;
          probs[q] = two_stage<true>(pl, g, t2y, t2x);
        }
        float best = 0.f;
        int arg = 0;
        for (int c0 = 0; c0 < K; c0 += kClasses) {
          float acc[kClasses] = {0.f};
          for (int q = 0; q < Q; ++q)
            for (int j = 0; j < kClasses; ++j)
              if (c0 + j < K) acc[j] += mask_cls[q * cls_stride + c0 + j] * probs[q];
          for (int j = 0; j < kClasses; ++j)
            if (c0 + j < K && (c0 + j == 0 || acc[j] > best)) { best = acc[j]; arg = c0 + j; }
        }
        out[((int64_t)t * g.Ho + oy) * g.Wo + ox] = arg;
      }
    }
}

}  // namespace

// dtype: 0 = f32, 2 = bf16 (DVIS_F32 / DVIS_BF16); argument order mirrors include/dvis_b200.h
extern "C" int hostcore_vis_masks(const void *logits, int dtype, int64_t q_stride, int64_t t_stride, const int64_t *sel, int n_sel,
                                  int frames, int h, int w, int H1, int W1, int Hc, int Wc, int Ho, int Wo, uint8_t *out) {
  const Geom g = make_geom(h, w, H1, W1, Hc, Wc, Ho, Wo);
  if (dtype == 0) vis_masks(static_cast<const float *>(logits), q_stride, t_stride, sel, n_sel, frames, g, out);
  else if (dtype == 2) vis_masks(static_cast<const uint16_t *>(logits), q_stride, t_stride, sel, n_sel, frames, g, out);
  else return 2;
  return 0;
}

extern "C" int hostcore_vps_argmax(const void *logits, int dtype, int64_t q_stride, int64_t t_stride, const int64_t *keep_idx,
                                   const float *keep_score, int n_keep, int frames, int h, int w, int H1, int W1, int Hc, int Wc,
                                   int Ho, int Wo, int32_t *win, unsigned long long *areas) {
  const Geom g = make_geom(h, w, H1, W1, Hc, Wc, Ho, Wo);
  if (dtype == 0) vps_argmax(static_cast<const float *>(logits), q_stride, t_stride, keep_idx, keep_score, n_keep, frames, g, win, areas);
  else if (dtype == 2) vps_argmax(static_cast<const uint16_t *>(logits), q_stride, t_stride, keep_idx, keep_score, n_keep, frames, g, win, areas);
  else return 2;
  return 0;
}

extern "C" int hostcore_vss_argmax(const void *logits, int dtype, int64_t q_stride, int64_t t_stride, const float *mask_cls,
                                   int64_t cls_stride, int Q, int K, int frames, int h, int w, int H1, int W1, int Hc, int Wc, int Ho,
                                   int Wo, int64_t *out) {
  const Geom g = make_geom(h, w, H1, W1, Hc, Wc, Ho, Wo);
  if (dtype == 0) vss_argmax(static_cast<const float *>(logits), q_stride, t_stride, mask_cls, cls_stride, Q, K, frames, g, out);
  else if (dtype == 2) vss_argmax(static_cast<const uint16_t *>(logits), q_stride, t_stride, mask_cls, cls_stride, Q, K, frames, g, out);
  else return 2;
  return 0;
}
