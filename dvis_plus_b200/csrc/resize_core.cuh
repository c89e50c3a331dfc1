// Per-pixel arithmetic of the video post-processing kernels (postproc.cu), written once as host/device inline code.
//
// Under nvcc every function here is __host__ __device__ and is what the kernels execute on the GPU.  The same header
// compiles with plain g++ into the test harness tests/hostcore/postproc_host.cpp, which lets the CPU test-suite check
// this arithmetic against the oracle without a GPU.  That harness is test infrastructure: libdvis_b200.so contains no
// host implementation of any kernel.
//
// The arithmetic follows what the reference calls: F.interpolate(mode="bilinear", align_corners=False) twice with a crop
// (and, for vps / vss, a sigmoid) in between -- P/dvis_Plus/meta_architecture.py:838-846, 889-895, 968-974 -- i.e.
// at::native::upsample_bilinear2d: scale = in / out (float), src = max(scale * (dst + 0.5) - 0.5, 0), i0 = int(src),
// i1 = i0 + (i0 < in - 1), lambda1 = src - i0, value = h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DVIS_HD __host__ __device__ __forceinline__
#define DVIS_UNROLL _Pragma("unroll")
#else
#include <cmath>
#define DVIS_HD inline
#define DVIS_UNROLL
#endif

namespace dvis {
namespace rc {

// out = w0 * in[i0] + w1 * in[i1]
struct Tap {
  int i0, i1;
  float w0, w1;
};

DVIS_HD Tap make_tap(int dst, float scale, int in_size) {
  float src = scale * (float(dst) + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  int i0 = int(src);
  if (i0 > in_size - 1) i0 = in_size - 1;
  const int i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  float l1 = src - float(i0);
  l1 = l1 < 0.f ? 0.f : (l1 > 1.f ? 1.f : l1);
  Tap t;
  t.i0 = i0;
  t.i1 = i1;
  t.w0 = 1.f - l1;
  t.w1 = l1;
  return t;
}

// Geometry of the chain (h, w) --resize--> (H1, W1) --crop--> (Hc, Wc) --resize--> (Ho, Wo); the uncropped size only
// enters through the first scale.
struct Geom {
  int h, w;        // mask logits as produced by the mask GEMM (stride-4 grid)
  int Hc, Wc;      // image size without padding (crop of the first resize)
  int Ho, Wo;      // output size
  float s1y, s1x;  // h / H1, w / W1
  float s2y, s2x;  // Hc / Ho, Wc / Wo
};

inline Geom make_geom(int h, int w, int H1, int W1, int Hc, int Wc, int Ho, int Wo) {
  Geom g;
  g.h = h; g.w = w; g.Hc = Hc; g.Wc = Wc; g.Ho = Ho; g.Wo = Wo;
  g.s1y = float(h) / float(H1); g.s1x = float(w) / float(W1);
  g.s2y = float(Hc) / float(Ho); g.s2x = float(Wc) / float(Wo);
  return g;
}

// element loads: fp32, or bf16 passed as its 16 raw bits
DVIS_HD float ld_elem(const float *p) { return *p; }
DVIS_HD float ld_elem(const uint16_t *p) {
  union { uint32_t u; float f; } c;
  c.u = uint32_t(*p) << 16;
  return c.f;
}

DVIS_HD float sigmoidf_(float x) {
#if defined(__CUDA_ARCH__)
  return 1.f / (1.f + expf(-x));   // full-precision expf: arg-max ties are decided on these values
#else
  return 1.f / (1.f + std::exp(-x));
#endif
}

// one plane (h, w) of mask logits, rows contiguous
template <typename T>
struct Plane {
  const T *p;
  int w;
  DVIS_HD float at(int y, int x) const { return ld_elem(p + (int64_t)y * w + x); }
};

// a rectangular window [y0, y0 + rows) x [x0, x0 + w) of a plane, staged as f32 (shared memory on the device)
struct Window {
  const float *p;
  int w, y0, x0;
  DVIS_HD float at(int y, int x) const { return p[(y - y0) * w + (x - x0)]; }
};

template <typename T>
DVIS_HD float bilinear(const Plane<T> &pl, const Tap &ty, const Tap &tx) {
  return ty.w0 * (tx.w0 * pl.at(ty.i0, tx.i0) + tx.w1 * pl.at(ty.i0, tx.i1)) +
         ty.w1 * (tx.w0 * pl.at(ty.i1, tx.i0) + tx.w1 * pl.at(ty.i1, tx.i1));
}

// value of the first resize at pixel (iy, ix) of the (cropped) intermediate image
template <bool kSigmoid, typename T>
DVIS_HD float stage1(const Plane<T> &pl, const Geom &g, int iy, int ix) {
  const float v = bilinear(pl, make_tap(iy, g.s1y, g.h), make_tap(ix, g.s1x, g.w));
  return kSigmoid ? sigmoidf_(v) : v;
}

// value of the whole chain at output pixel taps (t2y, t2x) = make_tap(oy, s2y, Hc), make_tap(ox, s2x, Wc).
// Taps with weight exactly 0 are not evaluated (identical for finite inputs; makes an identity second resize cost one
// first-stage sample instead of four).
template <bool kSigmoid, typename T>
DVIS_HD float two_stage(const Plane<T> &pl, const Geom &g, const Tap &t2y, const Tap &t2x) {
  const bool x1 = t2x.w1 != 0.f, y1 = t2y.w1 != 0.f;
  float top = t2x.w0 * stage1<kSigmoid>(pl, g, t2y.i0, t2x.i0);
  if (x1) top += t2x.w1 * stage1<kSigmoid>(pl, g, t2y.i0, t2x.i1);
  float v = t2y.w0 * top;
  if (y1) {
    float bot = t2x.w0 * stage1<kSigmoid>(pl, g, t2y.i1, t2x.i0);
    if (x1) bot += t2x.w1 * stage1<kSigmoid>(pl, g, t2y.i1, t2x.i1);
    v += t2y.w1 * bot;
  }
  return v;
}

// ---- single-resize strip walker (second resize is the identity: Ho == Hc, Wo == Wc) -------------------------------
// A strip is PX consecutive output pixels of one row.  Walking down the rows of a band, the two source rows change only
// every ~1/s1y output rows, so the horizontally blended source rows are kept and only re-read when they change.
template <int PX, typename T>
struct Strip {
  Tap tx[PX];
  float top[PX], bot[PX];
  int y0, y1;

  DVIS_HD void init(const Geom &g, int ox0) {
    DVIS_UNROLL
    for (int i = 0; i < PX; ++i) {
      const int ox = ox0 + i < g.Wo ? ox0 + i : g.Wo - 1;   // columns past the edge are computed but never stored
      tx[i] = make_tap(ox, g.s1x, g.w);
    }
    y0 = -1;
    y1 = -1;
  }
  // -> bit i of the result = (resized logit of pixel ox0 + i on row oy) > 0
  template <typename PL>
  DVIS_HD uint32_t row(const PL &pl, const Geom &g, int oy) {
    const Tap ty = make_tap(oy, g.s1y, g.h);
    if (ty.i0 != y0 || ty.i1 != y1) {
      y0 = ty.i0;
      y1 = ty.i1;
      DVIS_UNROLL
      for (int i = 0; i < PX; ++i) {
        top[i] = tx[i].w0 * pl.at(y0, tx[i].i0) + tx[i].w1 * pl.at(y0, tx[i].i1);
        bot[i] = tx[i].w0 * pl.at(y1, tx[i].i0) + tx[i].w1 * pl.at(y1, tx[i].i1);
      }
    }
    uint32_t bits = 0;
    DVIS_UNROLL
    for (int i = 0; i < PX; ++i) bits |= uint32_t(ty.w0 * top[i] + ty.w1 * bot[i] > 0.f) << i;
    return bits;
  }
};

// ---- two-resize strip walker (general chain, no sigmoid) -----------------------------------------------------------
// Same walk as Strip, for a second resize that changes the size.  Per thread: PX output pixels need 2*PX columns of the
// intermediate (cropped, first-resize) image -- ix0 / ix1 of every pixel -- in two intermediate rows (iy0, iy1), and every
// intermediate row needs two source rows.  Both levels are cached and shifted as the walk moves down: a source row is
// read (2 loads per intermediate column) only when it first enters, an intermediate row is blended from the two cached
// source rows only when it first enters.  All cache decisions depend on row indices only, i.e. are warp-uniform.
// The arithmetic per value is exactly the two nested bilinear formulas of the reference (see the file header).
template <int PX, typename T>
struct Strip2 {
  float w2x0[PX], w2x1[PX];     // second-resize x weights of the PX output pixels
  Tap ctap[2 * PX];             // first-resize x taps of intermediate columns (ix0_0, ix1_0, ix0_1, ix1_1, ...)
  float lowA[2 * PX], lowB[2 * PX];   // horizontally blended source rows ya / yb at the intermediate columns
  float midA[2 * PX], midB[2 * PX];   // intermediate rows ma / mb at the intermediate columns
  int ya, yb, ma, mb;

  DVIS_HD void init(const Geom &g, int ox0) {
    DVIS_UNROLL
    for (int i = 0; i < PX; ++i) {
      const int ox = ox0 + i < g.Wo ? ox0 + i : g.Wo - 1;   // columns past the edge are computed but never stored
      const Tap t2 = make_tap(ox, g.s2x, g.Wc);
      w2x0[i] = t2.w0;
      w2x1[i] = t2.w1;
      ctap[2 * i] = make_tap(t2.i0, g.s1x, g.w);
      ctap[2 * i + 1] = make_tap(t2.i1, g.s1x, g.w);
    }
    ya = yb = ma = mb = -1;
  }
  template <typename PL>
  DVIS_HD void load_low(const PL &pl, int y, float *dst) {
    DVIS_UNROLL
    for (int j = 0; j < 2 * PX; ++j) dst[j] = ctap[j].w0 * pl.at(y, ctap[j].i0) + ctap[j].w1 * pl.at(y, ctap[j].i1);
  }
  DVIS_HD void copy(float *dst, const float *src) {
    DVIS_UNROLL
    for (int j = 0; j < 2 * PX; ++j) dst[j] = src[j];
  }
  // make (lowA, lowB) hold source rows (y0, y1)
  template <typename PL>
  DVIS_HD void ensure_low(const PL &pl, int y0, int y1) {
    if (ya == y0 && yb == y1) return;
    if (yb == y0) { copy(lowA, lowB); ya = yb; }
    else if (ya != y0) { load_low(pl, y0, lowA); ya = y0; }
    if (ya == y1) { copy(lowB, lowA); yb = ya; }
    else if (yb != y1) { load_low(pl, y1, lowB); yb = y1; }
  }
  template <typename PL>
  DVIS_HD void make_mid(const PL &pl, const Geom &g, int iy, float *dst) {
    const Tap ty = make_tap(iy, g.s1y, g.h);
    ensure_low(pl, ty.i0, ty.i1);
    DVIS_UNROLL
    for (int j = 0; j < 2 * PX; ++j) dst[j] = ty.w0 * lowA[j] + ty.w1 * lowB[j];
  }
  // make (midA, midB) hold intermediate rows (i0, i1)
  template <typename PL>
  DVIS_HD void ensure_mid(const PL &pl, const Geom &g, int i0, int i1) {
    if (ma == i0 && mb == i1) return;
    if (mb == i0) { copy(midA, midB); ma = mb; }
    else if (ma != i0) { make_mid(pl, g, i0, midA); ma = i0; }
    if (ma == i1) { copy(midB, midA); mb = ma; }
    else if (mb != i1) { make_mid(pl, g, i1, midB); mb = i1; }
  }
  // -> bit i of the result = (chain value of pixel ox0 + i on row oy) > 0
  template <typename PL>
  DVIS_HD uint32_t row(const PL &pl, const Geom &g, int oy) {
    const Tap t2y = make_tap(oy, g.s2y, g.Hc);
    ensure_mid(pl, g, t2y.i0, t2y.i1);
    uint32_t bits = 0;
    DVIS_UNROLL
    for (int i = 0; i < PX; ++i) {
      const float v = t2y.w0 * (w2x0[i] * midA[2 * i] + w2x1[i] * midA[2 * i + 1]) +
                      t2y.w1 * (w2x0[i] * midB[2 * i] + w2x1[i] * midB[2 * i + 1]);
      bits |= uint32_t(v > 0.f) << i;
    }
    return bits;
  }
};

// Source index range [lo, hi] touched by output indices [o0, o1] of one axis (taps are monotone in the output index).
// two_stage: through the second resize (crop size `mid`, scale s2) and then the first (source size `in`, scale s1).
DVIS_HD void source_range(int o0, int o1, bool two_stage, float s2, int mid, float s1, int in, int *lo, int *hi) {
  int a = o0, b = o1;
  if (two_stage) {
    a = make_tap(o0, s2, mid).i0;
    b = make_tap(o1, s2, mid).i1;
  }
  *lo = make_tap(a, s1, in).i0;
  *hi = make_tap(b, s1, in).i1;
}

// ---- vps: per-pixel arg-max over the kept queries of score * probability (py:897,917) -----------------------------
// visit(k, v) is called with every kept query's resized probability (for the "original area" counts, py:924).
// Returns the winner (first maximum, like torch.argmax) and whether ITS probability is >= 0.5 (py:925).
template <typename T, typename Visit>
DVIS_HD int vps_pixel(const T *logits, int64_t q_stride, const int64_t *keep_idx, const float *keep_score, int n_keep,
                      const Geom &g, const Tap &t2y, const Tap &t2x, bool *winner_solid, Visit &&visit) {
  int best = 0;
  float best_p = 0.f, best_v = 0.f;
  for (int k = 0; k < n_keep; ++k) {
    Plane<T> pl;
    pl.p = logits + keep_idx[k] * q_stride;
    pl.w = g.w;
    const float v = two_stage<true>(pl, g, t2y, t2x);
    visit(k, v);
    const float p = keep_score[k] * v;
    if (k == 0 || p > best_p) {
      best = k;
      best_p = p;
      best_v = v;
    }
  }
  *winner_solid = best_v >= 0.5f;
  return best;
}

}  // namespace rc
}  // namespace dvis
