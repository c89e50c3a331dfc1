// Fused video post-processing: what DVIS_Plus_online.inference_video_vis / _vps / _vss do
// after the final mask GEMM (P/dvis_Plus/meta_architecture.py:818-979).
//
// The reference materialises, per kept query and frame, the mask logits at the padded input resolution (fp32), crops
// them, resizes again to the output resolution (fp32), and only then thresholds / arg-maxes: ~8 bytes of HBM traffic per
// full-resolution pixel and query, twice.  Here the whole chain is evaluated per OUTPUT pixel straight from the stride-4
// logits (which stay L2-resident: 235 KB per query-frame at 720p), so the only full-resolution traffic is the result:
// 1 byte per pixel (vis masks), 4 bytes per pixel and frame (vps ids), 8 bytes per pixel and frame (vss labels).
// All of it is gather/elementwise work: HBM-write bound, no tensor cores (the vss class contraction is the exception,
// see vss_argmax_kernel).  The per-pixel arithmetic lives in resize_core.cuh.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "resize_core.cuh"

namespace dvis {
namespace {

using rc::Geom;
using rc::Plane;
using rc::Tap;

template <typename T>
struct Raw;  // storage type the core's ld_elem understands
template <>
struct Raw<float> { using type = float; };
template <>
struct Raw<__nv_bfloat16> { using type = uint16_t; };

// ---- class scores + top-k ----------------------------------------------------------------------------------------
// scores[q, c] = softmax(cls[q, :])[c]; for c < K1 - 1 optionally max'ed with softmax(aux[q, :])[c]   (py:823-827,873-876)
__device__ void class_scores_rows(const float *__restrict__ cls, const float *__restrict__ aux, int Q, int K1,
                                  float *__restrict__ scores, int warp, int num_warps, int lane) {
  for (int q = warp; q < Q; q += num_warps) {
    const float *rows[2] = {cls + (size_t)q * K1, aux ? aux + (size_t)q * K1 : nullptr};
    float mx[2] = {-INFINITY, -INFINITY}, sum[2] = {0.f, 0.f};
    for (int r = 0; r < 2; ++r) {
      if (!rows[r]) continue;
      for (int c = lane; c < K1; c += 32) mx[r] = fmaxf(mx[r], rows[r][c]);
      for (int o = 16; o; o >>= 1) mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], o));
      for (int c = lane; c < K1; c += 32) sum[r] += expf(rows[r][c] - mx[r]);
      for (int o = 16; o; o >>= 1) sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], o);
    }
    for (int c = lane; c < K1; c += 32) {
      float s = expf(rows[0][c] - mx[0]) / sum[0];
      if (rows[1] && c < K1 - 1) s = fmaxf(s, expf(rows[1][c] - mx[1]) / sum[1]);
      scores[(size_t)q * K1 + c] = s;
    }
  }
}

__global__ void __launch_bounds__(256) class_scores_kernel(const float *cls, const float *aux, int Q, int K1, float *scores) {
  class_scores_rows(cls, aux, Q, K1, scores, blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), gridDim.x * (blockDim.x >> 5),
                    threadIdx.x & 31);
}

// Ordering of the top-k: larger score first, NaN above everything (torch.topk's convention), lower flat index on ties.
__device__ __forceinline__ bool topk_before(float a, int ia, float b, int ib) {
  const bool na = a != a, nb = b != b;
  if (na != nb) return na;
  if (!na && a != b) return a > b;
  return ia < ib;
}

// One CTA: class scores, then max_num rounds of a block-wide arg-max over the Q x (K1 - 1) object scores, each winner
// removed before the next round (py:831-835).  Entries that were taken hold -inf and a sentinel index that loses every
// comparison, so NaN / inf logits can neither stall the selection nor produce an out-of-range index.
__global__ void __launch_bounds__(1024) vis_topk_kernel(const float *cls, const float *aux, int Q, int K1, int max_num,
                                                        float *scores, float *out_scores, int64_t *out_labels,
                                                        int64_t *out_query) {
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  constexpr int kNone = 0x7fffffff;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  class_scores_rows(cls, aux, Q, K1, scores, warp, blockDim.x >> 5, lane);
  __syncthreads();
  const int K = K1 - 1, n = Q * K;
  for (int r = 0; r < max_num; ++r) {
    float best = -INFINITY;
    int arg = kNone;
    for (int j = tid; j < n; j += blockDim.x) {
      const float v = scores[(size_t)(j / K) * K1 + (j % K)];
      if (arg == kNone || topk_before(v, j, best, arg)) { best = v; arg = j; }
    }
    for (int o = 16; o; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
      if (oi != kNone && (arg == kNone || topk_before(ov, oi, best, arg))) { best = ov; arg = oi; }
    }
    if (lane == 0) { s_val[warp] = best; s_idx[warp] = arg; }
    __syncthreads();
    if (warp == 0) {
      best = lane < (blockDim.x >> 5) ? s_val[lane] : -INFINITY;
      arg = lane < (blockDim.x >> 5) ? s_idx[lane] : kNone;
      for (int o = 16; o; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
        if (oi != kNone && (arg == kNone || topk_before(ov, oi, best, arg))) { best = ov; arg = oi; }
      }
      if (lane == 0) {
        if (arg == kNone) arg = 0;                                // cannot happen for max_num <= n; never index out of range
        out_scores[r] = best;
        out_labels[r] = arg % K;
        out_query[r] = arg / K;
        scores[(size_t)(arg / K) * K1 + (arg % K)] = -INFINITY;  // taken
      }
    }
    __syncthreads();
  }
}

// ---- vis masks -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread4(uint32_t b) {  // bit i -> byte i
  return (b & 1u) | ((b & 2u) << 7) | ((b & 4u) << 14) | ((b & 8u) << 21);
}

constexpr int kStripPx = 8;     // output pixels per thread and row: one 8-byte store
constexpr int kStripWarps = 4;  // warps per CTA, each walking its own band of rows

// Second resize is the identity (output size == un-padded input size, the benchmark's 720p case): one bilinear sample
// per pixel.  A warp covers 256 consecutive pixels of a row and walks `rows_per_warp` rows, re-reading its source rows
// only when they change (every 4th output row at the 4x up-scale of the stride-4 logits).
// kPacked: one BIT per pixel instead of one byte -- out (planes, Ho, ceil(Wo / 8)), bit i of byte b = pixel 8 * b + i
// (numpy.unpackbits(bitorder="little")), bits past Wo are 0: 8x less to write and to copy to the host.
template <typename T, bool kPacked>
__global__ void __launch_bounds__(32 * kStripWarps) vis_masks_strip_kernel(const T *__restrict__ logits, int64_t q_stride,
                                                                           int64_t t_stride, const int64_t *__restrict__ sel,
                                                                           int frames, Geom g, int rows_per_warp,
                                                                           uint8_t *__restrict__ out) {
  using R = typename Raw<T>::type;
  const int lane = threadIdx.x, warp = threadIdx.y;
  const int plane = blockIdx.z, n = plane / frames, t = plane % frames;
  const int ox0 = (blockIdx.x * 32 + lane) * kStripPx;
  const int oy_begin = (blockIdx.y * kStripWarps + warp) * rows_per_warp;
  const int oy_end = min(oy_begin + rows_per_warp, g.Ho);
  if (ox0 >= g.Wo || oy_begin >= oy_end) return;
  const int64_t q = sel ? sel[n] : n;
  Plane<R> pl;
  pl.p = reinterpret_cast<const R *>(logits) + q * q_stride + t * t_stride;
  pl.w = g.w;
  rc::Strip<kStripPx, R> strip;
  strip.init(g, ox0);
  if constexpr (kPacked) {
    static_assert(kStripPx == 8, "one byte of 8 pixels per thread");
    const int Wb = (g.Wo + 7) >> 3;
    const uint32_t valid = ox0 + 8 <= g.Wo ? 0xffu : (1u << (g.Wo - ox0)) - 1u;          // pixels of this byte inside the row
    uint8_t *ob = out + (int64_t)plane * g.Ho * Wb + (ox0 >> 3);
    for (int oy = oy_begin; oy < oy_end; ++oy) ob[(int64_t)oy * Wb] = uint8_t(strip.row(pl, g, oy) & valid);
    return;
  }
  uint8_t *o = out + (int64_t)plane * g.Ho * g.Wo + ox0;
  const bool vec = (g.Wo % kStripPx == 0) && ((reinterpret_cast<uintptr_t>(out) & 7u) == 0);
  for (int oy = oy_begin; oy < oy_end; ++oy) {
    const uint32_t bits = strip.row(pl, g, oy);
    uint8_t *orow = o + (int64_t)oy * g.Wo;
    if (vec) {
      *reinterpret_cast<uint2 *>(orow) = make_uint2(spread4(bits & 15u), spread4(bits >> 4));
    } else {
      for (int i = 0; i < kStripPx; ++i)
        if (ox0 + i < g.Wo) orow[i] = uint8_t((bits >> i) & 1u);
    }
  }
}

// General chain (second resize changes the size): the same walk with two cache levels (rc::Strip2) -- source rows and
// intermediate rows are each computed once per strip when they first enter, instead of 16 source reads per output pixel.
constexpr int kStrip2Px = 4;    // output pixels per thread and row: one 4-byte store, a warp stores 128 contiguous bytes

template <typename T, bool kPacked>
__global__ void __launch_bounds__(32 * kStripWarps) vis_masks_two_stage_kernel(const T *__restrict__ logits, int64_t q_stride,
                                                                               int64_t t_stride, const int64_t *__restrict__ sel,
                                                                               int frames, Geom g, int rows_per_warp,
                                                                               uint8_t *__restrict__ out) {
  using R = typename Raw<T>::type;
  const int lane = threadIdx.x, warp = threadIdx.y;
  const int plane = blockIdx.z, n = plane / frames, t = plane % frames;
  const int ox0 = (blockIdx.x * 32 + lane) * kStrip2Px;
  const int oy_begin = (blockIdx.y * kStripWarps + warp) * rows_per_warp;
  const int oy_end = min(oy_begin + rows_per_warp, g.Ho);
  if (oy_begin >= oy_end) return;                       // warp-uniform
  // packed: lanes past the row end stay (their columns are clamped, their bits masked off) -- the nibble exchange below
  // needs every lane of the warp
  if (!kPacked && ox0 >= g.Wo) return;
  const int64_t q = sel ? sel[n] : n;
  Plane<R> pl;
  pl.p = reinterpret_cast<const R *>(logits) + q * q_stride + t * t_stride;
  pl.w = g.w;
  rc::Strip2<kStrip2Px, R> strip;
  strip.init(g, ox0);
  if constexpr (kPacked) {
    static_assert(kStrip2Px == 4, "two lanes of 4 pixels per byte");
    const int Wb = (g.Wo + 7) >> 3;
    const int left = g.Wo - ox0;                                                           // pixels of this nibble inside the row
    const uint32_t valid = left >= 4 ? 0xfu : left > 0 ? (1u << left) - 1u : 0u;
    uint8_t *ob = out + (int64_t)plane * g.Ho * Wb + (ox0 >> 3);
    for (int oy = oy_begin; oy < oy_end; ++oy) {
      const uint32_t mine = strip.row(pl, g, oy) & valid;
      const uint32_t hi = __shfl_down_sync(0xffffffffu, mine, 1);                          // the odd neighbour's nibble
      if ((lane & 1) == 0 && ox0 < g.Wo) ob[(int64_t)oy * Wb] = uint8_t(mine | (hi << 4));
    }
    return;
  }
  uint8_t *o = out + (int64_t)plane * g.Ho * g.Wo + ox0;
  const bool vec = (g.Wo % kStrip2Px == 0) && ((reinterpret_cast<uintptr_t>(out) & 3u) == 0);
  for (int oy = oy_begin; oy < oy_end; ++oy) {
    const uint32_t bits = strip.row(pl, g, oy);
    uint8_t *orow = o + (int64_t)oy * g.Wo;
    if (vec) {
      *reinterpret_cast<uint32_t *>(orow) = spread4(bits);
    } else {
      for (int i = 0; i < kStrip2Px; ++i)
        if (ox0 + i < g.Wo) orow[i] = uint8_t((bits >> i) & 1u);
    }
  }
}

// CTA-tiled variant of both strip kernels (selected with DVIS_VIS_MASKS_TILED=1; not yet timed on a B200, the strip
// kernels above are the measured default).  The strip kernels stall on dependent L2 loads every time a source row enters
// (26 % / 7 % of the HBM roofline); here the source window of the CTA's whole output tile is staged in shared memory once
// -- one coalesced, latency-exposed read per CTA, converted to f32 -- and the very same walkers then run on the window,
// so the results are bit-identical to the strip kernels'.
template <typename T, bool kTwoStage, bool kPacked>
__global__ void __launch_bounds__(32 * kStripWarps) vis_masks_tiled_kernel(const T *__restrict__ logits, int64_t q_stride,
                                                                           int64_t t_stride, const int64_t *__restrict__ sel,
                                                                           int frames, Geom g, int rows_per_warp,
                                                                           uint8_t *__restrict__ out) {
  using R = typename Raw<T>::type;
  constexpr int PX = kTwoStage ? kStrip2Px : kStripPx;
  extern __shared__ float s_win[];
  const int lane = threadIdx.x, warp = threadIdx.y, tid = warp * 32 + lane;
  const int plane = blockIdx.z, n = plane / frames, t = plane % frames;
  const int64_t q = sel ? sel[n] : n;
  const R *src = reinterpret_cast<const R *>(logits) + q * q_stride + t * t_stride;
  // the CTA's output tile and its source window
  const int tx0 = blockIdx.x * 32 * PX, tx1 = min(tx0 + 32 * PX, g.Wo) - 1;
  const int ty0 = blockIdx.y * kStripWarps * rows_per_warp, ty1 = min(ty0 + kStripWarps * rows_per_warp, g.Ho) - 1;
  int r_lo, r_hi, c_lo, c_hi;
  rc::source_range(ty0, ty1, kTwoStage, g.s2y, g.Hc, g.s1y, g.h, &r_lo, &r_hi);
  rc::source_range(tx0, tx1, kTwoStage, g.s2x, g.Wc, g.s1x, g.w, &c_lo, &c_hi);
  const int ww = c_hi - c_lo + 1, cells = (r_hi - r_lo + 1) * ww;      // the host sized the dynamic shared memory for this
  for (int i = tid; i < cells; i += 32 * kStripWarps) {
    const int y = i / ww, x = i - y * ww;
    s_win[i] = rc::ld_elem(src + (int64_t)(r_lo + y) * g.w + c_lo + x);
  }
  __syncthreads();
  rc::Window pl;
  pl.p = s_win;
  pl.w = ww;
  pl.y0 = r_lo;
  pl.x0 = c_lo;

  const int ox0 = tx0 + lane * PX;
  const int oy_begin = ty0 + warp * rows_per_warp;
  const int oy_end = min(oy_begin + rows_per_warp, g.Ho);
  if (oy_begin >= oy_end) return;                                        // warp-uniform
  if (!(kTwoStage && kPacked) && ox0 >= g.Wo) return;                    // packed two-stage: every lane feeds the nibble exchange
  typename std::conditional<kTwoStage, rc::Strip2<PX, R>, rc::Strip<PX, R>>::type strip;
  strip.init(g, ox0);
  if constexpr (kPacked) {
    const int Wb = (g.Wo + 7) >> 3;
    const int left = g.Wo - ox0;
    const uint32_t valid = left >= PX ? (1u << PX) - 1u : left > 0 ? (1u << left) - 1u : 0u;
    uint8_t *ob = out + (int64_t)plane * g.Ho * Wb + (ox0 >> 3);
    for (int oy = oy_begin; oy < oy_end; ++oy) {
      const uint32_t mine = strip.row(pl, g, oy) & valid;
      if constexpr (kTwoStage) {
        const uint32_t hi = __shfl_down_sync(0xffffffffu, mine, 1);
        if ((lane & 1) == 0 && ox0 < g.Wo) ob[(int64_t)oy * Wb] = uint8_t(mine | (hi << 4));
      } else {
        ob[(int64_t)oy * Wb] = uint8_t(mine);
      }
    }
    return;
  }
  uint8_t *o = out + (int64_t)plane * g.Ho * g.Wo + ox0;
  const bool vec = (g.Wo % PX == 0) && ((reinterpret_cast<uintptr_t>(out) & (PX - 1)) == 0);
  for (int oy = oy_begin; oy < oy_end; ++oy) {
    const uint32_t bits = strip.row(pl, g, oy);
    uint8_t *orow = o + (int64_t)oy * g.Wo;
    if (vec) {
      if constexpr (PX == 8) *reinterpret_cast<uint2 *>(orow) = make_uint2(spread4(bits & 15u), spread4(bits >> 4));
      else *reinterpret_cast<uint32_t *>(orow) = spread4(bits);
    } else {
      for (int i = 0; i < PX; ++i)
        if (ox0 + i < g.Wo) orow[i] = uint8_t((bits >> i) & 1u);
    }
  }
}

// ---- vps ---------------------------------------------------------------------------------------------------------
// One thread per output pixel of frame blockIdx.z.  win[t, y, x] = k (winner's probability >= 0.5) or ~k (< 0.5);
// areas[0:n] = #pixels won by k (py:923), areas[n:2n] = #pixels with probability_k >= 0.5 (py:924),
// areas[2n:3n] = #pixels won by k with probability_k >= 0.5 (py:925-926).
struct AreaVisitor {
  int *cnt;
  int n_keep;
  bool active;
  int lane;
  __device__ __forceinline__ void operator()(int k, float v) const {
    const unsigned m = __ballot_sync(0xffffffffu, active && v >= 0.5f);
    if (lane == 0 && m) atomicAdd(&cnt[n_keep + k], __popc(m));
  }
};

template <typename T>
__global__ void __launch_bounds__(128) vps_argmax_kernel(const T *__restrict__ logits, int64_t q_stride, int64_t t_stride,
                                                         const int64_t *__restrict__ keep_idx,
                                                         const float *__restrict__ keep_score, int n_keep, Geom g,
                                                         int32_t *__restrict__ win, unsigned long long *__restrict__ areas) {
  using R = typename Raw<T>::type;
  extern __shared__ int cnt[];  // 3 * n_keep
  for (int i = threadIdx.x; i < 3 * n_keep; i += blockDim.x) cnt[i] = 0;
  __syncthreads();
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y, t = blockIdx.z;
  const bool active = ox < g.Wo;
  const Tap t2y = rc::make_tap(oy, g.s2y, g.Hc);
  const Tap t2x = rc::make_tap(active ? ox : g.Wo - 1, g.s2x, g.Wc);
  AreaVisitor visit;
  visit.cnt = cnt;
  visit.n_keep = n_keep;
  visit.active = active;
  visit.lane = threadIdx.x & 31;
  bool solid = false;
  const int best = rc::vps_pixel(reinterpret_cast<const R *>(logits) + t * t_stride, q_stride, keep_idx, keep_score, n_keep,
                                 g, t2y, t2x, &solid, visit);
  if (active) {
    atomicAdd(&cnt[best], 1);
    if (solid) atomicAdd(&cnt[2 * n_keep + best], 1);
    win[((int64_t)t * g.Ho + oy) * g.Wo + ox] = solid ? best : ~best;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * n_keep; i += blockDim.x)
    if (cnt[i]) atomicAdd(&areas[i], (unsigned long long)cnt[i]);
}

// panoptic[i] = segment id of the pixel's winner if the winner's probability is >= 0.5, else 0 (py:925,934,939)
__global__ void __launch_bounds__(256) vps_paint_kernel(const int32_t *__restrict__ win, const int32_t *__restrict__ seg_of_k,
                                                        int64_t total, int32_t *__restrict__ panoptic) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t k = win[i];
    panoptic[i] = k >= 0 ? seg_of_k[k] : 0;
  }
}

// ---- vss ---------------------------------------------------------------------------------------------------------
// semseg[c] = sum_q mask_cls[q, c] * probability_q (py:972), label = first arg-max over c (py:973).  One thread per
// output pixel: the Q resized probabilities of the pixel are computed once into the thread's own shared-memory column,
// then the K classes are accumulated 8 at a time (mask_cls reads are warp-uniform).  fp32 CUDA-core math, the same
// precision as the reference's einsum; the contraction is K x Q MACs per pixel and would sit on the tensor pipe with
// tf32 operands -- left as the next step for this kernel, see DESIGN.md.
constexpr int kVssThreads = 128;
constexpr int kVssClasses = 8;

template <typename T>
__global__ void __launch_bounds__(kVssThreads) vss_argmax_kernel(const T *__restrict__ logits, int64_t q_stride,
                                                                 int64_t t_stride, const float *__restrict__ mask_cls,
                                                                 int64_t cls_stride, int Q, int K, Geom g,
                                                                 int64_t *__restrict__ out) {
  using R = typename Raw<T>::type;
  extern __shared__ float probs[];  // [Q][kVssThreads]
  const int ox = blockIdx.x * kVssThreads + threadIdx.x, oy = blockIdx.y, t = blockIdx.z;
  if (ox >= g.Wo) return;           // no block-wide synchronisation below: every thread owns its column
  const Tap t2y = rc::make_tap(oy, g.s2y, g.Hc);
  const Tap t2x = rc::make_tap(ox, g.s2x, g.Wc);
  float *mine = probs + threadIdx.x;
  for (int q = 0; q < Q; ++q) {
    Plane<R> pl;
    pl.p = reinterpret_cast<const R *>(logits) + q * q_stride + t * t_stride;
    pl.w = g.w;
    mine[q * kVssThreads] = rc::two_stage<true>(pl, g, t2y, t2x);
  }
  float best = 0.f;
  int arg = 0;
  for (int c0 = 0; c0 < K; c0 += kVssClasses) {
    float acc[kVssClasses];
#pragma unroll
    for (int j = 0; j < kVssClasses; ++j) acc[j] = 0.f;
    for (int q = 0; q < Q; ++q) {
      const float m = mine[q * kVssThreads];
      const float *row = mask_cls + q * cls_stride + c0;
#pragma unroll
      for (int j = 0; j < kVssClasses; ++j)
        if (c0 + j < K) acc[j] += __ldg(row + j) * m;
    }
#pragma unroll
    for (int j = 0; j < kVssClasses; ++j)
      if (c0 + j < K && (c0 + j == 0 || acc[j] > best)) { best = acc[j]; arg = c0 + j; }
  }
  out[((int64_t)t * g.Ho + oy) * g.Wo + ox] = arg;
}

int check_geom(const char *who, int frames, int h, int w, int H1, int W1, int Hc, int Wc, int Ho, int Wo) {
  DVIS_REQUIRE(frames > 0 && h > 0 && w > 0 && H1 > 0 && W1 > 0 && Ho > 0 && Wo > 0, "%s: sizes must be positive", who);
  DVIS_REQUIRE(Hc > 0 && Wc > 0 && Hc <= H1 && Wc <= W1, "%s: crop (%d, %d) must lie inside the first resize (%d, %d)", who, Hc, Wc, H1, W1);
  DVIS_REQUIRE(Ho <= 65535 && frames <= 65535, "%s: output height and frame count are limited to 65535 per launch", who);
  return DVIS_OK;
}

}  // namespace
}  // namespace dvis

using namespace dvis;

extern "C" int dvis_class_scores(const float *pred_cls, const float *aux_cls, int Q, int K1, float *scores, void *stream) {
  DVIS_REQUIRE(pred_cls && scores, "class_scores: null pointer argument");
  DVIS_REQUIRE(Q > 0 && K1 > 1, "class_scores: need Q > 0 and at least one class besides no-object");
  class_scores_kernel<<<(Q + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(pred_cls, aux_cls, Q, K1, scores);
  return check_launch("class_scores_kernel");
}

extern "C" int dvis_vis_topk(const float *pred_cls, const float *aux_cls, int Q, int K1, int max_num, float *scores_workspace,
                             float *out_scores, int64_t *out_labels, int64_t *out_query, void *stream) {
  DVIS_REQUIRE(pred_cls && scores_workspace && out_scores && out_labels && out_query, "vis_topk: null pointer argument");
  DVIS_REQUIRE(Q > 0 && K1 > 1, "vis_topk: need Q > 0 and at least one class besides no-object");
  // torch.topk raises for k larger than the number of candidates (py:831)
  DVIS_REQUIRE(max_num > 0 && (int64_t)max_num <= (int64_t)Q * (K1 - 1), "vis_topk: selected index k out of range (max_num %d, %lld scores)",
               max_num, (long long)Q * (K1 - 1));
  DVIS_REQUIRE((int64_t)Q * (K1 - 1) < (int64_t)1 << 31, "vis_topk: too many scores");
  vis_topk_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(pred_cls, aux_cls, Q, K1, max_num, scores_workspace,
                                                                     out_scores, out_labels, out_query);
  return check_launch("vis_topk_kernel");
}

namespace dvis {
namespace {
template <bool kPacked>
int launch_vis_masks(const void *logits, int logits_dtype, int64_t q_stride, int64_t t_stride, const int64_t *sel, int n_sel,
                     int frames, int h, int w, int H1, int W1, int Hc, int Wc, int Ho, int Wo, uint8_t *out, void *stream) {
  DVIS_REQUIRE(logits && out, "vis_masks: null pointer argument");
  DVIS_REQUIRE(n_sel > 0, "vis_masks: nothing selected");
  if (int rc_ = check_geom("vis_masks", frames, h, w, H1, W1, Hc, Wc, Ho, Wo)) return rc_;
  DVIS_REQUIRE((int64_t)n_sel * frames <= 65535, "vis_masks: n_sel * frames is limited to 65535 per launch");
  if (logits_dtype != DVIS_F32 && logits_dtype != DVIS_BF16) return fail(DVIS_ERR_UNSUPPORTED, "vis_masks: logits dtype must be f32 or bf16");
  const Geom g = rc::make_geom(h, w, H1, W1, Hc, Wc, Ho, Wo);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned planes = unsigned(n_sel * frames);
  const dim3 block(32, kStripWarps);
  const char *tiled_env = getenv("DVIS_VIS_MASKS_TILED");
  if (tiled_env && atoi(tiled_env) != 0) {
    const bool two = !(Ho == Hc && Wo == Wc);
    const int px = two ? kStrip2Px : kStripPx, rows_per_warp = two ? 24 : 16;
    const int tile_w = 32 * px, tile_h = kStripWarps * rows_per_warp;
    // exact size of the largest source window over all tile rows / tile columns (a few hundred make_tap evaluations)
    int max_rows = 0, max_cols = 0, lo, hi;
    for (int y0 = 0; y0 < Ho; y0 += tile_h) {
      rc::source_range(y0, std::min(y0 + tile_h, Ho) - 1, two, g.s2y, g.Hc, g.s1y, g.h, &lo, &hi);
      max_rows = std::max(max_rows, hi - lo + 1);
    }
    for (int x0 = 0; x0 < Wo; x0 += tile_w) {
      rc::source_range(x0, std::min(x0 + tile_w, Wo) - 1, two, g.s2x, g.Wc, g.s1x, g.w, &lo, &hi);
      max_cols = std::max(max_cols, hi - lo + 1);
    }
    const size_t smem = sizeof(float) * (size_t)max_rows * max_cols;
    if (smem <= 48 * 1024) {                    // strongly down-scaling chains have windows too large to stage: strip kernels
      const dim3 grid((Wo + tile_w - 1) / tile_w, (Ho + tile_h - 1) / tile_h, planes);
#define DVIS_TILED(TWO)                                                                                                          \
      do {                                                                                                                       \
        if (logits_dtype == DVIS_F32)                                                                                            \
          vis_masks_tiled_kernel<float, TWO, kPacked><<<grid, block, smem, s>>>(static_cast<const float *>(logits), q_stride, t_stride, sel, frames, g, rows_per_warp, out); \
        else                                                                                                                     \
          vis_masks_tiled_kernel<__nv_bfloat16, TWO, kPacked><<<grid, block, smem, s>>>(static_cast<const __nv_bfloat16 *>(logits), q_stride, t_stride, sel, frames, g, rows_per_warp, out); \
      } while (0)
      if (two) DVIS_TILED(true); else DVIS_TILED(false);
#undef DVIS_TILED
      return check_launch("vis_masks_tiled_kernel");
    }
  }
  if (Ho == Hc && Wo == Wc) {
    // rows per warp: enough CTAs to fill the machine a few times over, bands long enough to reuse source rows
    const int rows_per_warp = 16;
    const dim3 grid((Wo + 32 * kStripPx - 1) / (32 * kStripPx), (Ho + kStripWarps * rows_per_warp - 1) / (kStripWarps * rows_per_warp), planes);
    if (logits_dtype == DVIS_F32)
      vis_masks_strip_kernel<float, kPacked><<<grid, block, 0, s>>>(static_cast<const float *>(logits), q_stride, t_stride, sel, frames, g, rows_per_warp, out);
    else
      vis_masks_strip_kernel<__nv_bfloat16, kPacked><<<grid, block, 0, s>>>(static_cast<const __nv_bfloat16 *>(logits), q_stride, t_stride, sel, frames, g, rows_per_warp, out);
    return check_launch("vis_masks_strip_kernel");
  }
  const int rows_per_warp = 24;
  const dim3 grid((Wo + 32 * kStrip2Px - 1) / (32 * kStrip2Px), (Ho + kStripWarps * rows_per_warp - 1) / (kStripWarps * rows_per_warp), planes);
  if (logits_dtype == DVIS_F32)
    vis_masks_two_stage_kernel<float, kPacked><<<grid, block, 0, s>>>(static_cast<const float *>(logits), q_stride, t_stride, sel, frames, g, rows_per_warp, out);
  else
    vis_masks_two_stage_kernel<__nv_bfloat16, kPacked><<<grid, block, 0, s>>>(static_cast<const __nv_bfloat16 *>(logits), q_stride, t_stride, sel, frames, g, rows_per_warp, out);
  return check_launch("vis_masks_two_stage_kernel");
}
}  // namespace
}  // namespace dvis

extern "C" int dvis_vis_masks(const void *logits, int logits_dtype, int64_t q_stride, int64_t t_stride, const int64_t *sel,
                              int n_sel, int frames, int h, int w, int H1, int W1, int Hc, int Wc, int Ho, int Wo,
                              uint8_t *out, void *stream) {
  return launch_vis_masks<false>(logits, logits_dtype, q_stride, t_stride, sel, n_sel, frames, h, w, H1, W1, Hc, Wc, Ho, Wo, out, stream);
}

extern "C" int dvis_vis_masks_packed(const void *logits, int logits_dtype, int64_t q_stride, int64_t t_stride, const int64_t *sel,
                                     int n_sel, int frames, int h, int w, int H1, int W1, int Hc, int Wc, int Ho, int Wo,
                                     uint8_t *out, void *stream) {
  return launch_vis_masks<true>(logits, logits_dtype, q_stride, t_stride, sel, n_sel, frames, h, w, H1, W1, Hc, Wc, Ho, Wo, out, stream);
}

extern "C" int dvis_vps_argmax(const void *logits, int logits_dtype, int64_t q_stride, int64_t t_stride, const int64_t *keep_idx,
                               const float *keep_score, int n_keep, int frames, int h, int w, int H1, int W1, int Hc, int Wc,
                               int Ho, int Wo, int32_t *win, unsigned long long *areas, void *stream) {
  DVIS_REQUIRE(logits && keep_idx && keep_score && win && areas, "vps_argmax: null pointer argument");
  DVIS_REQUIRE(n_keep > 0 && n_keep <= 4096, "vps_argmax: 1 <= n_keep <= 4096");
  if (int rc_ = check_geom("vps_argmax", frames, h, w, H1, W1, Hc, Wc, Ho, Wo)) return rc_;
  const Geom g = rc::make_geom(h, w, H1, W1, Hc, Wc, Ho, Wo);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(areas, 0, sizeof(unsigned long long) * 3 * n_keep, s);
  if (e != cudaSuccess) return fail(DVIS_ERR_CUDA, "vps_argmax: memset: %s", cudaGetErrorString(e));
  const dim3 grid((Wo + 127) / 128, Ho, frames);
  const size_t smem = sizeof(int) * 3 * n_keep;
  if (logits_dtype == DVIS_F32)
    vps_argmax_kernel<float><<<grid, 128, smem, s>>>(static_cast<const float *>(logits), q_stride, t_stride, keep_idx, keep_score, n_keep, g, win, areas);
  else if (logits_dtype == DVIS_BF16)
    vps_argmax_kernel<__nv_bfloat16><<<grid, 128, smem, s>>>(static_cast<const __nv_bfloat16 *>(logits), q_stride, t_stride, keep_idx, keep_score, n_keep, g, win, areas);
  else
    return fail(DVIS_ERR_UNSUPPORTED, "vps_argmax: logits dtype must be f32 or bf16");
  return check_launch("vps_argmax_kernel");
}

extern "C" int dvis_vps_paint(const int32_t *win, const int32_t *seg_of_k, int64_t total, int32_t *panoptic, void *stream) {
  DVIS_REQUIRE(win && seg_of_k && panoptic, "vps_paint: null pointer argument");
  DVIS_REQUIRE(total > 0, "vps_paint: nothing to paint");
  const int blocks = int(std::min<int64_t>((total + 255) / 256, (int64_t)kNumSMs * 16));
  vps_paint_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(win, seg_of_k, total, panoptic);
  return check_launch("vps_paint_kernel");
}

extern "C" int dvis_vss_argmax(const void *logits, int logits_dtype, int64_t q_stride, int64_t t_stride, const float *mask_cls,
                               int64_t cls_stride, int Q, int K, int frames, int h, int w, int H1, int W1, int Hc, int Wc,
                               int Ho, int Wo, int64_t *out, void *stream) {
  DVIS_REQUIRE(logits && mask_cls && out, "vss_argmax: null pointer argument");
  DVIS_REQUIRE(Q > 0 && K > 0 && cls_stride >= K, "vss_argmax: need Q > 0, K > 0, cls_stride >= K");
  if (int rc_ = check_geom("vss_argmax", frames, h, w, H1, W1, Hc, Wc, Ho, Wo)) return rc_;
  const size_t smem = sizeof(float) * (size_t)Q * kVssThreads;
  if (smem > 200 * 1024) return fail(DVIS_ERR_UNSUPPORTED, "vss_argmax: Q = %d needs %zu bytes of shared memory (limit 400 queries)", Q, smem);
  const Geom g = rc::make_geom(h, w, H1, W1, Hc, Wc, Ho, Wo);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const dim3 grid((Wo + kVssThreads - 1) / kVssThreads, Ho, frames);
  cudaError_t e;
  if (logits_dtype == DVIS_F32) {
    e = cudaFuncSetAttribute(vss_argmax_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return fail(DVIS_ERR_CUDA, "vss_argmax: shared memory opt-in: %s", cudaGetErrorString(e));
    vss_argmax_kernel<float><<<grid, kVssThreads, smem, s>>>(static_cast<const float *>(logits), q_stride, t_stride, mask_cls, cls_stride, Q, K, g, out);
  } else if (logits_dtype == DVIS_BF16) {
    e = cudaFuncSetAttribute(vss_argmax_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return fail(DVIS_ERR_CUDA, "vss_argmax: shared memory opt-in: %s", cudaGetErrorString(e));
    vss_argmax_kernel<__nv_bfloat16><<<grid, kVssThreads, smem, s>>>(static_cast<const __nv_bfloat16 *>(logits), q_stride, t_stride, mask_cls, cls_stride, Q, K, g, out);
  } else {
    return fail(DVIS_ERR_UNSUPPORTED, "vss_argmax: logits dtype must be f32 or bf16");
  }
  return check_launch("vss_argmax_kernel");
}
