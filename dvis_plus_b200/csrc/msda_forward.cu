// Multi-scale deformable attention forward for sm_100a.
//
// Semantics follow the reference kernel (OPS/src/cuda/ms_deform_im2col_cuda.cuh:242-304 with the bilinear
// tap of :38-89): pixel coordinate = loc * size - 0.5, a point contributes iff -1 < h < H and -1 < w < W,
// each of the 4 corners contributes iff it lies inside the map (zero padding).
//
// Design (B200): this op is a gather, bounded by L1/L2 line throughput rather than tensor math.
//   * a group of LPR = D*sizeof(T)/16 lanes owns one (query, head) item and reads each corner row of D
//     channels as one 16-byte load per lane (a full 128-byte line for D=32 fp32); a warp carries
//     G = 32/LPR independent items, so no cross-lane reduction is needed at all;
//   * all P points of a level are issued back to back (4*P independent 16-byte loads in flight per lane);
//   * a CTA walks a contiguous chunk of an optional `item_order` permutation -- the host orders items as
//     2-D query tiles per head so that the lines a CTA touches stay L1-resident (see host/locality.py);
//   * the fused variant also does softmax(logits) and loc = ref + offset / (W, H) in registers, removing the
//     sampling_locations / attention_weights round trip through HBM (OPS/modules/ms_deform_attn.py:101-112).
#include <algorithm>
#include <type_traits>

#include "common.cuh"

namespace dvis {
namespace {

constexpr int kMaxLevels = 8;
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

struct MsdaParams {
  const void *value;
  const int64_t *shapes;
  const int64_t *level_start;
  const void *loc;    // plain: sampling locations; fused: offsets (f32)
  const void *attn;   // plain: attention weights;  fused: logits (f32)
  const float *ref;   // fused only
  int64_t loc_stride;   // fused: elements between consecutive queries of `loc`
  int64_t attn_stride;  // fused: elements between consecutive queries of `attn`
  int ref_dim;
  const int32_t *order;
  void *out;
  int N, S, M, L, Lq, P;
  int items_per_cta;
};

template <typename T>
__device__ __forceinline__ Vec16<T> ldg16(const T *p) {
  Vec16<T> r;
  r.v = __ldg(reinterpret_cast<const decltype(r.v) *>(p));
  return r;
}

template <typename T>
struct AccOf { using type = float; };
template <>
struct AccOf<double> { using type = double; };

template <typename TO, typename A, int N>
__device__ __forceinline__ void store_vec(TO *dst, const A (&acc)[N]);

template <>
__device__ __forceinline__ void store_vec<float, float, 4>(float *dst, const float (&a)[4]) {
  *reinterpret_cast<float4 *>(dst) = make_float4(a[0], a[1], a[2], a[3]);
}
template <>
__device__ __forceinline__ void store_vec<double, double, 2>(double *dst, const double (&a)[2]) {
  *reinterpret_cast<double2 *>(dst) = make_double2(a[0], a[1]);
}
template <>
__device__ __forceinline__ void store_vec<float, float, 8>(float *dst, const float (&a)[8]) {
  reinterpret_cast<float4 *>(dst)[0] = make_float4(a[0], a[1], a[2], a[3]);
  reinterpret_cast<float4 *>(dst)[1] = make_float4(a[4], a[5], a[6], a[7]);
}
template <>
__device__ __forceinline__ void store_vec<__nv_bfloat16, float, 8>(__nv_bfloat16 *dst, const float (&a)[8]) {
  *reinterpret_cast<uint4 *>(dst) = make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]),
                                               pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7]));
}
template <>
__device__ __forceinline__ void store_vec<__nv_bfloat16, float, 4>(__nv_bfloat16 *dst, const float (&a)[4]) {
  *reinterpret_cast<uint2 *>(dst) = make_uint2(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]));
}

// One sampling point: 4 predicated corner loads + weighted accumulate.  `base` points at
// value[n, level_start, m, j*VEC]; consecutive pixels are `row` elements apart.
template <typename T, typename A, int VEC>
__device__ __forceinline__ void sample_point(const T *__restrict__ base, int H, int W, int row, A x, A y, A a,
                                             A (&acc)[VEC]) {
  // un-fused multiply / subtract like the reference (cuh:290-291) so that floor() sees the same value
  const A h_im = y * A(H) - A(0.5);
  const A w_im = x * A(W) - A(0.5);
  const bool inr = (h_im > A(-1)) && (w_im > A(-1)) && (h_im < A(H)) && (w_im < A(W));
  const A hf = floor(h_im), wf = floor(w_im);
  const int h0 = int(hf), w0 = int(wf);
  const A lh = h_im - hf, lw = w_im - wf;
  const A hh = A(1) - lh, hw = A(1) - lw;
  const bool top = inr && h0 >= 0, bot = inr && h0 + 1 <= H - 1;
  const bool lef = w0 >= 0, rig = w0 + 1 <= W - 1;
  const T *p00 = base + (long)(h0 * W + w0) * row;
  Vec16<T> v00{}, v01{}, v10{}, v11{};
  if (top && lef) v00 = ldg16(p00);
  if (top && rig) v01 = ldg16(p00 + row);
  if (bot && lef) v10 = ldg16(p00 + (long)W * row);
  if (bot && rig) v11 = ldg16(p00 + (long)W * row + row);
  const A w00 = hh * hw * a, w01 = hh * lw * a, w10 = lh * hw * a, w11 = lh * lw * a;
#pragma unroll
  for (int k = 0; k < VEC; ++k)
    acc[k] += w00 * A(v00.get(k)) + w01 * A(v01.get(k)) + w10 * A(v10.get(k)) + w11 * A(v11.get(k));
}

// ---------------------------------------------------------------------------------------------------
// vectorised kernel: T = value type, TO = output type, D = channels per head, PCT = points (0 = runtime)
// FUSED: loc/attn are raw offsets / logits + reference points.  LCT = levels for FUSED (softmax needs L*P regs)
// ---------------------------------------------------------------------------------------------------
template <typename T, typename TO, int D, int PCT, bool FUSED, int LCT>
__global__ void __launch_bounds__(kThreads) msda_fwd_vec_kernel(const MsdaParams p) {
  using A = typename AccOf<T>::type;
  constexpr int VEC = Vec16<T>::N;
  constexpr int LPR = D / VEC;  // lanes per row
  constexpr int G = 32 / LPR;   // items per warp
  static_assert(D % VEC == 0 && LPR >= 1 && LPR <= 32 && (LPR & (LPR - 1)) == 0, "unsupported head dim");

  __shared__ int sH[kMaxLevels], sW[kMaxLevels], sStart[kMaxLevels];
  if (threadIdx.x < p.L) {
    sH[threadIdx.x] = int(p.shapes[2 * threadIdx.x]);
    sW[threadIdx.x] = int(p.shapes[2 * threadIdx.x + 1]);
    sStart[threadIdx.x] = int(p.level_start[threadIdx.x]);
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / LPR, j = lane % LPR;
  const int n = blockIdx.y;
  const int M = p.M, L = FUSED ? LCT : p.L, P = PCT ? PCT : p.P;
  const int per_batch = p.Lq * M;
  const int row = M * D;
  const T *value_n = static_cast<const T *>(p.value) + (size_t)n * p.S * row;

  const int chunk_begin = blockIdx.x * p.items_per_cta;
  const int chunk_end = min(chunk_begin + p.items_per_cta, per_batch);
  for (int it = chunk_begin + warp * G + g; it < chunk_end; it += kWarps * G) {
    const int item = p.order ? p.order[it] : it;
    const int q = item / M, m = item - q * M;
    const size_t nq = (size_t)n * p.Lq + q;
    A acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = A(0);

    if constexpr (FUSED) {
      // softmax over L*P logits of this (query, head): OPS/modules/ms_deform_attn.py:103-104
      constexpr int LP = LCT * PCT;
      const float *lg = static_cast<const float *>(p.attn) + nq * p.attn_stride + (size_t)m * LP;
      const float *of = static_cast<const float *>(p.loc) + nq * p.loc_stride + (size_t)m * LP * 2;
      const float *rf = p.ref + nq * (size_t)(LCT * p.ref_dim);
      float w[LP];
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < LP; ++i) { w[i] = __ldg(lg + i); mx = fmaxf(mx, w[i]); }
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < LP; ++i) { w[i] = __expf(w[i] - mx); sum += w[i]; }
      const float inv = 1.f / sum;
#pragma unroll
      for (int l = 0; l < LCT; ++l) {
        const int H = sH[l], W = sW[l];
        const T *base = value_n + (size_t)sStart[l] * row + m * D + j * VEC;
        // loc = r + off * s with s = 1/(W_l, H_l) for 2-d reference points (py:106-109) or
        // s = 0.5 * (w, h) / P for reference boxes (py:110-112)
        const int rd = p.ref_dim;
        const float rx = __ldg(rf + rd * l), ry = __ldg(rf + rd * l + 1);
        const float sx = rd == 2 ? 1.f / float(W) : __ldg(rf + 4 * l + 2) * (0.5f / float(PCT));
        const float sy = rd == 2 ? 1.f / float(H) : __ldg(rf + 4 * l + 3) * (0.5f / float(PCT));
#pragma unroll
        for (int pt = 0; pt < PCT; ++pt) {
          const float2 o = __ldg(reinterpret_cast<const float2 *>(of) + l * PCT + pt);
          sample_point<T, A, VEC>(base, H, W, row, fmaf(o.x, sx, rx), fmaf(o.y, sy, ry), w[l * PCT + pt] * inv, acc);
        }
      }
    } else {
      const size_t lp_base = (nq * M + m) * (size_t)(L * P);
      const T *loc = static_cast<const T *>(p.loc) + lp_base * 2;
      const T *att = static_cast<const T *>(p.attn) + lp_base;
      for (int l = 0; l < L; ++l) {
        const int H = sH[l], W = sW[l];
        const T *base = value_n + (size_t)sStart[l] * row + m * D + j * VEC;
        if constexpr (PCT == 4 && std::is_same<T, float>::value) {
          // (x,y) of the 4 points of this level are 32 contiguous bytes, their weights 16 (host checked alignment)
          const float4 l01 = __ldg(reinterpret_cast<const float4 *>(loc) + 2 * l);
          const float4 l23 = __ldg(reinterpret_cast<const float4 *>(loc) + 2 * l + 1);
          const float4 a4 = __ldg(reinterpret_cast<const float4 *>(att) + l);
          sample_point<T, A, VEC>(base, H, W, row, l01.x, l01.y, a4.x, acc);
          sample_point<T, A, VEC>(base, H, W, row, l01.z, l01.w, a4.y, acc);
          sample_point<T, A, VEC>(base, H, W, row, l23.x, l23.y, a4.z, acc);
          sample_point<T, A, VEC>(base, H, W, row, l23.z, l23.w, a4.w, acc);
        } else {
          for (int pt = 0; pt < P; ++pt)
            sample_point<T, A, VEC>(base, H, W, row, A(__ldg(loc + 2 * (l * P + pt))),
                                    A(__ldg(loc + 2 * (l * P + pt) + 1)), A(__ldg(att + l * P + pt)), acc);
        }
      }
    }
    TO *dst = static_cast<TO *>(p.out) + (nq * M + m) * (size_t)D + j * VEC;
    store_vec<TO, A, VEC>(dst, acc);
  }
}

// ---------------------------------------------------------------------------------------------------
// generic kernel: any D, float or double; one thread per output element (the reference's mapping, cuh:258-266)
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) msda_fwd_generic_kernel(const MsdaParams p, int D, size_t total) {
  const T *value = static_cast<const T *>(p.value);
  const T *locs = static_cast<const T *>(p.loc);
  const T *attn = static_cast<const T *>(p.attn);
  T *out = static_cast<T *>(p.out);
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = int(idx % D);
    const size_t qm = idx / D;              // (n*Lq + q)*M + m
    const int m = int(qm % p.M);
    const size_t n = qm / p.M / p.Lq;
    const size_t row = (size_t)p.M * D;
    T acc = 0;
    for (int l = 0; l < p.L; ++l) {
      const int H = int(p.shapes[2 * l]), W = int(p.shapes[2 * l + 1]);
      const T *base = value + (n * p.S + (size_t)p.level_start[l]) * row + (size_t)m * D + c;
      for (int pt = 0; pt < p.P; ++pt) {
        const size_t i = (qm * p.L + l) * p.P + pt;
        const T x = locs[2 * i], y = locs[2 * i + 1], a = attn[i];
        const T h_im = y * T(H) - T(0.5), w_im = x * T(W) - T(0.5);
        if (h_im > T(-1) && w_im > T(-1) && h_im < T(H) && w_im < T(W)) {
          const T hf = floor(h_im), wf = floor(w_im);
          const int h0 = int(hf), w0 = int(wf), h1 = h0 + 1, w1 = w0 + 1;
          const T lh = h_im - hf, lw = w_im - wf, hh = T(1) - lh, hw = T(1) - lw;
          T v00 = 0, v01 = 0, v10 = 0, v11 = 0;
          if (h0 >= 0 && w0 >= 0) v00 = base[((size_t)h0 * W + w0) * row];
          if (h0 >= 0 && w1 <= W - 1) v01 = base[((size_t)h0 * W + w1) * row];
          if (h1 <= H - 1 && w0 >= 0) v10 = base[((size_t)h1 * W + w0) * row];
          if (h1 <= H - 1 && w1 <= W - 1) v11 = base[((size_t)h1 * W + w1) * row];
          acc += (hh * hw * v00 + hh * lw * v01 + lh * hw * v10 + lh * lw * v11) * a;
        }
      }
    }
    out[idx] = acc;
  }
}

int choose_items_per_cta(int per_batch, int batch, int items_per_pass) {
  // aim for >= ~6 CTAs per SM in flight across the grid, but keep chunks long enough for L1 reuse
  int ipc = 256;
  while (ipc > items_per_pass && (long)((per_batch + ipc - 1) / ipc) * batch < 6L * kNumSMs) ipc >>= 1;
  if (ipc < items_per_pass) ipc = items_per_pass;
  return ipc;
}

template <typename T, typename TO, int D, int PCT, bool FUSED, int LCT>
int launch_vec(MsdaParams p, cudaStream_t stream) {
  constexpr int G = 32 / (D / Vec16<T>::N);
  const int per_batch = p.Lq * p.M;
  p.items_per_cta = choose_items_per_cta(per_batch, p.N, kWarps * G);
  dim3 grid((per_batch + p.items_per_cta - 1) / p.items_per_cta, p.N);
  msda_fwd_vec_kernel<T, TO, D, PCT, FUSED, LCT><<<grid, kThreads, 0, stream>>>(p);
  return check_launch("msda_fwd_vec_kernel");
}

template <typename T>
int launch_plain(const MsdaParams &p, int D, cudaStream_t stream) {
  const bool vec_ok = aligned16(p.value) && aligned16(p.out);
  const bool p4_ok = p.P == 4 && aligned16(p.loc) && aligned16(p.attn);
  if (vec_ok && std::is_same<T, float>::value) {
#define DVIS_CASE(DD)                                                                  \
  case DD:                                                                             \
    return p4_ok ? launch_vec<float, float, DD, 4, false, 0>(p, stream)                \
                 : launch_vec<float, float, DD, 0, false, 0>(p, stream);
    switch (D) {
      DVIS_CASE(8)
      DVIS_CASE(16)
      DVIS_CASE(32)
      DVIS_CASE(64)
      DVIS_CASE(128)
      default: break;
    }
#undef DVIS_CASE
  }
  const size_t total = (size_t)p.N * p.Lq * p.M * D;
  const int blocks = int(std::min<size_t>((total + 255) / 256, (size_t)kNumSMs * 32));
  msda_fwd_generic_kernel<T><<<blocks, 256, 0, stream>>>(p, D, total);
  return check_launch("msda_fwd_generic_kernel");
}

int validate_common(const void *value, const int64_t *shapes, const int64_t *ls, const void *loc, const void *attn,
                    const void *out, int batch, int S, int M, int D, int L, int Lq, int P) {
  DVIS_REQUIRE(value && shapes && ls && loc && attn && out, "msda: null pointer argument");
  DVIS_REQUIRE(batch > 0 && S > 0 && M > 0 && D > 0 && L > 0 && Lq > 0 && P > 0, "msda: sizes must be positive");
  DVIS_REQUIRE(L <= kMaxLevels, "msda: num_levels %d > %d", L, kMaxLevels);
  DVIS_REQUIRE((long)S * M * D < (1L << 31), "msda: spatial_size*heads*channels must fit in int32");
  DVIS_REQUIRE((long)Lq * M < (1L << 31) && batch < 65536, "msda: query/batch extent too large");
  return DVIS_OK;
}

}  // namespace
}  // namespace dvis

using namespace dvis;

extern "C" int dvis_msda_forward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start,
                                 const void *sampling_loc, const void *attn_weight, int batch, int spatial_size,
                                 int num_heads, int channels, int num_levels, int num_query, int num_point, int dtype,
                                 const int32_t *item_order, void *out, void *stream) {
  if (int rc = validate_common(value, spatial_shapes, level_start, sampling_loc, attn_weight, out, batch, spatial_size,
                               num_heads, channels, num_levels, num_query, num_point))
    return rc;
  MsdaParams p{};
  p.value = value; p.shapes = spatial_shapes; p.level_start = level_start; p.loc = sampling_loc; p.attn = attn_weight;
  p.order = item_order; p.out = out;
  p.N = batch; p.S = spatial_size; p.M = num_heads; p.L = num_levels; p.Lq = num_query; p.P = num_point;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == DVIS_F32) return launch_plain<float>(p, channels, s);
  if (dtype == DVIS_F64) return launch_plain<double>(p, channels, s);
  return fail(DVIS_ERR_UNSUPPORTED, "msda_forward: dtype %d (the op is float/double only, like the reference)", dtype);
}

extern "C" int dvis_msda_fused_forward(const void *value, int value_dtype, const int64_t *spatial_shapes,
                                       const int64_t *level_start, const float *offsets, int64_t offsets_stride,
                                       const float *logits, int64_t logits_stride, const float *ref, int ref_dim,
                                       int batch, int spatial_size, int num_heads, int channels, int num_levels,
                                       int num_query, int num_point, const int32_t *item_order, void *out,
                                       int out_dtype, void *stream) {
  if (int rc = validate_common(value, spatial_shapes, level_start, offsets, logits, out, batch, spatial_size, num_heads,
                               channels, num_levels, num_query, num_point))
    return rc;
  DVIS_REQUIRE(ref && (ref_dim == 2 || ref_dim == 4), "msda_fused: reference points must be 2-d or 4-d");
  DVIS_REQUIRE(aligned16(value) && aligned16(out) && (reinterpret_cast<uintptr_t>(offsets) & 7u) == 0 &&
                   offsets_stride % 2 == 0,
               "msda_fused: value/out must be 16-byte aligned, offsets 8-byte aligned");
  MsdaParams p{};
  p.value = value; p.shapes = spatial_shapes; p.level_start = level_start; p.loc = offsets; p.attn = logits;
  p.ref = ref; p.ref_dim = ref_dim; p.loc_stride = offsets_stride; p.attn_stride = logits_stride;
  p.order = item_order; p.out = out;
  p.N = batch; p.S = spatial_size; p.M = num_heads; p.L = num_levels; p.Lq = num_query; p.P = num_point;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (num_point != 4)
    return fail(DVIS_ERR_UNSUPPORTED, "msda_fused: num_point %d (only 4 is built; use dvis_msda_forward)", num_point);
#define DVIS_FUSED(TV, TOUT, DD, LL) return launch_vec<TV, TOUT, DD, 4, true, LL>(p, s)
#define DVIS_FUSED_L(TV, TOUT, DD)                    \
  switch (num_levels) {                               \
    case 1: DVIS_FUSED(TV, TOUT, DD, 1);              \
    case 3: DVIS_FUSED(TV, TOUT, DD, 3);              \
    case 4: DVIS_FUSED(TV, TOUT, DD, 4);              \
    default: break;                                   \
  }
#define DVIS_FUSED_D(TV, TOUT)                        \
  switch (channels) {                                 \
    case 32: DVIS_FUSED_L(TV, TOUT, 32) break;        \
    case 64: DVIS_FUSED_L(TV, TOUT, 64) break;        \
    default: break;                                   \
  }
  if (value_dtype == DVIS_F32 && out_dtype == DVIS_F32) { DVIS_FUSED_D(float, float) }
  else if (value_dtype == DVIS_BF16 && out_dtype == DVIS_BF16) { DVIS_FUSED_D(__nv_bfloat16, __nv_bfloat16) }
  else if (value_dtype == DVIS_BF16 && out_dtype == DVIS_F32) { DVIS_FUSED_D(__nv_bfloat16, float) }
  else if (value_dtype == DVIS_F32 && out_dtype == DVIS_BF16) { DVIS_FUSED_D(float, __nv_bfloat16) }
#undef DVIS_FUSED_D
#undef DVIS_FUSED_L
#undef DVIS_FUSED
  return fail(DVIS_ERR_UNSUPPORTED,
              "msda_fused: no kernel for value_dtype=%d out_dtype=%d channels=%d levels=%d (built: f32/bf16, D in {32,64}, L in {1,3,4})",
              value_dtype, out_dtype, channels, num_levels);
}
