// GroupNorm for channels-last (N, HW, C) maps, split into a statistics pass and a fused apply pass.
//
// The pixel decoder normalises every conv output with GroupNorm(32, C): input_proj (P/mask2former/modeling/
// pixel_decoder/msdeformattn.py:213-226,321), FPN lateral and output convs (:262-286,346-351).  The reference runs them
// on NCHW fp32 tensors through at::group_norm plus separate add / interpolate / relu / cast kernels; here the maps stay
// channels-last (the layout the GEMMs and the tcgen05 mask GEMM want) and the apply pass fuses what follows the norm:
//   y = GN(x) [+ bilinear_upsample(low-res map)] [ReLU]          (FPN top-down add, msdeformattn.py:349; F.relu :278)
//   outputs (each optional): fp32, GEMM-dtype copy, GEMM-dtype (y + pos)   (encoder inputs: src, with_pos_embed(src, pos))
// written with an arbitrary batch stride so each level lands directly in its slice of the (N, S, C) token buffer
// (replaces flatten / transpose / cat, msdeformattn.py:70-80).
#include "common.cuh"

namespace dvis {
namespace {

template <typename T>
__device__ __forceinline__ float4 ld4(const T *p);
template <>
__device__ __forceinline__ float4 ld4<float>(const float *p) { return *reinterpret_cast<const float4 *>(p); }
template <>
__device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16 *p) {
  const uint2 u = *reinterpret_cast<const uint2 *>(p);
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                     __uint_as_float(u.y & 0xffff0000u));
}
template <typename T>
__device__ __forceinline__ void st4(T *p, float4 v);
template <>
__device__ __forceinline__ void st4<float>(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
template <>
__device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16 *p, float4 v) {
  *reinterpret_cast<uint2 *>(p) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

constexpr int kMaxGroups = 64;

// ---- pass 1: per (n, group) shifted sum and sum of squares ------------------------------------------------------
// Deterministic (no atomics): the statistics -- and everything downstream: bf16 roundings, thresholded attention masks,
// Hungarian decisions -- are bit-identical from run to run.  The first version combined partial sums with shared-memory
// float atomics, whose arrival order made two runs of the same clip differ by one bf16 ulp (profiles/r2_determinism.md).
//   gn_partial_kernel: a CTA reads a contiguous chunk of 256 pixels x all channels (fully coalesced, the map is read once
//     at HBM speed), every thread reduces its pixels in a fixed order, threads of a group are combined in a fixed order
//     -> one (sum, sumsq) double pair per (sample, chunk, group) in the workspace;
//   gn_finalize_kernel: one thread per (sample, group) adds the chunks' partials in chunk order.
// Values are accumulated relative to a pivot K = x[n, pixel 0, first channel of the group] (shifted-data variance: no
// cancellation when |mean| >> std), fp32 within a thread's <= 64 pixels, double across threads and chunks.
constexpr int kGnChunk = 256;   // pixels per CTA of the partial pass

template <typename T>
__device__ __forceinline__ float gn_pivot(const T *x, int64_t batch_stride, int n, int g, int cpg) {
  return float(ld4<T>(x + (size_t)n * batch_stride + (size_t)g * cpg).x);
}

template <typename T>
__global__ void __launch_bounds__(256) gn_partial_kernel(const T *__restrict__ x, int64_t batch_stride, int HW, int C, int G,
                                                         double *__restrict__ partials) {
  __shared__ float s_s[256], s_q[256];
  const int n = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
  const int quads = C / 4, cpg = C / G;
  const int cq = threadIdx.x % quads, pl = threadIdx.x / quads, prow = blockDim.x / quads;
  const int p0 = chunk * kGnChunk, p1 = min(p0 + kGnChunk, HW);
  const int g = (cq * 4) / cpg;
  float s = 0.f, q = 0.f;
  if (pl < prow) {
    const float K = gn_pivot<T>(x, batch_stride, n, g, cpg);
    const T *xn = x + (size_t)n * batch_stride + cq * 4;
    // 8 loads in flight per thread (the sums stay in pixel order): one load per iteration left the SM with ~16 KB in flight,
    // a third of what the HBM latency needs
    for (int pb = p0 + pl; pb < p1; pb += 8 * prow) {
      float4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int p = pb + k * prow;
        v[k] = p < p1 ? ld4<T>(xn + (size_t)p * C) : make_float4(K, K, K, K);       // K - K = 0: contributes nothing
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float a = v[k].x - K, b = v[k].y - K, c = v[k].z - K, d = v[k].w - K;
        s += (a + b) + (c + d);
        q += (a * a + b * b) + (c * c + d * d);
      }
    }
  }
  s_s[threadIdx.x] = s;
  s_q[threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.x < G) {                       // thread g adds its group's threads in a fixed order: quads of the group, then rows
    const int qpg = cpg / 4;
    double ts = 0.0, tq = 0.0;
    for (int r = 0; r < prow; ++r)
      for (int c = 0; c < qpg; ++c) {
        const int t = r * quads + threadIdx.x * qpg + c;
        ts += double(s_s[t]);
        tq += double(s_q[t]);
      }
    double *dst = partials + (((size_t)n * nchunks + chunk) * G + threadIdx.x) * 2;
    dst[0] = ts;
    dst[1] = tq;
  }
}

// one WARP per (sample, group): lane l adds chunks l, l + 32, ... in order, then the 32 lane sums are added in lane order by a
// fixed shuffle tree -- a fixed order (deterministic), 32 loads in flight instead of a 230-long dependent chain per thread
__global__ void __launch_bounds__(128) gn_finalize_kernel(const double *__restrict__ partials, int nchunks, int G, int total,
                                                          double *__restrict__ sums) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;   // (n, g)
  if (i >= total) return;                                                                  // whole warps leave together
  const int n = i / G, g = i - n * G;
  double ts = 0.0, tq = 0.0;
  for (int c = lane; c < nchunks; c += 32) {
    const double *src = partials + (((size_t)n * nchunks + c) * G + g) * 2;
    ts += src[0];
    tq += src[1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ts += __shfl_down_sync(0xffffffffu, ts, o);
    tq += __shfl_down_sync(0xffffffffu, tq, o);
  }
  if (lane == 0) {
    sums[(size_t)i * 2] = ts;
    sums[(size_t)i * 2 + 1] = tq;
  }
}

struct GnApplyParams {
  const void *x;
  int64_t x_batch_stride;
  const double *sums;
  const float *gamma, *beta;
  int N, HW, C, G;
  float eps;
  int relu;
  const float *up;           // optional low-res fp32 map (N, uh, uw, C) added after bilinear upsampling to (H, W)
  int64_t up_batch_stride;
  int uh, uw, H, W;
  const float *pos;          // optional (HW, C) fp32, for out_lp_pos
  float *out_f32;
  void *out_lp, *out_lp_pos;
  int64_t out_batch_stride;  // elements between batch items of every output
};

// grid (pixel chunks, N); a thread owns 4 fixed channels (its group's mean / rstd and gamma / beta stay in registers) and
// walks pixels with stride blockDim / (C/4)
template <typename T, typename TL, bool kUp>
__global__ void __launch_bounds__(256) gn_apply_kernel(const GnApplyParams p, int pix_per_cta) {
  const int quads = p.C / 4;
  const int cq = threadIdx.x % quads, pl = threadIdx.x / quads, prow = blockDim.x / quads;
  if (pl >= prow) return;
  const int n = blockIdx.y;
  const int c = cq * 4, g = c / (p.C / p.G);
  const double inv_cnt = 1.0 / (double(p.HW) * double(p.C / p.G));
  const double su = p.sums[((size_t)n * p.G + g) * 2] * inv_cnt, sq = p.sums[((size_t)n * p.G + g) * 2 + 1] * inv_cnt;
  const float mean = gn_pivot<T>(static_cast<const T *>(p.x), p.x_batch_stride, n, g, p.C / p.G) + float(su);   // sums are pivot-shifted
  const float rstd = rsqrtf(fmaxf(float(sq - su * su), 0.f) + p.eps);
  const float4 ga = *reinterpret_cast<const float4 *>(p.gamma + c), be = *reinterpret_cast<const float4 *>(p.beta + c);
  const float4 sc = make_float4(rstd * ga.x, rstd * ga.y, rstd * ga.z, rstd * ga.w);
  const float4 sh = make_float4(be.x - mean * sc.x, be.y - mean * sc.y, be.z - mean * sc.z, be.w - mean * sc.w);
  const T *xn = static_cast<const T *>(p.x) + (size_t)n * p.x_batch_stride + c;
  const float *u = kUp ? p.up + (size_t)n * p.up_batch_stride + c : nullptr;
  const float ry = float(p.uh) / float(max(p.H, 1)), rx = float(p.uw) / float(max(p.W, 1));
  const int p0 = blockIdx.x * pix_per_cta, p1 = min(p0 + pix_per_cta, p.HW);
  // 4 pixels per iteration; ALL their loads -- the map itself and, for the FPN's top-down add, the 4 bilinear taps of the low-res
  // map -- are issued before any of them is used: with one dependent load chain per pixel an SM kept ~16 KB in flight and the
  // kernel ran at 48 % of the HBM roofline (the taps are L2 round trips of ~1 us each)
  constexpr int KP = 4;
  for (int pb = p0 + pl; pb < p1; pb += KP * prow) {
    float4 vv[KP], ta[KP], tb[KP], tc[KP], td[KP];
    float wy[KP], wx[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      const int pix = min(pb + k * prow, p1 - 1);                 // tail: repeat the last pixel (not stored)
      vv[k] = ld4<T>(xn + (size_t)pix * p.C);
      if constexpr (kUp) {
        // F.interpolate(mode="bilinear", align_corners=False): src = (dst + 0.5) * in/out - 0.5, clamped at 0
        const int oy = pix / p.W, ox = pix - oy * p.W;
        const float fy = fmaxf((oy + 0.5f) * ry - 0.5f, 0.f), fx = fmaxf((ox + 0.5f) * rx - 0.5f, 0.f);
        const int y0 = min(int(fy), p.uh - 1), x0 = min(int(fx), p.uw - 1);
        const int y1 = min(y0 + 1, p.uh - 1), x1 = min(x0 + 1, p.uw - 1);
        wy[k] = fy - y0; wx[k] = fx - x0;
        ta[k] = *reinterpret_cast<const float4 *>(u + ((size_t)y0 * p.uw + x0) * p.C);
        tb[k] = *reinterpret_cast<const float4 *>(u + ((size_t)y0 * p.uw + x1) * p.C);
        tc[k] = *reinterpret_cast<const float4 *>(u + ((size_t)y1 * p.uw + x0) * p.C);
        td[k] = *reinterpret_cast<const float4 *>(u + ((size_t)y1 * p.uw + x1) * p.C);
      }
    }
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      const int pix = pb + k * prow;
      if (pix >= p1) break;
      const float4 v = vv[k];
      float4 y = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
      if constexpr (kUp) {
        const float ly = wy[k], lx = wx[k];
        const float4 a = ta[k], b = tb[k], cc = tc[k], d = td[k];
        const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
        y.x += w00 * a.x + w01 * b.x + w10 * cc.x + w11 * d.x;
        y.y += w00 * a.y + w01 * b.y + w10 * cc.y + w11 * d.y;
        y.z += w00 * a.z + w01 * b.z + w10 * cc.z + w11 * d.z;
        y.w += w00 * a.w + w01 * b.w + w10 * cc.w + w11 * d.w;
      }
      if (p.relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
      const size_t o = (size_t)n * p.out_batch_stride + (size_t)pix * p.C + c;
      if (p.out_f32) st4<float>(p.out_f32 + o, y);
      if (p.out_lp) st4<TL>(static_cast<TL *>(p.out_lp) + o, y);
      if (p.out_lp_pos) {
        const float4 q = *reinterpret_cast<const float4 *>(p.pos + (size_t)pix * p.C + c);
        st4<TL>(static_cast<TL *>(p.out_lp_pos) + o, make_float4(y.x + q.x, y.y + q.y, y.z + q.z, y.w + q.w));
      }
    }
  }
}

}  // namespace
}  // namespace dvis

using namespace dvis;

extern "C" int dvis_groupnorm_nhwc(const void *x, int x_dtype, int64_t x_batch_stride, int N, int HW, int C, int G,
                                   const float *gamma, const float *beta, float eps, int relu, double *sums_workspace,
                                   const float *up, int64_t up_batch_stride, int up_h, int up_w, int H, int W,
                                   const float *pos, float *out_f32, void *out_lp, void *out_lp_pos, int lp_dtype,
                                   int64_t out_batch_stride, void *stream) {
  DVIS_REQUIRE(x && gamma && beta && sums_workspace, "groupnorm_nhwc: null pointer argument");
  DVIS_REQUIRE(N > 0 && HW > 0 && C > 0 && G > 0 && C % G == 0, "groupnorm_nhwc: bad sizes");
  DVIS_REQUIRE(C % 4 == 0 && (C / G) % 4 == 0 && C <= 1024 && G <= kMaxGroups,
               "groupnorm_nhwc: need C %% 4 == 0, (C/G) %% 4 == 0, C <= 1024, G <= 64 (C=%d G=%d)", C, G);
  DVIS_REQUIRE(out_f32 || out_lp || out_lp_pos, "groupnorm_nhwc: no output requested");
  DVIS_REQUIRE(!out_lp_pos || pos, "groupnorm_nhwc: out_lp_pos needs pos");
  DVIS_REQUIRE(!up || (up_h > 0 && up_w > 0 && H * W == HW), "groupnorm_nhwc: upsample-add needs H*W == HW");
  if ((x_dtype != DVIS_F32 && x_dtype != DVIS_BF16) || (lp_dtype != DVIS_F32 && lp_dtype != DVIS_BF16))
    return fail(DVIS_ERR_UNSUPPORTED, "groupnorm_nhwc: dtypes must be f32 or bf16");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nchunks = (HW + kGnChunk - 1) / kGnChunk;
  double *partials = sums_workspace + (size_t)2 * N * G;      // workspace layout: [N*G*2 sums][N*nchunks*G*2 partials]
  dim3 grid(nchunks, N);
  prefer_carveout(gn_partial_kernel<__nv_bfloat16>);
  if (x_dtype == DVIS_F32)
    gn_partial_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float *>(x), x_batch_stride, HW, C, G, partials);
  else
    gn_partial_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16 *>(x), x_batch_stride, HW, C, G, partials);
  if (int rc = check_launch("gn_partial_kernel")) return rc;
  gn_finalize_kernel<<<(N * G + 3) / 4, 128, 0, s>>>(partials, nchunks, G, N * G, sums_workspace);   // 4 warps = 4 (sample, group) pairs per CTA
  if (int rc = check_launch("gn_finalize_kernel")) return rc;
  GnApplyParams p{x, x_batch_stride, sums_workspace, gamma, beta, N, HW, C, G, eps, relu, up, up_batch_stride, up_h, up_w,
                  H, W, pos, out_f32, out_lp, out_lp_pos, out_batch_stride};
  const int apply_pix = 64;                                     // 16 pixels per thread row at C = 256
  dim3 agrid((HW + apply_pix - 1) / apply_pix, N);
  using bf = __nv_bfloat16;
#define DVIS_GN_APPLY(TX, TLP)                                                               \
  do {                                                                                       \
    if (up) gn_apply_kernel<TX, TLP, true><<<agrid, 256, 0, s>>>(p, apply_pix);              \
    else gn_apply_kernel<TX, TLP, false><<<agrid, 256, 0, s>>>(p, apply_pix);                \
  } while (0)
  prefer_carveout(gn_apply_kernel<bf, bf, false>);
  if (x_dtype == DVIS_F32 && lp_dtype == DVIS_F32) DVIS_GN_APPLY(float, float);
  else if (x_dtype == DVIS_F32) DVIS_GN_APPLY(float, bf);
  else if (lp_dtype == DVIS_F32) DVIS_GN_APPLY(bf, float);
  else DVIS_GN_APPLY(bf, bf);
#undef DVIS_GN_APPLY
  return check_launch("gn_apply_kernel");
}
