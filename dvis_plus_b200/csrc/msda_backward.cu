// Multi-scale deformable attention backward for sm_100a.
//
// Replaces ms_deform_attn_cuda_backward (OPS/src/cuda/ms_deform_attn_cuda.cu:88-158) and the col2im kernel family
// (OPS/src/cuda/ms_deform_im2col_cuda.cuh:306-925, dispatch :961-1331).  Arithmetic follows
// ms_deform_attn_col2im_bilinear (cuh:92-164): grad_value += corner_weight * attn * grad_out (scatter),
// grad_attn = <grad_out, sample>, grad_loc = (W * <gw, grad_out*attn>, H * <gh, grad_out*attn>).
//
// Instead of the reference's seven kernel variants selected by `channels`, one vectorised kernel: a group of
// LPR = D/4 lanes owns a (query, head) item, the per-point reductions over channels are warp shuffles inside the
// group (no shared memory, no serial thread-0 loop), and grad_value is scattered with 16-byte vector atomics
// (red.global.add.v4.f32, sm_90+).  grad_loc / grad_attn of an item are written by exactly one group, so they need
// no atomics and are deterministic; only grad_value accumulation order is not (as in the reference).
// A scalar kernel covers double precision and head dims that are not 8/16/32/64/128.
#include <algorithm>

#include "common.cuh"

namespace dvis {
namespace {

constexpr int kMaxLevels = 8;

struct BwdParams {
  const void *value, *loc, *attn, *gout;
  const int64_t *shapes, *level_start;
  void *gvalue, *gloc, *gattn;
  int N, S, M, L, Lq, P;
};

template <int D>
__global__ void __launch_bounds__(256) msda_bwd_vec_kernel(const BwdParams p) {
  constexpr int LPR = D / 4, G = 32 / LPR;
  __shared__ int sH[kMaxLevels], sW[kMaxLevels], sStart[kMaxLevels];
  if (threadIdx.x < p.L) {
    sH[threadIdx.x] = int(p.shapes[2 * threadIdx.x]);
    sW[threadIdx.x] = int(p.shapes[2 * threadIdx.x + 1]);
    sStart[threadIdx.x] = int(p.level_start[threadIdx.x]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / LPR, j = lane % LPR;
  const int M = p.M, L = p.L, P = p.P;
  const long total = (long)p.N * p.Lq * M;
  const long item = ((long)blockIdx.x * 8 + warp) * G + g;
  const bool active = item < total;
  const long it = active ? item : 0;                 // inactive groups still take part in the shuffles
  const int m = int(it % M);
  const long n = it / M / p.Lq;
  const int row = M * D;
  const float *value_n = static_cast<const float *>(p.value) + (size_t)n * p.S * row;
  float *gvalue_n = static_cast<float *>(p.gvalue) + (size_t)n * p.S * row;
  const float4 go = active ? *reinterpret_cast<const float4 *>(static_cast<const float *>(p.gout) + it * D + j * 4)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
  const float *loc = static_cast<const float *>(p.loc) + it * (size_t)(L * P) * 2;
  const float *att = static_cast<const float *>(p.attn) + it * (size_t)(L * P);
  float *gloc = static_cast<float *>(p.gloc) + it * (size_t)(L * P) * 2;
  float *gatt = static_cast<float *>(p.gattn) + it * (size_t)(L * P);
  for (int l = 0; l < L; ++l) {
    const int H = sH[l], W = sW[l];
    const size_t lvl = (size_t)sStart[l] * row + m * D + j * 4;
    for (int pt = 0; pt < P; ++pt) {
      const int i = l * P + pt;
      const float x = __ldg(loc + 2 * i), y = __ldg(loc + 2 * i + 1), a = __ldg(att + i);
      const float h_im = y * float(H) - 0.5f, w_im = x * float(W) - 0.5f;
      const bool inr = active && h_im > -1.f && w_im > -1.f && h_im < float(H) && w_im < float(W);
      const float hf = floorf(h_im), wf = floorf(w_im);
      const int h0 = inr ? int(hf) : 0, w0 = inr ? int(wf) : 0;
      const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
      const bool top = inr && h0 >= 0, bot = inr && h0 + 1 <= H - 1, lef = w0 >= 0, rig = w0 + 1 <= W - 1;
      const size_t o00 = lvl + (size_t)((long)h0 * W + w0) * row;
      const size_t offs[4] = {o00, o00 + row, o00 + (size_t)W * row, o00 + (size_t)W * row + row};
      const bool ok[4] = {top && lef, top && rig, bot && lef, bot && rig};
      const float cw[4] = {hh * hw, hh * lw, lh * hw, lh * lw};
      const float ch[4] = {-hw, -lw, hw, lw};        // d(sample)/d(h) coefficient of each corner (cuh:126-155)
      const float cx[4] = {-hh, hh, -lh, lh};        // d(sample)/d(w)
      float ga = 0.f, gh = 0.f, gw = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (ok[c]) {
          const float4 v = __ldg(reinterpret_cast<const float4 *>(value_n + offs[c]));
          const float dot = v.x * go.x + v.y * go.y + v.z * go.z + v.w * go.w;
          ga += cw[c] * dot;
          gh += ch[c] * dot;
          gw += cx[c] * dot;
          const float s = cw[c] * a;
          atomicAdd(reinterpret_cast<float4 *>(gvalue_n + offs[c]), make_float4(s * go.x, s * go.y, s * go.z, s * go.w));
        }
      }
#pragma unroll
      for (int o = LPR / 2; o; o >>= 1) {
        ga += __shfl_xor_sync(0xffffffffu, ga, o);
        gh += __shfl_xor_sync(0xffffffffu, gh, o);
        gw += __shfl_xor_sync(0xffffffffu, gw, o);
      }
      if (inr && j == 0) {                            // out-of-range points keep the caller's zeros (cu:127-128)
        gatt[i] = ga;
        gloc[2 * i] = float(W) * gw * a;
        gloc[2 * i + 1] = float(H) * gh * a;
      }
    }
  }
}

// scalar kernel: any D, float / double; one thread per (n, q, m, l, p), channel loop inside, scalar atomics
template <typename T>
__global__ void __launch_bounds__(256) msda_bwd_generic_kernel(const BwdParams p, int D, long total) {
  const T *value = static_cast<const T *>(p.value), *locs = static_cast<const T *>(p.loc), *attn = static_cast<const T *>(p.attn);
  const T *gout = static_cast<const T *>(p.gout);
  T *gvalue = static_cast<T *>(p.gvalue), *gloc = static_cast<T *>(p.gloc), *gattn = static_cast<T *>(p.gattn);
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int pt = int(idx % p.P);
    const int l = int((idx / p.P) % p.L);
    const long qm = idx / p.P / p.L;                 // (n*Lq + q)*M + m
    const int m = int(qm % p.M);
    const long n = qm / p.M / p.Lq;
    const int H = int(p.shapes[2 * l]), W = int(p.shapes[2 * l + 1]);
    const size_t row = (size_t)p.M * D;
    const size_t lvl = ((size_t)n * p.S + (size_t)p.level_start[l]) * row + (size_t)m * D;
    const T x = locs[2 * idx], y = locs[2 * idx + 1], a = attn[idx];
    const T h_im = y * T(H) - T(0.5), w_im = x * T(W) - T(0.5);
    if (!(h_im > T(-1) && w_im > T(-1) && h_im < T(H) && w_im < T(W))) continue;
    const T hf = floor(h_im), wf = floor(w_im);
    const int h0 = int(hf), w0 = int(wf), h1 = h0 + 1, w1 = w0 + 1;
    const T lh = h_im - hf, lw = w_im - wf, hh = T(1) - lh, hw = T(1) - lw;
    const bool ok[4] = {h0 >= 0 && w0 >= 0, h0 >= 0 && w1 <= W - 1, h1 <= H - 1 && w0 >= 0, h1 <= H - 1 && w1 <= W - 1};
    const size_t offs[4] = {lvl + ((size_t)h0 * W + w0) * row, lvl + ((size_t)h0 * W + w1) * row,
                            lvl + ((size_t)h1 * W + w0) * row, lvl + ((size_t)h1 * W + w1) * row};
    const T cw[4] = {hh * hw, hh * lw, lh * hw, lh * lw}, ch[4] = {-hw, -lw, hw, lw}, cx[4] = {-hh, hh, -lh, lh};
    T ga = 0, gh = 0, gw = 0;
    for (int c = 0; c < D; ++c) {
      const T tg = gout[qm * D + c];
      for (int k = 0; k < 4; ++k) {
        if (!ok[k]) continue;
        const T v = value[offs[k] + c];
        ga += cw[k] * v * tg;
        gh += ch[k] * v * tg;
        gw += cx[k] * v * tg;
        atomicAdd(gvalue + offs[k] + c, cw[k] * a * tg);
      }
    }
    gattn[idx] = ga;
    gloc[2 * idx] = T(W) * gw * a;
    gloc[2 * idx + 1] = T(H) * gh * a;
  }
}

}  // namespace
}  // namespace dvis

using namespace dvis;

extern "C" int dvis_msda_backward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start,
                                  const void *sampling_loc, const void *attn_weight, const void *grad_out, int batch,
                                  int spatial_size, int num_heads, int channels, int num_levels, int num_query,
                                  int num_point, int dtype, void *grad_value, void *grad_sampling_loc,
                                  void *grad_attn_weight, void *stream) {
  DVIS_REQUIRE(value && spatial_shapes && level_start && sampling_loc && attn_weight && grad_out && grad_value &&
                   grad_sampling_loc && grad_attn_weight, "msda_backward: null pointer argument");
  DVIS_REQUIRE(batch > 0 && spatial_size > 0 && num_heads > 0 && channels > 0 && num_levels > 0 && num_query > 0 && num_point > 0,
               "msda_backward: sizes must be positive");
  DVIS_REQUIRE(num_levels <= kMaxLevels, "msda_backward: num_levels %d > %d", num_levels, kMaxLevels);
  BwdParams p{value, sampling_loc, attn_weight, grad_out, spatial_shapes, level_start, grad_value, grad_sampling_loc,
              grad_attn_weight, batch, spatial_size, num_heads, num_levels, num_query, num_point};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long items = (long)batch * num_query * num_heads;
  if (dtype == DVIS_F32 && aligned16(value) && aligned16(grad_value) && aligned16(grad_out)) {
#define DVIS_BWD(DD)                                                                          \
  case DD: {                                                                                  \
    const int per_block = 8 * (32 / (DD / 4));                                                \
    msda_bwd_vec_kernel<DD><<<unsigned((items + per_block - 1) / per_block), 256, 0, s>>>(p); \
    return check_launch("msda_bwd_vec_kernel");                                               \
  }
    switch (channels) {
      DVIS_BWD(8) DVIS_BWD(16) DVIS_BWD(32) DVIS_BWD(64) DVIS_BWD(128)
      default: break;
    }
#undef DVIS_BWD
  }
  const long total = items * num_levels * num_point;
  const int blocks = int(std::min<long>((total + 255) / 256, (long)kNumSMs * 32));
  if (dtype == DVIS_F32) msda_bwd_generic_kernel<float><<<blocks, 256, 0, s>>>(p, channels, total);
  else if (dtype == DVIS_F64) msda_bwd_generic_kernel<double><<<blocks, 256, 0, s>>>(p, channels, total);
  else return fail(DVIS_ERR_UNSUPPORTED, "msda_backward: dtype %d (float/double only, like the reference)", dtype);
  return check_launch("msda_bwd_generic_kernel");
}
