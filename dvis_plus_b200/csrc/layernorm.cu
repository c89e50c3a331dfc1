// Fused residual-add + LayerNorm (+ optional positional add) for the post-norm blocks on the path:
//   y = LayerNorm(x + residual) * gamma + beta
// as in MSDeformAttnTransformerEncoderLayer.forward (P/mask2former/modeling/pixel_decoder/msdeformattn.py:125-126,
// 118-119) and SelfAttentionLayer / CrossAttentionLayer / FFNLayer.forward_post
// (P/mask2former_video/modeling/transformer_decoder/video_mask2former_transformer_decoder.py:48-49,109-110,165-166).
// One pass over HBM: reads x (+ residual, + pos), writes the fp32 stream and/or a low-precision copy for the next
// GEMM, and optionally the low-precision (y + pos) the next deformable-attention query projection consumes
// (with_pos_embed, msdeformattn.py:112-114,124) -- instead of add, layer_norm, cast, add, cast as five kernels.
//
// One warp per row; each lane keeps C/32 values in registers, statistics by warp shuffles, two-pass variance.
#include "common.cuh"

namespace dvis {
namespace {

template <typename T>
__device__ __forceinline__ float4 load4(const T *p);
template <>
__device__ __forceinline__ float4 load4<float>(const float *p) { return *reinterpret_cast<const float4 *>(p); }
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16 *p) {
  const uint2 u = *reinterpret_cast<const uint2 *>(p);
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                     __uint_as_float(u.y & 0xffff0000u));
}
template <typename T>
__device__ __forceinline__ void store4(T *p, float4 v);
template <>
__device__ __forceinline__ void store4<float>(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16 *p, float4 v) {
  *reinterpret_cast<uint2 *>(p) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

struct LnParams {
  const void *x;
  const void *residual;
  const float *gamma, *beta;
  const float *pos;
  int64_t rows, pos_rows;
  int C;
  float eps;
  float *out_f32;
  void *out_lp, *out_lp_pos;
};

// VPL = float4 vectors per lane (C <= 128 * VPL, C % 4 == 0; vectors past C are predicated off)
template <typename TX, typename TR, typename TL, int VPL>
__global__ void __launch_bounds__(256) add_layernorm_kernel(const LnParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const int C = p.C;
  const TX *x = static_cast<const TX *>(p.x) + row * C;
  const TR *r = p.residual ? static_cast<const TR *>(p.residual) + row * C : nullptr;
  float4 v[VPL];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C) {
      v[i] = load4<TX>(x + c);
      if (r) {
        const float4 t = load4<TR>(r + c);
        v[i].x += t.x; v[i].y += t.y; v[i].z += t.z; v[i].w += t.w;
      }
    }
    sum += v[i].x + v[i].y + v[i].z + v[i].w;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / float(C);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    if ((i * 32 + lane) * 4 < C) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      sq += a * a + b * b + c * c + d * d;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / float(C) + p.eps);
  const float *pos = p.pos ? p.pos + (row % p.pos_rows) * C : nullptr;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c >= C) continue;
    const float4 g = *reinterpret_cast<const float4 *>(p.gamma + c), b = *reinterpret_cast<const float4 *>(p.beta + c);
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x + b.x;
    y.y = (v[i].y - mean) * rstd * g.y + b.y;
    y.z = (v[i].z - mean) * rstd * g.z + b.z;
    y.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (p.out_f32) store4<float>(p.out_f32 + row * C + c, y);
    if (p.out_lp) store4<TL>(static_cast<TL *>(p.out_lp) + row * C + c, y);
    if (p.out_lp_pos) {
      const float4 q = *reinterpret_cast<const float4 *>(pos + c);
      store4<TL>(static_cast<TL *>(p.out_lp_pos) + row * C + c, make_float4(y.x + q.x, y.y + q.y, y.z + q.z, y.w + q.w));
    }
  }
}

template <typename TX, typename TR, typename TL>
int launch_ln(const LnParams &p, cudaStream_t s) {
  const int warps = 8;
  const unsigned grid = unsigned((p.rows + warps - 1) / warps);
  int vpl = (p.C + 127) / 128;
  if (vpl == 5) vpl = 6;                       // built: 1, 2, 3, 4, 6, 8, 16 vectors per lane (extra vectors are predicated off)
  else if (vpl == 7) vpl = 8;
  else if (vpl > 8) vpl = 16;
  switch (vpl) {
#define DVIS_LN(V) case V: prefer_carveout(add_layernorm_kernel<TX, TR, TL, V>); add_layernorm_kernel<TX, TR, TL, V><<<grid, warps * 32, 0, s>>>(p); break;
    DVIS_LN(1) DVIS_LN(2) DVIS_LN(3) DVIS_LN(4) DVIS_LN(6) DVIS_LN(8) DVIS_LN(16)
#undef DVIS_LN
    default: return fail(DVIS_ERR_UNSUPPORTED, "add_layernorm: C=%d", p.C);
  }
  return check_launch("add_layernorm_kernel");
}

}  // namespace
}  // namespace dvis

using namespace dvis;

extern "C" int dvis_add_layernorm(const void *x, int x_dtype, const void *residual, int residual_dtype,
                                  const float *gamma, const float *beta, const float *pos, int64_t pos_rows, int64_t rows,
                                  int C, float eps, float *out_f32, void *out_lp, void *out_lp_pos, int lp_dtype,
                                  void *stream) {
  DVIS_REQUIRE(x && gamma && beta, "add_layernorm: null pointer argument");
  DVIS_REQUIRE(rows > 0 && C > 0 && C % 4 == 0 && C <= 2048, "add_layernorm: rows > 0, C %% 4 == 0 and C <= 2048 required (C=%d)", C);
  DVIS_REQUIRE(out_f32 || out_lp || out_lp_pos, "add_layernorm: no output requested");
  DVIS_REQUIRE(!out_lp_pos || (pos && pos_rows > 0), "add_layernorm: out_lp_pos needs pos");
  DVIS_REQUIRE((rows + 7) / 8 < (int64_t(1) << 31), "add_layernorm: too many rows");
  LnParams p{x, residual, gamma, beta, pos, rows, pos_rows > 0 ? pos_rows : 1, C, eps, out_f32, out_lp, out_lp_pos};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool xf = x_dtype == DVIS_F32, rf = residual_dtype == DVIS_F32 || !residual, lf = lp_dtype == DVIS_F32;
  if ((x_dtype != DVIS_F32 && x_dtype != DVIS_BF16) || (residual && residual_dtype != DVIS_F32 && residual_dtype != DVIS_BF16) ||
      (lp_dtype != DVIS_F32 && lp_dtype != DVIS_BF16))
    return fail(DVIS_ERR_UNSUPPORTED, "add_layernorm: dtypes must be f32 or bf16");
  using bf = __nv_bfloat16;
  if (xf && rf && lf) return launch_ln<float, float, float>(p, s);
  if (xf && rf && !lf) return launch_ln<float, float, bf>(p, s);
  if (xf && !rf && lf) return launch_ln<float, bf, float>(p, s);
  if (xf && !rf && !lf) return launch_ln<float, bf, bf>(p, s);
  if (!xf && rf && lf) return launch_ln<bf, float, float>(p, s);
  if (!xf && rf && !lf) return launch_ln<bf, float, bf>(p, s);
  if (!xf && !rf && lf) return launch_ln<bf, bf, float>(p, s);
  return launch_ln<bf, bf, bf>(p, s);
}
