// Small fused helpers around the mask head of the masked-attention decoder
// (P/dvis_Plus/video_mask2former_transformer_decoder.py:358-374, :297):
//   * bilinear resize of a channels-last bf16 map (F.interpolate(mode="bilinear", align_corners=False), py:367) used to
//     bring mask_features to each attention level ONCE (interpolate(E @ F) == E @ interpolate(F));
//   * attention bias from low-resolution mask logits: sigmoid(x) < 0.5 <=> x < 0 -> -inf (may not attend), rows that would
//     be fully masked attend everywhere (py:297) -- replaces compare / all-reduce / and / fill / masked_fill kernels.
#include <algorithm>

#include "common.cuh"

namespace dvis {
namespace {

__device__ __forceinline__ float4 ld_bf4(const __nv_bfloat16 *p) {
  const uint2 u = *reinterpret_cast<const uint2 *>(p);
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                     __uint_as_float(u.y & 0xffff0000u));
}

// in (N, h, w, C) bf16 -> out (N, H, W, C) bf16; one thread per 4 channels of an output pixel
__global__ void __launch_bounds__(256) resize_bilinear_nhwc_kernel(const __nv_bfloat16 *__restrict__ in, __nv_bfloat16 *__restrict__ out,
                                                                   int N, int h, int w, int H, int W, int C) {
  const int quads = C / 4;
  const int64_t total = (int64_t)N * H * W * quads;
  const float sy = float(h) / float(H), sx = float(w) / float(W);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cq = int(i % quads);
    int64_t r = i / quads;
    const int ox = int(r % W); r /= W;
    const int oy = int(r % H);
    const int n = int(r / H);
    const float fy = fmaxf((oy + 0.5f) * sy - 0.5f, 0.f), fx = fmaxf((ox + 0.5f) * sx - 0.5f, 0.f);
    const int y0 = min(int(fy), h - 1), x0 = min(int(fx), w - 1);
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = fy - y0, lx = fx - x0;
    const __nv_bfloat16 *b = in + (size_t)n * h * w * C + cq * 4;
    const float4 a = ld_bf4(b + ((size_t)y0 * w + x0) * C), bb = ld_bf4(b + ((size_t)y0 * w + x1) * C);
    const float4 c = ld_bf4(b + ((size_t)y1 * w + x0) * C), d = ld_bf4(b + ((size_t)y1 * w + x1) * C);
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    const float4 y = make_float4(w00 * a.x + w01 * bb.x + w10 * c.x + w11 * d.x, w00 * a.y + w01 * bb.y + w10 * c.y + w11 * d.y,
                                 w00 * a.z + w01 * bb.z + w10 * c.z + w11 * d.z, w00 * a.w + w01 * bb.w + w10 * c.w + w11 * d.w);
    *reinterpret_cast<uint2 *>(out + (size_t)i * 4) = make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
  }
}

// logits (rows, hw) f32 -> bias (rows, hw) TB: 0 where the query may attend, -inf where not; one warp per row
template <typename TB>
__global__ void __launch_bounds__(256) attn_bias_kernel(const float *__restrict__ logits, TB *__restrict__ bias, int64_t rows, int hw) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float *l = logits + row * hw;
  bool any_open = false;                                  // some position with sigmoid >= 0.5
  for (int i = lane; i < hw; i += 32) any_open |= !(l[i] < 0.f);
  any_open = __any_sync(0xffffffffu, any_open);
  TB *b = bias + row * hw;
  const TB ninf = TB(-INFINITY), zero = TB(0.f);
  for (int i = lane; i < hw; i += 32) b[i] = (any_open && l[i] < 0.f) ? ninf : zero;
}

// The decoder's memory of one attention level (py:270-279: src = input_proj(x) + level_embed, key input = src + pos):
//   tok[b, p, :] = x[b, p, :] + level[:],   key[b, p, :] = tok[b, p, :] + pos[p, :]      both written in bf16
// x: fp32 or bf16 rows of C channels (pixel stride C, batch stride free -- a slice of the encoder's (N, S, C) token buffer).
// Replaces float() / add / add / cast / cast passes over up to 241 MB per level (0.46 ms of aten adds per clip).
template <typename TX>
__global__ void __launch_bounds__(256) level_tokens_kernel(const TX *__restrict__ x, int64_t x_batch, const float *__restrict__ level,
                                                           const float *__restrict__ pos, __nv_bfloat16 *__restrict__ tok,
                                                           __nv_bfloat16 *__restrict__ key, int B, int HW, int C) {
  const int quads = C / 4;
  const int64_t total = (int64_t)B * HW * quads;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cq = int(i % quads);
    const int64_t r = i / quads;
    const int pix = int(r % HW), b = int(r / HW);
    const TX *src = x + (size_t)b * x_batch + (size_t)pix * C + cq * 4;
    float4 v;
    if constexpr (sizeof(TX) == 4) v = *reinterpret_cast<const float4 *>(src);
    else v = ld_bf4(reinterpret_cast<const __nv_bfloat16 *>(src));
    const float4 le = *reinterpret_cast<const float4 *>(level + cq * 4);
    v = make_float4(v.x + le.x, v.y + le.y, v.z + le.z, v.w + le.w);
    *reinterpret_cast<uint2 *>(tok + (size_t)i * 4) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    if (key) {
      const float4 q = *reinterpret_cast<const float4 *>(pos + (size_t)pix * C + cq * 4);
      *reinterpret_cast<uint2 *>(key + (size_t)i * 4) = make_uint2(pack_bf16x2(v.x + q.x, v.y + q.y), pack_bf16x2(v.z + q.z, v.w + q.w));
    }
  }
}

}  // namespace
}  // namespace dvis

using namespace dvis;

extern "C" int dvis_level_tokens(const void *x, int x_dtype, int64_t x_batch_stride, const float *level_embed, const float *pos, int B,
                                 int HW, int C, void *tok, void *key, void *stream) {
  DVIS_REQUIRE(x && level_embed && tok && (!key || pos), "level_tokens: null pointer argument");
  DVIS_REQUIRE(B > 0 && HW > 0 && C > 0 && C % 4 == 0, "level_tokens: bad sizes (C %% 4 == 0)");
  DVIS_REQUIRE(aligned16(x) && aligned16(tok) && aligned16(level_embed) && (!key || (aligned16(key) && aligned16(pos))) &&
                   x_batch_stride % 4 == 0, "level_tokens: 16-byte aligned rows required");
  const int64_t total = (int64_t)B * HW * (C / 4);
  const int blocks = int(std::min<int64_t>((total + 255) / 256, (int64_t)kNumSMs * 32));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto *t = static_cast<__nv_bfloat16 *>(tok), *k = static_cast<__nv_bfloat16 *>(key);
  prefer_carveout(level_tokens_kernel<__nv_bfloat16>);
  if (x_dtype == DVIS_F32)
    level_tokens_kernel<float><<<blocks, 256, 0, s>>>(static_cast<const float *>(x), x_batch_stride, level_embed, pos, t, k, B, HW, C);
  else if (x_dtype == DVIS_BF16)
    level_tokens_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(static_cast<const __nv_bfloat16 *>(x), x_batch_stride, level_embed, pos, t, k, B, HW, C);
  else return fail(DVIS_ERR_UNSUPPORTED, "level_tokens: x must be f32 or bf16");
  return check_launch("level_tokens_kernel");
}

extern "C" int dvis_resize_bilinear_nhwc(const void *in, int N, int h, int w, int C, void *out, int H, int W, void *stream) {
  DVIS_REQUIRE(in && out, "resize_bilinear_nhwc: null pointer argument");
  DVIS_REQUIRE(N > 0 && h > 0 && w > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "resize_bilinear_nhwc: bad sizes (C %% 4 == 0)");
  const int64_t total = (int64_t)N * H * W * (C / 4);
  const int blocks = int(std::min<int64_t>((total + 255) / 256, (int64_t)kNumSMs * 32));
  prefer_carveout(resize_bilinear_nhwc_kernel);
  resize_bilinear_nhwc_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16 *>(in), static_cast<__nv_bfloat16 *>(out), N, h, w, H, W, C);
  return check_launch("resize_bilinear_nhwc_kernel");
}

extern "C" int dvis_attn_bias_from_logits(const float *logits, int64_t rows, int hw, void *bias, int bias_dtype, void *stream) {
  DVIS_REQUIRE(logits && bias, "attn_bias_from_logits: null pointer argument");
  DVIS_REQUIRE(rows > 0 && hw > 0, "attn_bias_from_logits: sizes must be positive");
  const unsigned grid = unsigned((rows + 7) / 8);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (bias_dtype == DVIS_F32) attn_bias_kernel<float><<<grid, 256, 0, s>>>(logits, static_cast<float *>(bias), rows, hw);
  else if (bias_dtype == DVIS_BF16) attn_bias_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(logits, static_cast<__nv_bfloat16 *>(bias), rows, hw);
  else return fail(DVIS_ERR_UNSUPPORTED, "attn_bias_from_logits: bias dtype must be f32 or bf16");
  return check_launch("attn_bias_kernel");
}
