// Pair-packed bf16 variant of the fused multi-scale deformable attention forward (inference path of the module
// drop-in, OPS/modules/ms_deform_attn.py:98-118; sampling semantics of OPS/src/cuda/ms_deform_im2col_cuda.cuh:38-89,
// 242-304 unchanged).
//
// Why: the staged kernel (msda_forward.cu) is bound by the L1 data pipe -- every bilinear corner row is its own
// 128-byte line, 4 lines per sampling point.  Here `value` is re-laid out so that the two x-adjacent corners share
// one line:    pairs[n, e, m, 0:2, 0:D]  =  [ value[n, e-1, m, :], value[n, e, m, :] ]     (bf16, D = 32 -> 128 bytes)
// with value[-1] = value[S] = 0, e in [0, S].  A sampling point then needs 2 lines (top pair, bottom pair) instead of 4,
// the per-point parameters shrink to 16 bytes (2 offsets + 4 half weights: ONE LDS.128) and a group of 8 lanes gathers a
// whole pair with one LDG.128: lanes 0-3 carry the left corner's channels, lanes 4-7 the right corner's, combined by
// one xor-4 shuffle per channel at the end.  Costs one extra pass that writes the packed copy (2x the value bytes).
#include <algorithm>
#include <cuda_fp16.h>

#include "common.cuh"

namespace dvis {
namespace {

constexpr int kMaxLevels = 8;
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kD = 32;            // channels per head (bf16): one pair = 2 * 32 * 2 B = 128 B
constexpr int kItems = 64;        // (query, head) items per CTA
constexpr int kMaxLP = 32;

struct PairParams {
  const void *pairs;              // (N, S+1, M, 2, D) bf16
  const int64_t *shapes, *level_start;
  const void *offsets, *logits;   // TP
  const float *ref;
  int64_t off_stride, logit_stride;
  int ref_dim;
  const int32_t *order;
  void *out;                      // (N, Lq, M*D) bf16
  int N, S, M, L, Lq, P;
  int m_shift, p_shift, lpc_shift;
};

struct __align__(16) PointRec {
  uint32_t off_top, off_bot;      // byte offsets of the top / bottom pair inside this batch element's pair buffer
  __half2 w_top, w_bot;           // (left, right) weights, already multiplied by the attention weight
};

template <typename TP>
__device__ __forceinline__ float ldp(const TP *p);
template <>
__device__ __forceinline__ float ldp<float>(const float *p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ldp<__nv_bfloat16>(const __nv_bfloat16 *p) { return __bfloat162float(__ldg(p)); }

__global__ void __launch_bounds__(256) pack_pairs_kernel(const uint4 *__restrict__ value, uint4 *__restrict__ pairs, int S, int M,
                                                         int64_t total) {
  // one thread per 16-byte chunk of the output; a (e, m) entry is 8 chunks: 0-3 = value[e-1, m], 4-7 = value[e, m]
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = int(i & 7);
    const int64_t em = i >> 3;                    // (n*(S+1) + e)*M + m
    const int m = int(em % M);
    const int64_t ne = em / M;
    const int e = int(ne % (S + 1));
    const int64_t n = ne / (S + 1);
    const int s = e - 1 + (c >> 2);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (s >= 0 && s < S) v = __ldg(value + ((n * S + s) * M + m) * 4 + (c & 3));
    pairs[i] = v;
  }
}

template <typename TP, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) msda_pair_kernel(const PairParams p) {
  extern __shared__ uint4 dyn_smem[];
  __shared__ int s_item[kItems];
  __shared__ int sH[kMaxLevels], sW[kMaxLevels], sStart[kMaxLevels];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < p.L) {
    sH[tid] = int(p.shapes[2 * tid]);
    sW[tid] = int(p.shapes[2 * tid + 1]);
    sStart[tid] = int(p.level_start[tid]);
  }
  const int n = blockIdx.y, M = p.M, P = p.P, LP = p.L * p.P, LPs = LP | 1;
  PointRec *s_rec = reinterpret_cast<PointRec *>(dyn_smem);                   // [kItems][LPs]
  float *s_prob = reinterpret_cast<float *>(dyn_smem + kItems * LPs);        // [kItems][LP]
  const int per_batch = p.Lq * M;
  const int chunk_begin = blockIdx.x * kItems;
  const int nitems = min(kItems, per_batch - chunk_begin);
  if (tid < nitems) s_item[tid] = p.order ? __ldg(p.order + chunk_begin + tid) : chunk_begin + tid;
  __syncthreads();

  // phase 0: softmax over the L*P logits of each item (4 threads per item)
  {
    const int il = tid >> 2, sub = tid & 3;
    const bool act = il < nitems;
    const int item = act ? s_item[il] : 0;
    const int q = p.m_shift >= 0 ? item >> p.m_shift : item / M, m = item - q * M;
    const TP *lg = static_cast<const TP *>(p.logits) + ((size_t)n * p.Lq + q) * p.logit_stride + (size_t)m * LP;
    float mx = -INFINITY;
    for (int i = sub; i < LP; i += 4) mx = fmaxf(mx, act ? ldp<TP>(lg + i) : 0.f);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float sum = 0.f;
    for (int i = sub; i < LP; i += 4) {
      const float e = act ? __expf(ldp<TP>(lg + i) - mx) : 0.f;
      if (act) s_prob[il * LP + i] = e;
      sum += e;
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = 1.f / sum;
    if (act)
      for (int i = sub; i < LP; i += 4) s_prob[il * LP + i] *= inv;
  }
  __syncthreads();

  // phase 1: one thread per (item, point)
  {
    const int pt = tid & ((1 << p.lpc_shift) - 1);
    if (pt < LP) {
      const int l = p.p_shift >= 0 ? pt >> p.p_shift : pt / P;
      const int H = sH[l], W = sW[l], start = sStart[l];
      for (int il = tid >> p.lpc_shift; il < nitems; il += kThreads >> p.lpc_shift) {
        const int item = s_item[il];
        const int q = p.m_shift >= 0 ? item >> p.m_shift : item / M, m = item - q * M;
        const size_t nq = (size_t)n * p.Lq + q;
        const TP *of = static_cast<const TP *>(p.offsets) + nq * p.off_stride + ((size_t)m * LP + pt) * 2;
        const float *rf = p.ref + (nq * p.L + l) * p.ref_dim;
        const float sx = p.ref_dim == 2 ? 1.f / float(W) : __ldg(rf + 2) * (0.5f / float(P));
        const float sy = p.ref_dim == 2 ? 1.f / float(H) : __ldg(rf + 3) * (0.5f / float(P));
        const float x = fmaf(ldp<TP>(of), sx, __ldg(rf)), y = fmaf(ldp<TP>(of + 1), sy, __ldg(rf + 1));
        const float a = s_prob[il * LP + pt];
        const float h_im = y * float(H) - 0.5f, w_im = x * float(W) - 0.5f;
        const bool inr = h_im > -1.f && w_im > -1.f && h_im < float(H) && w_im < float(W);
        const float hf = floorf(h_im), wf = floorf(w_im);
        const int h0 = inr ? int(hf) : 0, w0 = inr ? int(wf) : 0;
        const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
        const bool top = inr && h0 >= 0, bot = inr && h0 + 1 <= H - 1, lef = w0 >= 0, rig = w0 + 1 <= W - 1;
        // pair entry e = s + 1 holds [value[s], value[s+1]]; s = start + h0*W + w0 is the LEFT corner (>= -1)
        const int e_top = start + h0 * W + w0 + 1, e_bot = e_top + W;
        const int e_safe = top ? e_top : (bot ? e_bot : 0);
        PointRec r;
        r.off_top = uint32_t(((top ? e_top : e_safe) * M + m)) * 128u;
        r.off_bot = uint32_t(((bot ? e_bot : e_safe) * M + m)) * 128u;
        r.w_top = __floats2half2_rn((top && lef) ? hh * hw * a : 0.f, (top && rig) ? hh * lw * a : 0.f);
        r.w_bot = __floats2half2_rn((bot && lef) ? lh * hw * a : 0.f, (bot && rig) ? lh * lw * a : 0.f);
        s_rec[il * LPs + pt] = r;
      }
    }
  }
  __syncthreads();

  // phase 2: 8 lanes per item; lanes 0-3 = left corner, 4-7 = right corner, 8 bf16 channels each
  const int g = lane >> 3, j = lane & 7;
  const bool right = j >= 4;
  const char *base = static_cast<const char *>(p.pairs) + (size_t)n * (p.S + 1) * M * 128 + j * 16;
  for (int il = warp * 4 + g; il < nitems; il += kWarps * 4) {
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    const uint4 *rec = reinterpret_cast<const uint4 *>(s_rec) + il * LPs;
#pragma unroll 4
    for (int pt = 0; pt < LP; ++pt) {
      const uint4 r = rec[pt];
      const uint4 vt = __ldg(reinterpret_cast<const uint4 *>(base + r.x));
      const uint4 vb = __ldg(reinterpret_cast<const uint4 *>(base + r.y));
      const __half2 wt2 = *reinterpret_cast<const __half2 *>(&r.z), wb2 = *reinterpret_cast<const __half2 *>(&r.w);
      const float wt = right ? __high2float(wt2) : __low2float(wt2);
      const float wb = right ? __high2float(wb2) : __low2float(wb2);
      const uint32_t t[4] = {vt.x, vt.y, vt.z, vt.w}, b[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        acc[2 * k] = fmaf(wt, __uint_as_float(t[k] << 16), acc[2 * k]);
        acc[2 * k + 1] = fmaf(wt, __uint_as_float(t[k] & 0xffff0000u), acc[2 * k + 1]);
        acc[2 * k] = fmaf(wb, __uint_as_float(b[k] << 16), acc[2 * k]);
        acc[2 * k + 1] = fmaf(wb, __uint_as_float(b[k] & 0xffff0000u), acc[2 * k + 1]);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 4);
    if (!right) {
      const int item = s_item[il];
      const int q = p.m_shift >= 0 ? item >> p.m_shift : item / M, m = item - q * M;
      __nv_bfloat16 *dst = static_cast<__nv_bfloat16 *>(p.out) + (((size_t)n * p.Lq + q) * M + m) * (size_t)kD + j * 8;
      *reinterpret_cast<uint4 *>(dst) = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]),
                                                   pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
    }
  }
}

int log2_exact(int v) {
  if (v <= 0 || (v & (v - 1))) return -1;
  int s = 0;
  while ((1 << s) < v) ++s;
  return s;
}

}  // namespace
}  // namespace dvis

using namespace dvis;

extern "C" int dvis_msda_pack_pairs(const void *value, int batch, int spatial_size, int num_heads, int channels, void *pairs,
                                    void *stream) {
  DVIS_REQUIRE(value && pairs, "msda_pack_pairs: null pointer argument");
  DVIS_REQUIRE(batch > 0 && spatial_size > 0 && num_heads > 0, "msda_pack_pairs: sizes must be positive");
  if (channels != kD) return fail(DVIS_ERR_UNSUPPORTED, "msda_pack_pairs: channels %d (built for 32 bf16 channels per head)", channels);
  DVIS_REQUIRE(aligned16(value) && aligned16(pairs), "msda_pack_pairs: pointers must be 16-byte aligned");
  const int64_t total = (int64_t)batch * (spatial_size + 1) * num_heads * 8;
  const int blocks = int(std::min<int64_t>((total + 255) / 256, (int64_t)kNumSMs * 32));
  pack_pairs_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4 *>(value),
                                                                         static_cast<uint4 *>(pairs), spatial_size, num_heads, total);
  return check_launch("pack_pairs_kernel");
}

extern "C" int dvis_msda_pair_forward(const void *pairs, const int64_t *spatial_shapes, const int64_t *level_start,
                                      const void *offsets, int64_t offsets_stride, const void *logits, int64_t logits_stride,
                                      int param_dtype, const float *ref, int ref_dim, int batch, int spatial_size,
                                      int num_heads, int channels, int num_levels, int num_query, int num_point,
                                      const int32_t *item_order, void *out, void *stream) {
  DVIS_REQUIRE(pairs && spatial_shapes && level_start && offsets && logits && ref && out, "msda_pair_forward: null pointer argument");
  DVIS_REQUIRE(batch > 0 && spatial_size > 0 && num_heads > 0 && num_levels > 0 && num_query > 0 && num_point > 0,
               "msda_pair_forward: sizes must be positive");
  DVIS_REQUIRE(ref_dim == 2 || ref_dim == 4, "msda_pair_forward: reference points must be 2-d or 4-d");
  DVIS_REQUIRE(num_levels <= kMaxLevels && num_levels * num_point <= kMaxLP, "msda_pair_forward: L <= 8 and L*P <= 32 required");
  DVIS_REQUIRE(aligned16(pairs) && aligned16(out), "msda_pair_forward: pairs / out must be 16-byte aligned");
  DVIS_REQUIRE((int64_t)(spatial_size + 1) * num_heads * 128 < (int64_t(1) << 32), "msda_pair_forward: pair buffer of one batch element must be < 4 GiB");
  if (channels != kD) return fail(DVIS_ERR_UNSUPPORTED, "msda_pair_forward: channels %d (built for 32)", channels);
  if (param_dtype != DVIS_F32 && param_dtype != DVIS_BF16) return fail(DVIS_ERR_UNSUPPORTED, "msda_pair_forward: offsets/logits must be f32 or bf16");
  PairParams p{};
  p.pairs = pairs; p.shapes = spatial_shapes; p.level_start = level_start; p.offsets = offsets; p.logits = logits; p.ref = ref;
  p.off_stride = offsets_stride; p.logit_stride = logits_stride; p.ref_dim = ref_dim; p.order = item_order; p.out = out;
  p.N = batch; p.S = spatial_size; p.M = num_heads; p.L = num_levels; p.Lq = num_query; p.P = num_point;
  p.m_shift = log2_exact(num_heads); p.p_shift = log2_exact(num_point);
  int lpc = 1, sh = 0;
  while (lpc < num_levels * num_point) { lpc <<= 1; ++sh; }
  p.lpc_shift = sh;
  const int per_batch = num_query * num_heads;
  dim3 grid((per_batch + kItems - 1) / kItems, batch);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int LP = num_levels * num_point;
  const size_t smem = size_t(kItems) * ((LP | 1) * sizeof(PointRec) + LP * sizeof(float));
  if (param_dtype == DVIS_F32) msda_pair_kernel<float, 6><<<grid, kThreads, smem, s>>>(p);
  else msda_pair_kernel<__nv_bfloat16, 6><<<grid, kThreads, smem, s>>>(p);
  return check_launch("msda_pair_kernel");
}
