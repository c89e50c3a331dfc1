// Fused multi-head attention core for the SHORT sequences of tracker / refiner / the predictor's query self-attention:
//   out[b, i, h, :] = softmax_j(scale * <q[b,i,h,:], k[b,j,h,:]>) @ v[b,j,h,:]
// i.e. what nn.MultiheadAttention does between its in- and out-projections inside SelfAttentionLayer /
// CrossAttentionLayer / ReferringCrossAttentionLayer
// (P/mask2former_video/modeling/transformer_decoder/video_mask2former_transformer_decoder.py:46,104;
//  P/dvis_Plus/tracker.py:45).  At 16..256 keys of 32..64 dims this is latency-, not tensor-bound: no MMA, one kernel,
// K and V of a (batch, head) staged once in shared memory, one warp per query row, fp32 math, exp2-based softmax.
// Strides are explicit so q / k / v can be slices of a packed projection output (no .contiguous() copies).
#include "common.cuh"

namespace dvis {
namespace {

constexpr int kWarpsA = 4;

struct AttnParams {
  const __nv_bfloat16 *q, *k, *v;
  __nv_bfloat16 *o;
  int64_t q_row, q_batch, k_row, k_batch, v_row, v_batch, o_row, o_batch;   // element strides; heads are Dh apart
  int B, Lq, Lk, H;
  float scale_log2e;
  int rows_per_cta;
};

template <int DH>
__global__ void __launch_bounds__(kWarpsA * 32) mha_core_kernel(const AttnParams p) {
  constexpr int KS = DH + 2;                      // padded row stride (bf16) -> conflict-free row-per-lane reads
  extern __shared__ __nv_bfloat16 smem[];
  __nv_bfloat16 *sK = smem;                       // [Lk][KS]
  __nv_bfloat16 *sV = sK + (size_t)p.Lk * KS;     // [Lk][DH]
  float *sP = reinterpret_cast<float *>(sV + (size_t)p.Lk * DH);   // [kWarpsA][Lk] probabilities of the warp's current row
  const int b = blockIdx.z, h = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const __nv_bfloat16 *kb = p.k + (size_t)b * p.k_batch + (size_t)h * DH;
  const __nv_bfloat16 *vb = p.v + (size_t)b * p.v_batch + (size_t)h * DH;
  // stage K and V of this (batch, head): 4-byte chunks, coalesced along the head dim
  for (int i = threadIdx.x; i < p.Lk * (DH / 2); i += blockDim.x) {
    const int j = i / (DH / 2), c = (i - j * (DH / 2)) * 2;
    *reinterpret_cast<uint32_t *>(sK + (size_t)j * KS + c) = *reinterpret_cast<const uint32_t *>(kb + (size_t)j * p.k_row + c);
    *reinterpret_cast<uint32_t *>(sV + (size_t)j * DH + c) = *reinterpret_cast<const uint32_t *>(vb + (size_t)j * p.v_row + c);
  }
  __syncthreads();
  float *myP = sP + (size_t)warp * p.Lk;
  const int row0 = blockIdx.x * p.rows_per_cta;
  const int row1 = min(row0 + p.rows_per_cta, p.Lq);
  for (int i = row0 + warp; i < row1; i += kWarpsA) {
    // the query row, pre-scaled, replicated in every lane's registers
    const __nv_bfloat16 *qp = p.q + (size_t)b * p.q_batch + (size_t)i * p.q_row + (size_t)h * DH;
    float q[DH];
#pragma unroll
    for (int c = 0; c < DH; c += 2) {
      const uint32_t u = *reinterpret_cast<const uint32_t *>(qp + c);
      q[c] = __uint_as_float(u << 16) * p.scale_log2e;
      q[c + 1] = __uint_as_float(u & 0xffff0000u) * p.scale_log2e;
    }
    // scores: lane j handles keys j, j+32, ...
    float mx = -INFINITY;
    for (int j = lane; j < p.Lk; j += 32) {
      const __nv_bfloat16 *kr = sK + (size_t)j * KS;
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < DH; c += 2) {
        const uint32_t u = *reinterpret_cast<const uint32_t *>(kr + c);
        s = fmaf(q[c], __uint_as_float(u << 16), s);
        s = fmaf(q[c + 1], __uint_as_float(u & 0xffff0000u), s);
      }
      myP[j] = s;
      mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < p.Lk; j += 32) {
      const float e = exp2f(myP[j] - mx);
      myP[j] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    __syncwarp();
    // out = P @ V: lane owns dims (2*lane, 2*lane+1) [+ 64 ...]
    float acc[DH / 32] = {};
    constexpr int PAIRS = DH / 64 > 0 ? DH / 64 : 1;       // DH=32: lanes 0..15 active with one pair; DH=64: one pair per lane
    (void)PAIRS;
    if (2 * lane < DH) {
      float a0 = 0.f, a1 = 0.f;
      for (int j = 0; j < p.Lk; ++j) {
        const float pj = myP[j];
        const uint32_t u = *reinterpret_cast<const uint32_t *>(sV + (size_t)j * DH + 2 * lane);
        a0 = fmaf(pj, __uint_as_float(u << 16), a0);
        a1 = fmaf(pj, __uint_as_float(u & 0xffff0000u), a1);
      }
      __nv_bfloat16 *op = p.o + (size_t)b * p.o_batch + (size_t)i * p.o_row + (size_t)h * DH + 2 * lane;
      *reinterpret_cast<uint32_t *>(op) = pack_bf16x2(a0 * inv, a1 * inv);
    }
    (void)acc;
    __syncwarp();
  }
}

}  // namespace
}  // namespace dvis

using namespace dvis;

extern "C" int dvis_mha_core(const void *q, int64_t q_row, int64_t q_batch, const void *k, int64_t k_row, int64_t k_batch,
                             const void *v, int64_t v_row, int64_t v_batch, void *out, int64_t o_row, int64_t o_batch,
                             int B, int Lq, int Lk, int H, int Dh, float scale, void *stream) {
  DVIS_REQUIRE(q && k && v && out, "mha_core: null pointer argument");
  DVIS_REQUIRE(B > 0 && Lq > 0 && Lk > 0 && H > 0, "mha_core: sizes must be positive");
  DVIS_REQUIRE(B <= 65535 && H <= 65535, "mha_core: batch / heads too large for the grid");
  if (Dh != 32 && Dh != 64) return fail(DVIS_ERR_UNSUPPORTED, "mha_core: head dim %d (built: 32, 64)", Dh);
  const auto even = [](int64_t s) { return (s & 1) == 0; };
  DVIS_REQUIRE(even(q_row) && even(q_batch) && even(k_row) && even(k_batch) && even(v_row) && even(v_batch) && even(o_row) && even(o_batch) &&
                   (reinterpret_cast<uintptr_t>(q) & 3) == 0 && (reinterpret_cast<uintptr_t>(k) & 3) == 0 &&
                   (reinterpret_cast<uintptr_t>(v) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0,
               "mha_core: pointers and strides must be 4-byte aligned");
  const size_t smem = (size_t)Lk * (Dh + 2) * 2 + (size_t)Lk * Dh * 2 + (size_t)kWarpsA * Lk * 4;
  if (smem > 200 * 1024) return fail(DVIS_ERR_UNSUPPORTED, "mha_core: Lk=%d does not fit in shared memory (use the library path)", Lk);
  AttnParams p{static_cast<const __nv_bfloat16 *>(q), static_cast<const __nv_bfloat16 *>(k), static_cast<const __nv_bfloat16 *>(v),
               static_cast<__nv_bfloat16 *>(out), q_row, q_batch, k_row, k_batch, v_row, v_batch, o_row, o_batch, B, Lq, Lk, H,
               scale * 1.4426950408889634f, 0};
  // enough CTAs to fill the machine, but not so many that K/V staging is repeated needlessly
  int rows = Lq;
  while (rows > kWarpsA && (int64_t)((Lq + rows - 1) / rows) * H * B < 2 * kNumSMs) rows = (rows + 1) / 2;
  p.rows_per_cta = rows;
  dim3 grid((Lq + rows - 1) / rows, H, B);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (Dh == 32) {
    cudaFuncSetAttribute(mha_core_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    mha_core_kernel<32><<<grid, kWarpsA * 32, smem, s>>>(p);
  } else {
    cudaFuncSetAttribute(mha_core_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    mha_core_kernel<64><<<grid, kWarpsA * 32, smem, s>>>(p);
  }
  return check_launch("mha_core_kernel");
}
