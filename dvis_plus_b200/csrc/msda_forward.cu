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
//   * per-point offsets and bilinear x attention weights are computed ONCE per (item, point) by one thread into shared
//     memory (phase 1) and then consumed by the lanes that gather (phase 2): 2 LDS.128 + 4 LDG.128 + 8 FFMA2 per point;
//   * a CTA walks a contiguous chunk of an optional `item_order` permutation -- the host orders items as
//     2-D query tiles per head so that the lines a CTA touches stay L1-resident (dvis_plus_b200/locality.py);
//   * the fused variant also does softmax(logits) and loc = ref + offset / (W, H) in the kernel, removing the
//     sampling_locations / attention_weights round trip through HBM (OPS/modules/ms_deform_attn.py:101-112).
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace dvis {
namespace {

constexpr int kMaxLevels = 8;
constexpr int kMaxStagedLP = 32;   // L*P beyond this goes to the generic kernel
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
  int m_shift, p_shift, lpc_shift;  // log2(M), log2(P) or -1 when not a power of two; log2(next_pow2(L*P))
  unsigned lp_magic;                // ceil(2^20 / (L*P)): e / (L*P) == (e * lp_magic) >> 20 for e < 2^12
};

template <typename T>
struct LocType { using type = float; };   // plain-op loc/attn dtype for value dtype T (bf16 value is fused-only)

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

// Parameters of one sampling point, computed once by one thread (phase 1) and then read by the LPR lanes that
// gather the point's rows (phase 2): byte offsets of the 4 corner rows relative to this batch element's
// value base, and the 4 bilinear weights already multiplied by the attention weight.  A weight of exactly 0
// marks a corner outside the map (or a whole point outside (-1, size)); its offset is redirected to a valid row.
struct PointOffsets { uint32_t o00, o01, o10, o11; };   // BYTE offsets (unsigned: one IADD3 + IADD3.X per address)

constexpr uint32_t kNoCorner = 0xffffffffu;   // offset of a corner outside the map when kSkip is set: phase 2 does not load it

// kSkip = false: corners outside the map keep weight 0 and are redirected to a valid row (unconditional loads: the inference
// kernels).  kSkip = true (the plain op = the reference's own contract): their offset is kNoCorner and phase 2 SKIPS the load like
// the reference does (cuh:61-83), so a NaN / Inf anywhere in `value` reaches exactly the outputs it reaches in the reference.
template <typename A, bool kSkip = false>
__device__ __forceinline__ void point_params(A x, A y, A a, int H, int W, int start, int M, int m, int lpr,
                                             PointOffsets &off, float4 &wt) {
  // un-fused multiply / subtract like the reference (cuh:290-291) so that floor() sees the same value
  const A h_im = y * A(H) - A(0.5);
  const A w_im = x * A(W) - A(0.5);
  const bool inr = (h_im > A(-1)) && (w_im > A(-1)) && (h_im < A(H)) && (w_im < A(W));   // cuh:293
  const A hf = floor(h_im), wf = floor(w_im);
  const int h0 = inr ? int(hf) : 0, w0 = inr ? int(wf) : 0;
  const A lh = h_im - hf, lw = w_im - wf;
  const A hh = A(1) - lh, hw = A(1) - lw;
  const bool top = inr && h0 >= 0, bot = inr && h0 + 1 <= H - 1;      // cuh:61-83 corner tests
  const bool lef = w0 >= 0, rig = w0 + 1 <= W - 1;
  const bool v00 = top && lef, v01 = top && rig, v10 = bot && lef, v11 = bot && rig;
  wt.x = v00 ? float(hh * hw * a) : 0.f;
  wt.y = v01 ? float(hh * lw * a) : 0.f;
  wt.z = v10 ? float(lh * hw * a) : 0.f;
  wt.w = v11 ? float(lh * lw * a) : 0.f;
  const int rowu = M * lpr;                                           // one pixel = M*D elements = M*LPR 16-byte units
  const int o00 = ((start + h0 * W + w0) * M + m) * lpr;
  const int o01 = o00 + rowu, o10 = o00 + W * rowu, o11 = o10 + rowu;
  // Phase 2 loads all four corners unconditionally; a corner outside the map (weight 0) is pointed at a valid
  // corner of the same point (or at pixel 0 of this head when the whole point is out of range), so the load is
  // always in bounds and contributes 0 * finite = 0.
  if constexpr (kSkip) {
    off.o00 = v00 ? uint32_t(o00) << 4 : kNoCorner;
    off.o01 = v01 ? uint32_t(o01) << 4 : kNoCorner;
    off.o10 = v10 ? uint32_t(o10) << 4 : kNoCorner;
    off.o11 = v11 ? uint32_t(o11) << 4 : kNoCorner;
    return;
  }
  const int safe = v00 ? o00 : v01 ? o01 : v10 ? o10 : v11 ? o11 : m * lpr;
  off.o00 = uint32_t(v00 ? o00 : safe) << 4;
  off.o01 = uint32_t(v01 ? o01 : safe) << 4;
  off.o10 = uint32_t(v10 ? o10 : safe) << 4;
  off.o11 = uint32_t(v11 ? o11 : safe) << 4;
}

// Head-major value layout (N, M, S, D) with D = 32 bf16 (64 bytes per pixel): the two x-adjacent corners of a sampling point are
// 128 CONTIGUOUS bytes, one L1 line when the left pixel index is even.  A point is described by the byte offsets of its two
// column PAIRS (row h0 and row h0 + 1; relative to the batch item's value base, head slab included) and the weights of the
// (left, right) pixel of each pair.  At the borders the pair is shifted inside the row -- w0 == -1 reads columns (0, 1) with
// the right weight 0, w0 == W - 1 reads (W - 2, W - 1) with the left weight 0 -- and a row outside the map is clamped to the
// nearest valid row with weights 0, so every load is in bounds and a zero weight multiplies a finite value.
//   rec.x / rec.y: byte offsets of the top / bottom pair      rec.z: bf16x2 (top, bottom) weight of the LEFT pixel
//   rec.w: bf16x2 (top, bottom) weight of the RIGHT pixel
__device__ __forceinline__ uint4 point_params_hm(float x, float y, float a, int H, int W, int start, int slab, int S) {
  const float h_im = y * float(H) - 0.5f;
  const float w_im = x * float(W) - 0.5f;
  const bool inr = (h_im > -1.f) && (w_im > -1.f) && (h_im < float(H)) && (w_im < float(W));   // cuh:293
  const float hf = floorf(h_im), wf = floorf(w_im);
  const int h0 = inr ? int(hf) : 0, w0 = inr ? int(wf) : 0;
  const float lh = h_im - hf, lw = w_im - wf;
  const float hh = 1.f - lh, hw = 1.f - lw;
  const bool top = inr && h0 >= 0, bot = inr && h0 + 1 <= H - 1;      // cuh:61-83 corner tests
  const bool lef = w0 >= 0, rig = w0 + 1 <= W - 1;
  const float w00 = (top && lef) ? hh * hw * a : 0.f, w01 = (top && rig) ? hh * lw * a : 0.f;
  const float w10 = (bot && lef) ? lh * hw * a : 0.f, w11 = (bot && rig) ? lh * lw * a : 0.f;
  int wb = w0;
  float tl = w00, tr = w01, bl = w10, br = w11;
  if (!lef) { wb = 0; tl = w01; tr = 0.f; bl = w11; br = 0.f; }                      // column 0 is the point's RIGHT corner
  else if (!rig) { wb = max(W - 2, 0); tl = 0.f; tr = w00; bl = 0.f; br = w10; }     // column W-1 is the point's LEFT corner
  const int r0 = max(h0, 0), r1 = min(h0 + 1, H - 1);
  const int p0 = min(start + r0 * W + wb, S - 2), p1 = min(start + r1 * W + wb, S - 2);   // (W == 1 maps: stay inside the slab)
  uint4 rec;
  rec.x = uint32_t(slab + p0) << 6;
  rec.y = uint32_t(slab + p1) << 6;
  rec.z = pack_bf16x2(tl, bl);
  rec.w = pack_bf16x2(tr, br);
  return rec;
}

// Blackwell mixed-precision FMA (PTX fma.rn.f32.bf16 -> SASS FHFMA.BF16 with .H0 / .H1 operand selectors): fp32 accumulator +=
// bf16 x bf16 straight from the halves of packed registers -- no bf16 -> fp32 conversion instructions (they were 32 of the
// 54 instructions per sampling point of the bf16 gather loop).  a0 += lo(v) * w, a1 += hi(v) * w with w = the HI-th half of wp.
template <int HI>
__device__ __forceinline__ void fhfma2(float &a0, float &a1, uint32_t v, uint32_t wp) {
#ifndef DVIS_SIMT_EMULATION
  if (HI)
    asm("{\n .reg .b16 vl, vh, wl, wh;\n mov.b32 {vl, vh}, %2;\n mov.b32 {wl, wh}, %3;\n"
        " fma.rn.f32.bf16 %0, vl, wh, %0;\n fma.rn.f32.bf16 %1, vh, wh, %1;\n}" : "+f"(a0), "+f"(a1) : "r"(v), "r"(wp));
  else
    asm("{\n .reg .b16 vl, vh, wl, wh;\n mov.b32 {vl, vh}, %2;\n mov.b32 {wl, wh}, %3;\n"
        " fma.rn.f32.bf16 %0, vl, wl, %0;\n fma.rn.f32.bf16 %1, vh, wl, %1;\n}" : "+f"(a0), "+f"(a1) : "r"(v), "r"(wp));
#else
  const float w = __uint_as_float(HI ? (wp & 0xffff0000u) : (wp << 16));
  a0 = fmaf(__uint_as_float(v << 16), w, a0);
  a1 = fmaf(__uint_as_float(v & 0xffff0000u), w, a1);
#endif
}

// ---------------------------------------------------------------------------------------------------
// staged kernel: T = value type, TO = output type, D = channels per head.
// A CTA owns `items_per_cta` (query, head) items.
//   phase 0 (FUSED): softmax over the L*P logits of each item -> smem
//   phase 1: one thread per (item, point) computes offsets + weights -> smem           (no redundancy)
//   phase 2: a group of LPR lanes per item walks its points: 2 LDS.128 + 4 predicated LDG.128 + 4*VEC FFMA each
// ---------------------------------------------------------------------------------------------------
template <typename TP>
__device__ __forceinline__ float ldp(const TP *p);
template <>
__device__ __forceinline__ float ldp<float>(const float *p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ldp<__nv_bfloat16>(const __nv_bfloat16 *p) {
  return __bfloat162float(__ldg(p));
}

// TP: dtype of the FUSED path's offsets / logits (float or bf16); unused for the plain op
// HM: `value` is head-major (N, M, S, D) -- bf16 in / out, D = 32, FUSED only; see point_params_hm
template <typename T, typename TO, int D, bool FUSED, typename TP = float, int MINB = 6, int UNROLL = 2, bool HM = false>
__global__ void __launch_bounds__(kThreads, MINB) msda_fwd_staged_kernel(const MsdaParams p) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int LPR = D / VEC;  // lanes per row
  constexpr int G = 32 / LPR;   // items per warp
  static_assert(D % VEC == 0 && LPR >= 1 && LPR <= 32 && (LPR & (LPR - 1)) == 0, "unsupported head dim");

  extern __shared__ uint4 dyn_smem[];
  __shared__ int sH[kMaxLevels], sW[kMaxLevels], sStart[kMaxLevels];
  __shared__ float sInvH[kMaxLevels], sInvW[kMaxLevels];
  __shared__ unsigned char sLevelOf[kMaxStagedLP];          // level of sampling point pt (pt / P), looked up instead of divided
  if (threadIdx.x < p.L) {
    sH[threadIdx.x] = int(p.shapes[2 * threadIdx.x]);
    sW[threadIdx.x] = int(p.shapes[2 * threadIdx.x + 1]);
    sStart[threadIdx.x] = int(p.level_start[threadIdx.x]);
    sInvH[threadIdx.x] = 1.f / float(int(p.shapes[2 * threadIdx.x]));
    sInvW[threadIdx.x] = 1.f / float(int(p.shapes[2 * threadIdx.x + 1]));
  }
  if (threadIdx.x < p.L * p.P) sLevelOf[threadIdx.x] = (unsigned char)(threadIdx.x / p.P);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.y, M = p.M, P = p.P, LP = p.L * p.P;
  const int LPs = LP | 1;       // odd stride (in 16-byte units) -> the G groups of a warp hit disjoint banks
  const int per_batch = p.Lq * M;
  const int chunk_begin = blockIdx.x * p.items_per_cta;
  const int nitems = min(p.items_per_cta, per_batch - chunk_begin);
  // bf16 in AND out (the encoder's production path): the 4 corner weights are kept as packed bf16 and the gather loop runs on
  // FHFMA (fhfma2 above); every other combination keeps exact fp32 weights (the plain op's 2e-5 contract)
  constexpr bool kMixed = std::is_same<T, __nv_bfloat16>::value && std::is_same<TO, __nv_bfloat16>::value;
  static_assert(!HM || (kMixed && FUSED && D == 32), "head-major value: bf16 in / out, 32 channels per head, fused parameters");
  PointOffsets *s_off = reinterpret_cast<PointOffsets *>(dyn_smem);
  float4 *s_wt = reinterpret_cast<float4 *>(dyn_smem + p.items_per_cta * LPs);   // kMixed: a dense uint2 array (2 x bf16x2) in the same region
  // head-major: ONE 16-byte record per point (no separate weight array) and 2 floats per item of softmax state: 14 KB per CTA
  // instead of 30 KB, i.e. ~100 KB more L1 for the gather at 6 CTAs per SM
  float *s_prob = reinterpret_cast<float *>(dyn_smem + (HM ? 1 : 2) * p.items_per_cta * LPs);
  int *s_item = reinterpret_cast<int *>(s_prob + (FUSED ? p.items_per_cta * (HM ? 2 : LP) : 0));   // s_prob: (max, 1/sum) per item

  for (int il = tid; il < nitems; il += kThreads) s_item[il] = p.order ? __ldg(p.order + chunk_begin + il) : chunk_begin + il;
  __syncthreads();

  if constexpr (FUSED) {
    // phase 0: ONE thread per item: softmax over its L*P contiguous logits (OPS/modules/ms_deform_attn.py:103-104).  The row
    // (24 bytes at L*P = 12, bf16) is read twice from L1 (max, then exp) -- the first version used 4 threads + shuffles per
    // item and cost 13 % of the kernel's instructions (ncu source view, profiles/r2_msda.md); this form costs ~1.5 %.
    // The row is fetched with 8-byte loads into registers (3 loads at L*P = 12, bf16): a 2-byte load per logit would cost one
    // L1 wavefront per lane and logit -- the kernel is L1-wavefront bound -- and made the first one-thread version slower.
    constexpr int EPV = 8 / int(sizeof(TP));                         // logits per 8-byte load
    const bool vec_ok = (LP % EPV) == 0 && (reinterpret_cast<uintptr_t>(p.attn) & 7) == 0 && (p.attn_stride % EPV) == 0;
    for (int il = tid; il < nitems; il += kThreads) {
      const int item = s_item[il];
      const int q = p.m_shift >= 0 ? item >> p.m_shift : item / M, m = item - q * M;
      const TP *lg = static_cast<const TP *>(p.attn) + ((size_t)n * p.Lq + q) * p.attn_stride + (size_t)m * LP;
      // only (max, 1 / sum) of the row go to shared memory; phase 1 recomputes exp(logit - max) for its own point (one L1-hit
      // load + one MUFU) -- storing and re-reading L*P probabilities cost 4.5 shared-memory wavefronts per item with 4-way
      // bank conflicts, on a kernel that is bound by L1 / shared-memory wavefronts (ncu: 91 % of the LSU data pipe)
      float mx = -INFINITY, sum = 0.f;
      if (vec_ok) {
        const int nv = LP / EPV;
        auto unpack = [](const uint2 u, float (&v)[EPV]) {
          if constexpr (EPV == 4) {
            v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
            v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
          } else {
            v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y);
          }
        };
        for (int i = 0; i < nv; ++i) {
          float v[EPV];
          unpack(__ldg(reinterpret_cast<const uint2 *>(lg) + i), v);
#pragma unroll
          for (int k = 0; k < EPV; ++k) mx = fmaxf(mx, v[k]);
        }
        for (int i = 0; i < nv; ++i) {                               // second read: L1 hit (the 40-register budget of the
          float v[EPV];                                               // 6-CTA/SM gather loop does not hold 32 logits)
          unpack(__ldg(reinterpret_cast<const uint2 *>(lg) + i), v);
#pragma unroll
          for (int k = 0; k < EPV; ++k) sum += __expf(v[k] - mx);
        }
      } else {
        for (int i = 0; i < LP; ++i) mx = fmaxf(mx, ldp<TP>(lg + i));
        for (int i = 0; i < LP; ++i) sum += __expf(ldp<TP>(lg + i) - mx);
      }
      *reinterpret_cast<float2 *>(s_prob + 2 * il) = make_float2(mx, 1.f / sum);
    }
    __syncthreads();
  }

  // phase 1: one thread per (item, point), ALL lanes busy: the flat index e = il * LP + pt is split with a multiply-shift
  // (magic = ceil(2^20 / LP), exact for e < 2^12), the point's level comes from a table, 1 / W and 1 / H from shared memory
  {
    const int total = nitems * LP;
    const unsigned magic = p.lp_magic;
    for (int e = tid; e < total; e += kThreads) {
      const int il = int((unsigned(e) * magic) >> 20), pt = e - il * LP;
      const int l = sLevelOf[pt];
      const int H = sH[l], W = sW[l], start = sStart[l];
      const int item = s_item[il];
      const int q = p.m_shift >= 0 ? item >> p.m_shift : item / M, m = item - q * M;
      const size_t nq = (size_t)n * p.Lq + q;
      PointOffsets off;
      float4 wt;
      if constexpr (FUSED) {
        const TP *of = static_cast<const TP *>(p.loc) + nq * p.loc_stride + ((size_t)m * LP + pt) * 2;
        const float *rf = p.ref + (nq * p.L + l) * p.ref_dim;
        float2 o;
        if constexpr (std::is_same<TP, __nv_bfloat16>::value) {       // x and y offsets: one 32-bit load
          const uint32_t u = __ldg(reinterpret_cast<const uint32_t *>(of));
          o = make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
        } else {
          o = __ldg(reinterpret_cast<const float2 *>(of));
        }
        // loc = ref + off / (W_l, H_l) for 2-d reference points (py:106-109);
        // loc = ref_xy + off / P * ref_wh * 0.5 for boxes (py:110-112)
        float rx, ry, sx, sy;
        if (p.ref_dim == 2) {
          const float2 r2 = __ldg(reinterpret_cast<const float2 *>(rf));
          rx = r2.x; ry = r2.y; sx = sInvW[l]; sy = sInvH[l];
        } else {
          const float4 r4 = __ldg(reinterpret_cast<const float4 *>(rf));
          rx = r4.x; ry = r4.y; sx = r4.z * (0.5f / float(P)); sy = r4.w * (0.5f / float(P));
        }
        const float2 ms = *reinterpret_cast<const float2 *>(s_prob + 2 * il);      // (max, 1 / sum) of the item's logits
        const TP *lgp = static_cast<const TP *>(p.attn) + nq * p.attn_stride + (size_t)m * LP + pt;
        const float prob = __expf(ldp<TP>(lgp) - ms.x) * ms.y;
        if constexpr (HM) {
          reinterpret_cast<uint4 *>(s_off)[il * LPs + pt] = point_params_hm(fmaf(o.x, sx, rx), fmaf(o.y, sy, ry), prob, H, W, start, m * p.S, p.S);
          continue;
        }
        point_params<float>(fmaf(o.x, sx, rx), fmaf(o.y, sy, ry), prob, H, W, start, M, m, LPR, off, wt);
      } else {
        using TL = typename LocType<T>::type;   // plain op: loc / attn have the value's dtype
        const size_t i = (nq * M + m) * (size_t)LP + pt;
        const TL *lc = static_cast<const TL *>(p.loc) + 2 * i;
        point_params<TL, true>(__ldg(lc), __ldg(lc + 1), __ldg(static_cast<const TL *>(p.attn) + i), H, W, start, M, m, LPR,
                               off, wt);
      }
      s_off[il * LPs + pt] = off;
      if constexpr (kMixed)
        reinterpret_cast<uint2 *>(s_wt)[il * LPs + pt] = make_uint2(pack_bf16x2(wt.x, wt.y), pack_bf16x2(wt.z, wt.w));   // dense 8-byte slots
      else
        s_wt[il * LPs + pt] = wt;
    }
  }
  __syncthreads();

  // phase 2: gather + accumulate
  if constexpr (HM) {
    // 8 lanes per item: lanes 0-3 read the 64 bytes of the LEFT pixel of a pair, lanes 4-7 the RIGHT pixel -- one 128-byte
    // request per pair (one L1 wavefront when aligned, else two; the token-major layout always costs two), two pairs per point.
    // Each half accumulates its own pixel's contribution for the same 8 channels; one xor-shuffle per channel joins them.
    const int g8 = lane >> 3, j8 = lane & 7, half = j8 >> 2;
    const char *vb8 = reinterpret_cast<const char *>(static_cast<const T *>(p.value) + (size_t)n * p.S * M * D) + j8 * 16;
    const int rounds = (nitems + kWarps * 4 - 1) / (kWarps * 4);
    for (int r = 0; r < rounds; ++r) {
      const int il_raw = (r * kWarps + warp) * 4 + g8;
      const bool live = il_raw < nitems;
      const int il = live ? il_raw : nitems - 1;             // idle groups repeat the last item (all lanes reach the shuffles)
      float acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
      const uint4 *so = reinterpret_cast<const uint4 *>(s_off) + il * LPs;
#pragma unroll UNROLL
      for (int pt = 0; pt < LP; ++pt) {
        const uint4 rec = so[pt];
        const uint32_t w = half ? rec.w : rec.z;
        const uint4 va = __ldg(reinterpret_cast<const uint4 *>(vb8 + rec.x)), vc = __ldg(reinterpret_cast<const uint4 *>(vb8 + rec.y));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          fhfma2<0>(acc[2 * k], acc[2 * k + 1], (&va.x)[k], w);
          fhfma2<1>(acc[2 * k], acc[2 * k + 1], (&vc.x)[k], w);
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 4);
      if (live && half == 0) {
        const int item = s_item[il];
        const int q = p.m_shift >= 0 ? item >> p.m_shift : item / M, m = item - q * M;
        TO *dst = static_cast<TO *>(p.out) + (((size_t)n * p.Lq + q) * M + m) * (size_t)D + j8 * 8;
        store_vec<TO, float, 8>(dst, acc);
      }
    }
    return;
  }
  const int g = lane / LPR, j = lane % LPR;
  using V = decltype(Vec16<T>::v);
  const char *vb = reinterpret_cast<const char *>(static_cast<const T *>(p.value) + (size_t)n * p.S * M * D) + j * 16;
  for (int il = warp * G + g; il < nitems; il += kWarps * G) {
    if constexpr (kMixed) {
      float acc[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
      const uint4 *so = reinterpret_cast<const uint4 *>(s_off) + il * LPs;
      const uint2 *sw = reinterpret_cast<const uint2 *>(s_wt) + il * LPs;
#pragma unroll UNROLL
      for (int pt = 0; pt < LP; ++pt) {
        const uint4 o = so[pt];
        const uint2 w = sw[pt];
        const uint4 v00 = __ldg(reinterpret_cast<const uint4 *>(vb + o.x)), v01 = __ldg(reinterpret_cast<const uint4 *>(vb + o.y));
        const uint4 v10 = __ldg(reinterpret_cast<const uint4 *>(vb + o.z)), v11 = __ldg(reinterpret_cast<const uint4 *>(vb + o.w));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          fhfma2<0>(acc[2 * k], acc[2 * k + 1], (&v00.x)[k], w.x);
          fhfma2<1>(acc[2 * k], acc[2 * k + 1], (&v01.x)[k], w.x);
          fhfma2<0>(acc[2 * k], acc[2 * k + 1], (&v10.x)[k], w.y);
          fhfma2<1>(acc[2 * k], acc[2 * k + 1], (&v11.x)[k], w.y);
        }
      }
      const int item = s_item[il];
      const int q = p.m_shift >= 0 ? item >> p.m_shift : item / M, m = item - q * M;
      TO *dst = static_cast<TO *>(p.out) + (((size_t)n * p.Lq + q) * M + m) * (size_t)D + j * VEC;
      store_vec<TO, float, VEC>(dst, acc);
      continue;
    }
    // accumulators as f32x2 pairs: Blackwell's packed FFMA2 (fma.rn.f32x2) halves the FMA issue slots
    float2 acc2[VEC / 2];
#pragma unroll
    for (int k = 0; k < VEC / 2; ++k) acc2[k] = make_float2(0.f, 0.f);
    const uint4 *so = reinterpret_cast<const uint4 *>(s_off) + il * LPs;
    const float4 *sw = s_wt + il * LPs;
#pragma unroll UNROLL
    for (int pt = 0; pt < LP; ++pt) {
      const uint4 o = so[pt];
      const float4 w = sw[pt];
      Vec16<T> v00, v01, v10, v11;
      if constexpr (!FUSED) {          // the plain op: corners outside the map are not loaded at all (reference semantics)
        v00.v = V{}; v01.v = V{}; v10.v = V{}; v11.v = V{};
        if (o.x != kNoCorner) v00.v = __ldg(reinterpret_cast<const V *>(vb + o.x));
        if (o.y != kNoCorner) v01.v = __ldg(reinterpret_cast<const V *>(vb + o.y));
        if (o.z != kNoCorner) v10.v = __ldg(reinterpret_cast<const V *>(vb + o.z));
        if (o.w != kNoCorner) v11.v = __ldg(reinterpret_cast<const V *>(vb + o.w));
      } else {
        v00.v = __ldg(reinterpret_cast<const V *>(vb + o.x));
        v01.v = __ldg(reinterpret_cast<const V *>(vb + o.y));
        v10.v = __ldg(reinterpret_cast<const V *>(vb + o.z));
        v11.v = __ldg(reinterpret_cast<const V *>(vb + o.w));
      }
      const float2 wx = make_float2(w.x, w.x), wy = make_float2(w.y, w.y), wz = make_float2(w.z, w.z), ww = make_float2(w.w, w.w);
#pragma unroll
      for (int k = 0; k < VEC / 2; ++k) {
        acc2[k] = __ffma2_rn(wx, make_float2(v00.get(2 * k), v00.get(2 * k + 1)), acc2[k]);
        acc2[k] = __ffma2_rn(wy, make_float2(v01.get(2 * k), v01.get(2 * k + 1)), acc2[k]);
        acc2[k] = __ffma2_rn(wz, make_float2(v10.get(2 * k), v10.get(2 * k + 1)), acc2[k]);
        acc2[k] = __ffma2_rn(ww, make_float2(v11.get(2 * k), v11.get(2 * k + 1)), acc2[k]);
      }
    }
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC / 2; ++k) { acc[2 * k] = acc2[k].x; acc[2 * k + 1] = acc2[k].y; }
    const int item = s_item[il];
    const int q = p.m_shift >= 0 ? item >> p.m_shift : item / M, m = item - q * M;
    TO *dst = static_cast<TO *>(p.out) + (((size_t)n * p.Lq + q) * M + m) * (size_t)D + j * VEC;
    store_vec<TO, float, VEC>(dst, acc);
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

int log2_exact(int v) {
  if (v <= 0 || (v & (v - 1))) return -1;
  int s = 0;
  while ((1 << s) < v) ++s;
  return s;
}

constexpr size_t kMaxStagedSmem = 227 * 1024;   // opt-in dynamic shared memory per CTA on sm_100

// dynamic shared memory the staged kernel needs (same formula as launch_staged): lets the plain op fall back to the generic
// kernel for shapes whose staging does not fit (D = 8 fp32 with L*P >= 24 asks for > 100 KB)
template <typename T, int D, bool FUSED>
size_t staged_smem(int L, int P) {
  constexpr int G = 32 / (D / Vec16<T>::N);
  const int LP = L * P, LPs = LP | 1;
  const int items = std::max(64, kWarps * G);
  return (size_t)items * (2 * LPs * 16 + (FUSED ? LP * 4 : 0) + 4);
}

template <typename T, typename TO, int D, bool FUSED, typename TP = float, bool HM = false>
int launch_staged(MsdaParams p, cudaStream_t stream) {
  constexpr int G = 32 / (D / Vec16<T>::N);
  const int per_batch = p.Lq * p.M, LP = p.L * p.P, LPs = LP | 1;
  p.items_per_cta = 64;            // = 2 phase-2 passes at G=4; phase 0 maps 4 threads per item
  static_assert(kWarps * G <= 64 || true, "");
  if (kWarps * G > p.items_per_cta) p.items_per_cta = kWarps * G;
  p.m_shift = log2_exact(p.M);
  p.p_shift = log2_exact(p.P);
  int lpc = 1, sh = 0;
  while (lpc < LP) { lpc <<= 1; ++sh; }
  p.lpc_shift = sh;
  p.lp_magic = ((1u << 20) + LP - 1) / LP;
  const size_t smem = HM ? (size_t)p.items_per_cta * (LPs * 16 + 2 * 4 + 4)
                         : (size_t)p.items_per_cta * (2 * LPs * 16 + (FUSED ? LP * 4 : 0) + 4);
  if (smem > kMaxStagedSmem)
    return fail(DVIS_ERR_UNSUPPORTED, "msda: %zu bytes of shared memory per CTA for D=%d, L*P=%d exceed the %d-byte opt-in limit", smem,
                D, LP, int(kMaxStagedSmem));
  // head-major: 8 CTAs per SM (32 registers, 14 KB of shared memory each) with the 12-point loop fully unrolled: 541 us isolated
  // at T = 16 vs 557 us for 6 CTAs per SM (profiles/r2_msda_hm_variants.log)
  auto kern = msda_fwd_staged_kernel<T, TO, D, FUSED, TP, HM ? 8 : 6, HM ? 12 : 2, HM>;
  if constexpr (HM) {   // the same experiment switch for the head-major gather (tests/perf/encoder_microbench.py)
    static const int variant = getenv("DVIS_MSDA_HM_VARIANT") ? atoi(getenv("DVIS_MSDA_HM_VARIANT")) : 0;
    if (variant == 1) kern = msda_fwd_staged_kernel<T, TO, D, FUSED, TP, 6, 4, HM>;
    if (variant == 2) kern = msda_fwd_staged_kernel<T, TO, D, FUSED, TP, 8, 2, HM>;
    if (variant == 3) kern = msda_fwd_staged_kernel<T, TO, D, FUSED, TP, 8, 4, HM>;
    if (variant == 4) kern = msda_fwd_staged_kernel<T, TO, D, FUSED, TP, 6, 2, HM>;
    if (variant == 5) kern = msda_fwd_staged_kernel<T, TO, D, FUSED, TP, 4, 4, HM>;
    if (variant == 6) kern = msda_fwd_staged_kernel<T, TO, D, FUSED, TP, 7, 12, HM>;
    if (variant == 7) kern = msda_fwd_staged_kernel<T, TO, D, FUSED, TP, 7, 3, HM>;
    if (variant == 8) kern = msda_fwd_staged_kernel<T, TO, D, FUSED, TP, 5, 12, HM>;
    if (variant == 9) kern = msda_fwd_staged_kernel<T, TO, D, FUSED, TP, 6, 6, HM>;
    if (variant == 10) kern = msda_fwd_staged_kernel<T, TO, D, FUSED, TP, 6, 12, HM>;
    if (variant == 11) kern = msda_fwd_staged_kernel<T, TO, D, FUSED, TP, 8, 3, HM>;
  } else {  // occupancy / unroll variants (DVIS_MSDA_VARIANT), kept for the micro-benchmark.  Default: 6 CTAs/SM, unroll 2
     // (40 registers): measured 23 % faster than 3 CTAs/SM x unroll 4 (80 registers) -- the gather is latency bound.
    static const int variant = getenv("DVIS_MSDA_VARIANT") ? atoi(getenv("DVIS_MSDA_VARIANT")) : 0;
    if (variant == 1) kern = msda_fwd_staged_kernel<T, TO, D, FUSED, TP, 3, 4>;
    if (variant == 2) kern = msda_fwd_staged_kernel<T, TO, D, FUSED, TP, 8, 2>;
    if (variant == 3) kern = msda_fwd_staged_kernel<T, TO, D, FUSED, TP, 8, 1>;
  }
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(std::max<size_t>(smem, 48 * 1024))) != cudaSuccess)
    return check_launch("msda_fwd_staged_kernel (shared-memory opt-in)");
  dim3 grid((per_batch + p.items_per_cta - 1) / p.items_per_cta, p.N);
  if (getenv("DVIS_SMEM_CARVEOUT_MSDA")) prefer_carveout(kern);   // the gather lives on its L1 hits: excluded from the experiment by default
  kern<<<grid, kThreads, smem, stream>>>(p);
  return check_launch("msda_fwd_staged_kernel");
}

template <typename T>
int launch_plain(const MsdaParams &p, int D, cudaStream_t stream) {
  const bool vec_ok = aligned16(p.value) && aligned16(p.out) && p.L * p.P <= kMaxStagedLP;
  if (vec_ok && std::is_same<T, float>::value) {
    switch (D) {   // shapes whose staging would not fit in shared memory take the generic kernel below
      case 8: if (staged_smem<float, 8, false>(p.L, p.P) <= kMaxStagedSmem) return launch_staged<float, float, 8, false>(p, stream); break;
      case 16: if (staged_smem<float, 16, false>(p.L, p.P) <= kMaxStagedSmem) return launch_staged<float, float, 16, false>(p, stream); break;
      case 32: return launch_staged<float, float, 32, false>(p, stream);
      case 64: return launch_staged<float, float, 64, false>(p, stream);
      case 128: return launch_staged<float, float, 128, false>(p, stream);
      default: break;
    }
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
  DVIS_REQUIRE((long)S * M * D < (1L << 29), "msda: spatial_size*heads*channels must be < 2^29 elements");
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
                                       const int64_t *level_start, const void *offsets, int64_t offsets_stride,
                                       const void *logits, int64_t logits_stride, int param_dtype, const float *ref, int ref_dim,
                                       int batch, int spatial_size, int num_heads, int channels, int num_levels,
                                       int num_query, int num_point, const int32_t *item_order, void *out,
                                       int out_dtype, void *stream) {
  if (int rc = validate_common(value, spatial_shapes, level_start, offsets, logits, out, batch, spatial_size, num_heads,
                               channels, num_levels, num_query, num_point))
    return rc;
  DVIS_REQUIRE(ref && (ref_dim == 2 || ref_dim == 4), "msda_fused: reference points must be 2-d or 4-d");
  DVIS_REQUIRE(num_levels * num_point >= 2, "msda_fused: needs at least 2 sampling points per head (L*P = %d)", num_levels * num_point);
  DVIS_REQUIRE(aligned16(value) && aligned16(out), "msda_fused: value/out must be 16-byte aligned");
  DVIS_REQUIRE(aligned16(ref) && (reinterpret_cast<uintptr_t>(offsets) & 7) == 0 && offsets_stride % 2 == 0,
               "msda_fused: reference points must be 16-byte aligned, offsets 8-byte aligned with an even row stride");
  DVIS_REQUIRE(param_dtype == DVIS_F32 || param_dtype == DVIS_BF16, "msda_fused: offsets/logits must be f32 or bf16");
  MsdaParams p{};
  p.value = value; p.shapes = spatial_shapes; p.level_start = level_start; p.loc = offsets; p.attn = logits;
  p.ref = ref; p.ref_dim = ref_dim; p.loc_stride = offsets_stride; p.attn_stride = logits_stride;
  p.order = item_order; p.out = out;
  p.N = batch; p.S = spatial_size; p.M = num_heads; p.L = num_levels; p.Lq = num_query; p.P = num_point;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (num_levels * num_point > kMaxStagedLP)
    return fail(DVIS_ERR_UNSUPPORTED, "msda_fused: num_levels*num_point %d > %d", num_levels * num_point, kMaxStagedLP);
#define DVIS_FUSED_D(TV, TOUT)                                                                          \
  switch (channels) {                                                                                   \
    case 16: return param_dtype == DVIS_F32 ? launch_staged<TV, TOUT, 16, true, float>(p, s)            \
                                            : launch_staged<TV, TOUT, 16, true, __nv_bfloat16>(p, s);   \
    case 32: return param_dtype == DVIS_F32 ? launch_staged<TV, TOUT, 32, true, float>(p, s)            \
                                            : launch_staged<TV, TOUT, 32, true, __nv_bfloat16>(p, s);   \
    case 64: return param_dtype == DVIS_F32 ? launch_staged<TV, TOUT, 64, true, float>(p, s)            \
                                            : launch_staged<TV, TOUT, 64, true, __nv_bfloat16>(p, s);   \
    default: break;                                                                                     \
  }
  if (value_dtype == DVIS_F32 && out_dtype == DVIS_F32) { DVIS_FUSED_D(float, float) }
  else if (value_dtype == DVIS_BF16 && out_dtype == DVIS_BF16) { DVIS_FUSED_D(__nv_bfloat16, __nv_bfloat16) }
  else if (value_dtype == DVIS_BF16 && out_dtype == DVIS_F32) { DVIS_FUSED_D(__nv_bfloat16, float) }
  else if (value_dtype == DVIS_F32 && out_dtype == DVIS_BF16) { DVIS_FUSED_D(float, __nv_bfloat16) }
#undef DVIS_FUSED_D
  return fail(DVIS_ERR_UNSUPPORTED,
              "msda_fused: no kernel for value_dtype=%d out_dtype=%d channels=%d (built: f32/bf16 value and output, channels in {16,32,64})",
              value_dtype, out_dtype, channels);
}

// The same operator on a HEAD-MAJOR value tensor (N, M, S, D) -- what csrc/linear_tc.cu's value-projection epilogue writes --
// for the production shape: bf16 value and output, 32 channels per head.  Everything else as dvis_msda_fused_forward.
extern "C" int dvis_msda_fused_forward_hm(const void *value_hm, const int64_t *spatial_shapes, const int64_t *level_start,
                                          const void *offsets, int64_t offsets_stride, const void *logits, int64_t logits_stride,
                                          int param_dtype, const float *ref, int ref_dim, int batch, int spatial_size,
                                          int num_heads, int channels, int num_levels, int num_query, int num_point,
                                          const int32_t *item_order, void *out, void *stream) {
  if (int rc = validate_common(value_hm, spatial_shapes, level_start, offsets, logits, out, batch, spatial_size, num_heads,
                               channels, num_levels, num_query, num_point))
    return rc;
  DVIS_REQUIRE(channels == 32, "msda_fused_hm: built for 32 channels per head (got %d)", channels);
  DVIS_REQUIRE(spatial_size >= 2 && (long)spatial_size * num_heads < (1L << 26), "msda_fused_hm: spatial extent out of range");
  DVIS_REQUIRE(ref && (ref_dim == 2 || ref_dim == 4), "msda_fused: reference points must be 2-d or 4-d");
  DVIS_REQUIRE(num_levels * num_point >= 2 && num_levels * num_point <= kMaxStagedLP, "msda_fused_hm: L*P = %d outside [2, %d]",
               num_levels * num_point, kMaxStagedLP);
  DVIS_REQUIRE(aligned16(value_hm) && aligned16(out), "msda_fused_hm: value / out must be 16-byte aligned (128 for the value to get one-line pairs)");
  DVIS_REQUIRE(aligned16(ref) && (reinterpret_cast<uintptr_t>(offsets) & 7) == 0 && offsets_stride % 2 == 0,
               "msda_fused: reference points must be 16-byte aligned, offsets 8-byte aligned with an even row stride");
  DVIS_REQUIRE(param_dtype == DVIS_F32 || param_dtype == DVIS_BF16, "msda_fused: offsets/logits must be f32 or bf16");
  MsdaParams p{};
  p.value = value_hm; p.shapes = spatial_shapes; p.level_start = level_start; p.loc = offsets; p.attn = logits;
  p.ref = ref; p.ref_dim = ref_dim; p.loc_stride = offsets_stride; p.attn_stride = logits_stride;
  p.order = item_order; p.out = out;
  p.N = batch; p.S = spatial_size; p.M = num_heads; p.L = num_levels; p.Lq = num_query; p.P = num_point;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return param_dtype == DVIS_F32 ? launch_staged<__nv_bfloat16, __nv_bfloat16, 32, true, float, true>(p, s)
                                 : launch_staged<__nv_bfloat16, __nv_bfloat16, 32, true, __nv_bfloat16, true>(p, s);
}
