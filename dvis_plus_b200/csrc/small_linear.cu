// Linear layers of the temporal stage (tracker / refiner blocks) as ONE kernel per dependent step:
//     Y = act( A @ W^T + bias ) [+ residual]            W (N, K) bf16 row-major (nn.Linear layout), fp32 accumulate
// where the A operand is either a bf16 matrix (optionally gathered over `taps` time-shifted rows: Conv1d with replicate
// padding as a GEMM, P/dvis_Plus/refiner.py:44-52,116-119) or is BUILT IN THE PROLOGUE from the fp32 residual stream:
//     A = LN1( LN0(src0) + src1 )                        (each step optional)
// The post-norm blocks of the reference (SelfAttentionLayer / CrossAttentionLayer / FFNLayer.forward_post,
// P/mask2former_video/modeling/transformer_decoder/video_mask2former_transformer_decoder.py:40-50,98-108,160-164;
// ReferringCrossAttentionLayer, P/dvis_Plus/tracker.py:37-50) are  x <- LN(x + f(x)).  Here every producer writes the
// PRE-norm sum (x + f(x), fp32) and every consumer normalises its own rows while its weight tiles are already in flight,
// so LayerNorm never is a kernel (or a dependent step) of its own: a tracker layer is 5 launches (QKV, attention, out-proj,
// FFN1, FFN2) instead of ~9 library kernels + 3 add_layernorm.  The normalised rows are also the next residual, so the
// kernel can store them as a side output (each CTA of a row block stores a different column slice).
//
// M = 200 rows x K = 512 is latency-bound: 32-row x 64-column CTA tiles (7 x N/64 CTAs), 4 warps of mma.sync m16n8k16,
// W (and bf16 A) tiles through a 4-stage cp.async ring; for the refiner's 3 200 rows the 64-row variant halves the
// weight re-reads.  Weight tiles never depend on the previous kernel, so with programmatic dependent launch their
// prefetch overlaps the predecessor's tail (griddepcontrol.wait sits between the W and the A loads).
#include <algorithm>
#include <cstdlib>

#include "mma.cuh"

namespace dvis {
namespace {

constexpr int kLsThreads = 256;             // 8 warps = 2 per scheduler: with 1 the kernel is bound by dependent-issue latency
constexpr int kLsBN = 64, kLsBK = 64, kLsMaxStages = 8;
constexpr int kLsRS = kLsBK + 8;            // ring row stride (bf16): 144 bytes, ldmatrix conflict-free
constexpr int kLsMaxProK = 512;             // widest LayerNorm the prologue handles (hidden size of tracker / refiner)

struct LsParams {
  // A operand, plain / conv mode
  const __nv_bfloat16 *x;
  int64_t ldx, x_batch;
  int taps, tap_pad, tap_period, tap_len;   // row r = t * period + q reads rows clamp(t + d - pad, 0, len - 1) * period + q
  // A operand, prologue mode (x == nullptr): A = LN1(LN0(src0) + src1), K <= 512, K % 128 == 0
  const float *src0;
  const float *ln0_g, *ln0_b;
  const void *src1;
  int src1_bf16;
  const float *ln1_g, *ln1_b;
  float eps;
  float *side0;                              // optional (M, K) fp32: LN0(src0)
  float *side1;                              // optional (M, K) fp32: the A rows before rounding to bf16
  // B operand and epilogue
  const __nv_bfloat16 *w;
  int64_t w_batch;
  const float *bias;
  int64_t bias_batch;
  const float *residual;                     // optional fp32 (M, N), row stride ldr
  int64_t ldr;
  int relu;
  float *y_f32;
  __nv_bfloat16 *y_bf16;
  int64_t ldy, y_batch;
  int M, N, K;
  // epilogue LayerNorm (plain mode, N <= 512 = one full row per row block): the LAST CTA of a row block to finish normalises
  // the block's rows from y_f32:  e1 = LN_e1(y);  optionally  e2 = LN_e2(e1 + e_src1).  Outputs in f32 and bf16.
  const float *e1_g, *e1_b, *e2_g, *e2_b;
  const void *e_src1;
  int e_src1_bf16;
  float *e1_f32, *e2_f32;
  __nv_bfloat16 *e1_bf16, *e2_bf16;
  int *rowblk_cnt;                           // one arrival counter per row block, zero between launches
  int splits;                                // split-K over gridDim.z (plain mode, batch == 1): K / splits per CTA, partial tiles
  float *splitk_ws;                          // summed in split order by the LAST CTA of a tile (deterministic): (splits, M, N) f32
  int *splitk_cnt;                           // one arrival counter per output tile, zero between launches (self-cleaning)
  int stages;                                // cp.async ring depth (2..8): bytes in flight cover the ~1 us L2 latency
  long long *prof;                           // debug (DVIS_LS_PROF=1): clock64 stamps of CTA (0,0,0): start, ring filled, first
};                                           // chunk landed, k loop done, end

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// LayerNorm of RG rows at once, each held as NV float4 per lane (columns lane*4 + 128*i), two-pass like csrc/layernorm.cu.
// The rows' shuffle chains are interleaved (RG independent shuffles per step): a warp that reduced its 4 rows one after the
// other spent ~2 000 cycles in dependent shuffle latency.  gamma / beta are already in registers.
template <int RG, int NV>
__device__ __forceinline__ void ln_rows(float4 (&v)[RG][NV], int nv, int K, const float4 (&ga)[NV], const float4 (&be)[NV], float eps) {
  float s[RG], q[RG];
#pragma unroll
  for (int j = 0; j < RG; ++j) {
    s[j] = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i < nv) s[j] += (v[j][i].x + v[j][i].y) + (v[j][i].z + v[j][i].w);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1)
#pragma unroll
    for (int j = 0; j < RG; ++j) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
#pragma unroll
  for (int j = 0; j < RG; ++j) {
    s[j] /= float(K);
    q[j] = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i < nv) {
        const float a = v[j][i].x - s[j], c = v[j][i].y - s[j], d = v[j][i].z - s[j], e = v[j][i].w - s[j];
        q[j] += (a * a + c * c) + (d * d + e * e);
      }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1)
#pragma unroll
    for (int j = 0; j < RG; ++j) q[j] += __shfl_xor_sync(0xffffffffu, q[j], o);
#pragma unroll
  for (int j = 0; j < RG; ++j) {
    const float mean = s[j], rstd = rsqrtf(q[j] / float(K) + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i < nv)
        v[j][i] = make_float4((v[j][i].x - mean) * rstd * ga[i].x + be[i].x, (v[j][i].y - mean) * rstd * ga[i].y + be[i].y,
                              (v[j][i].z - mean) * rstd * ga[i].z + be[i].z, (v[j][i].w - mean) * rstd * ga[i].w + be[i].w);
  }
}

__device__ __forceinline__ void cp_async_wait_dyn(int n) {   // cp.async.wait_group takes an immediate
  switch (n) {
    case 0: cp_async_wait<0>(); break;
    case 1: cp_async_wait<1>(); break;
    case 2: cp_async_wait<2>(); break;
    case 3: cp_async_wait<3>(); break;
    case 4: cp_async_wait<4>(); break;
    case 5: cp_async_wait<5>(); break;
    default: cp_async_wait<6>(); break;
  }
}

// 8 warps.  BM x BN = 32 x 64: warps 2 x 4, warp tile 16 x 16;  64 x 64: warps 4 x 2, warp tile 16 x 32;  32 x 32 (N <= 512:
// twice the CTAs, 2/3 of the bytes per CTA -- the L2 -> SM fill is the floor of these kernels): warps 2 x 4, warp tile 16 x 8.
// Every per-thread source pointer / shared address is computed once; a k-chunk costs each thread 2-4 cp.async, 4-12 ldmatrix
// and 4-16 mma.
template <int BM, bool PRO, int BN = kLsBN>
__global__ void __launch_bounds__(kLsThreads) small_linear_kernel(const LsParams p) {
  constexpr int kLsBN = BN;                                                        // shadows the default tile width below
  constexpr int WM = BM / 16, WN = 8 / WM, WCOLS = kLsBN / WN, NT = WCOLS / 8;
  constexpr int A_TILE = BM * kLsRS, W_TILE = kLsBN * kLsRS;
  constexpr int AI = BM / 32, WI = BN / 32;                                        // A / W rows per thread and chunk
  extern __shared__ uint4 ls_smem[];
  __nv_bfloat16 *sw = reinterpret_cast<__nv_bfloat16 *>(ls_smem);                  // [stages][64][RS]
  __nv_bfloat16 *sa = sw + p.stages * W_TILE;                                      // plain: [stages][BM][RS]; PRO: [BM][K + 8]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int wm = warp / WN, wn = warp % WN;
  const int n0 = blockIdx.x * kLsBN, m0 = blockIdx.y * BM;
  const int bz = p.splits > 1 ? 0 : blockIdx.z, sz = p.splits > 1 ? blockIdx.z : 0;
  const int nk = p.K / kLsBK / p.splits, kc0 = sz * nk, S = p.stages;          // this CTA's k-chunks: kc0 .. kc0 + nk
  const int ars = PRO ? p.K + 8 : kLsRS;                                           // A row stride in smem (elements)

  // ---- per-thread copy assignments: row (tid / 8) + 32 i, 16-byte piece (tid % 8) ----
  const int lr = tid >> 3, lc = (tid & 7) * 8;
  const __nv_bfloat16 *w_src[WI];
  int w_bytes[WI];
#pragma unroll
  for (int i = 0; i < WI; ++i) {
    const int n = n0 + lr + 32 * i;
    w_bytes[i] = n < p.N ? 16 : 0;
    w_src[i] = p.w + (size_t)bz * p.w_batch + (size_t)(n < p.N ? n : 0) * p.K + lc;
  }
  const saddr_t w_dst = saddr(sw) + (lr * kLsRS + lc) * 2, a_dst = saddr(sa) + (lr * kLsRS + lc) * 2;
  const __nv_bfloat16 *a_src[AI];
  int a_bytes[AI], a_t[AI], a_q[AI];
  const int cin = PRO ? p.K : p.K / p.taps;
  if constexpr (!PRO) {
#pragma unroll
    for (int i = 0; i < AI; ++i) {
      const int m = m0 + lr + 32 * i, mm = m < p.M ? m : 0;
      a_bytes[i] = m < p.M ? 16 : 0;
      a_t[i] = p.taps > 1 ? mm / p.tap_period : 0;
      a_q[i] = p.taps > 1 ? mm - a_t[i] * p.tap_period : mm;
      a_src[i] = p.x + (size_t)bz * p.x_batch + (size_t)mm * p.ldx + lc;           // taps == 1; conv mode re-points per tap
    }
  }
  auto issue_w = [&](int kc, int slot) {
#pragma unroll
    for (int i = 0; i < WI; ++i) cp_async_16(w_dst + (slot * W_TILE + i * 32 * kLsRS) * 2, w_src[i] + (kc0 + kc) * kLsBK, w_bytes[i]);
  };
  // chunks are issued in order 0, 1, 2, ...: the Conv1d tap / column of the next chunk is tracked incrementally (no division)
  int a_col = kc0 * kLsBK, a_tap = 0;          // (split-K is not combined with taps)
  auto issue_a = [&](int slot) {
    if constexpr (!PRO) {
      if (p.taps > 1 && a_col == 0) {                                              // first chunk of a tap: re-point the source rows
#pragma unroll
        for (int i = 0; i < AI; ++i) {
          const int ts = min(max(a_t[i] + a_tap - p.tap_pad, 0), p.tap_len - 1);
          a_src[i] = p.x + (size_t)(ts * p.tap_period + a_q[i]) * p.ldx + lc;
        }
      }
#pragma unroll
      for (int i = 0; i < AI; ++i) cp_async_16(a_dst + (slot * A_TILE + i * 32 * kLsRS) * 2, a_src[i] + a_col, a_bytes[i]);
      a_col += kLsBK;
      if (a_col == cin) { a_col = 0; ++a_tap; }
    }
  };

  const bool prof = p.prof && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  if (prof) p.prof[0] = clock64();
  if constexpr (!PRO) {
    // weights first: they do not depend on the kernel before this one
    for (int s = 0; s < S - 1; ++s)
      if (s < nk) issue_w(s, s);
    pdl_wait();
    for (int s = 0; s < S - 1; ++s) {
      if (s < nk) issue_a(s);
      cp_async_commit();
    }
  } else {
    pdl_wait();                                  // the LayerNorm sources are the critical path: their loads go out first
  }

  if constexpr (PRO) {
    // ---- prologue: A rows = LN1(LN0(src0) + src1) -> bf16 in shared memory; warp w owns rows w, w+8, ... ----
    // All global loads of a group of 4 rows (both sources, gamma / beta) are issued before anything is computed: ONE L2
    // round trip per group instead of one per row and source (the first version spent 14 600 of its 20 000 cycles here).
    constexpr int NV = kLsMaxProK / 128, RG = 4;
    const int nv = p.K / 128;
    const int nseg = p.K / 64, gx = gridDim.x, my_seg = (int)blockIdx.x % min(gx, nseg);
    float4 g0[NV], b0[NV], g1[NV], b1[NV];
    bool mine[NV];                               // does this CTA store side-output column segment (lane*4 + 128 i) / 64 ?
#pragma unroll
    for (int i = 0; i < NV; ++i) mine[i] = i < nv && ((lane * 4 + 128 * i) >> 6) % gx == my_seg;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i < nv) {
        const int c = lane * 4 + 128 * i;
        if (p.ln0_g) { g0[i] = *reinterpret_cast<const float4 *>(p.ln0_g + c); b0[i] = *reinterpret_cast<const float4 *>(p.ln0_b + c); }
        if (p.ln1_g) { g1[i] = *reinterpret_cast<const float4 *>(p.ln1_g + c); b1[i] = *reinterpret_cast<const float4 *>(p.ln1_b + c); }
      }
    bool ring_issued = false;
    for (int r0 = warp; r0 < BM; r0 += 8 * RG) {
      float4 v[RG][NV], u[RG][NV];
#pragma unroll
      for (int j = 0; j < RG; ++j) {
        const int m = m0 + r0 + 8 * j;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          v[j][i] = u[j][i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < nv && m < p.M) {
            const size_t o = (size_t)m * p.K + lane * 4 + 128 * i;
            v[j][i] = *reinterpret_cast<const float4 *>(p.src0 + o);
            if (p.src1) {
              if (p.src1_bf16) {
                const uint2 w2 = *reinterpret_cast<const uint2 *>(static_cast<const __nv_bfloat16 *>(p.src1) + o);
                u[j][i] = make_float4(__uint_as_float(w2.x << 16), __uint_as_float(w2.x & 0xffff0000u), __uint_as_float(w2.y << 16),
                                      __uint_as_float(w2.y & 0xffff0000u));
              } else {
                u[j][i] = *reinterpret_cast<const float4 *>(static_cast<const float *>(p.src1) + o);
              }
            }
          }
        }
      }
      if (!ring_issued) {                        // the weight ring goes out behind the first group's row loads
        ring_issued = true;
        for (int s = 0; s < S - 1; ++s) {
          if (s < nk) issue_w(s, s);
          cp_async_commit();
        }
      }
      // rows past M hold zeros: their LayerNorm is finite garbage that is never stored
      if (p.ln0_g) ln_rows<RG, NV>(v, nv, p.K, g0, b0, p.eps);
#pragma unroll
      for (int j = 0; j < RG; ++j) {
        const int m = m0 + r0 + 8 * j;
        if (p.side0 && m < p.M) {
#pragma unroll
          for (int i = 0; i < NV; ++i)
            if (mine[i]) *reinterpret_cast<float4 *>(p.side0 + (size_t)m * p.K + lane * 4 + 128 * i) = v[j][i];
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) { v[j][i].x += u[j][i].x; v[j][i].y += u[j][i].y; v[j][i].z += u[j][i].z; v[j][i].w += u[j][i].w; }
      }
      if (p.ln1_g) ln_rows<RG, NV>(v, nv, p.K, g1, b1, p.eps);
#pragma unroll
      for (int j = 0; j < RG; ++j) {
        const int r = r0 + 8 * j, m = m0 + r;
        if (p.side1 && m < p.M) {
#pragma unroll
          for (int i = 0; i < NV; ++i)
            if (mine[i]) *reinterpret_cast<float4 *>(p.side1 + (size_t)m * p.K + lane * 4 + 128 * i) = v[j][i];
        }
#pragma unroll
        for (int i = 0; i < NV; ++i)
          if (i < nv) {
            const bool live = m < p.M;            // rows past M feed zeros to the MMA
            *reinterpret_cast<uint2 *>(sa + r * ars + lane * 4 + 128 * i) =
                live ? make_uint2(pack_bf16x2(v[j][i].x, v[j][i].y), pack_bf16x2(v[j][i].z, v[j][i].w)) : make_uint2(0u, 0u);
          }
      }
    }
  }

  float acc[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  // epilogue operands (bias, residual) are fetched now, so their L2 round trip hides behind the k loop
  const float *bias = p.bias ? p.bias + (size_t)bz * p.bias_batch : nullptr;
  float2 e_bias[NT], e_res[2][NT];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int n = n0 + wn * WCOLS + nt * 8 + 2 * t;
    e_bias[nt] = (bias && n < p.N) ? *reinterpret_cast<const float2 *>(bias + n) : make_float2(0.f, 0.f);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int m = m0 + wm * 16 + g + half * 8;
      e_res[half][nt] = (p.residual && n < p.N && m < p.M) ? *reinterpret_cast<const float2 *>(p.residual + (size_t)m * p.ldr + n)
                                                           : make_float2(0.f, 0.f);
    }
  }
  // ldmatrix addresses of this lane inside a ring slot (A: 16 x 16 tile of the warp's rows; B: pairs of 8-column tiles)
  const saddr_t a_lds = saddr(sa) + ((wm * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * ars + (lane >> 4) * 8) * 2;
  const saddr_t b_lds = NT == 1 ? saddr(sw) + ((wn * 8 + (lane & 7)) * kLsRS + (lane >> 3) * 8) * 2      // 4 k-pieces of one column tile
                                : saddr(sw) + ((wn * WCOLS + (lane >> 4) * 8 + (lane & 7)) * kLsRS + ((lane >> 3) & 1) * 8) * 2;

  // resident: the ring holds ALL of K (S - 1 >= nk): everything is already in flight, one wait + one barrier, then a pure
  // ldmatrix / mma loop -- the per-chunk wait + barrier of the streaming form cost 550 cycles a chunk with 8 warps
  const bool resident = nk <= S - 1;
  int slot = 0, fill = S - 1;                  // ring slots of chunk kc and of chunk kc + S - 1 (kept without % S)
  if (prof) p.prof[1] = clock64();
  if (resident) {
    cp_async_wait<0>();
    __syncthreads();
    if (prof) p.prof[2] = clock64();
  }
  for (int kc = 0; kc < nk; ++kc) {
    if (!resident) {
      cp_async_wait_dyn(S - 2);
      __syncthreads();                         // chunk kc has landed for everyone; the slot of chunk kc - 1 is free again
      if (prof && kc == 0) p.prof[2] = clock64();
      const int nx = kc + S - 1;
      if (nx < nk) { issue_w(nx, fill); issue_a(fill); }
      cp_async_commit();
    }
    const saddr_t ca = PRO ? a_lds + kc * (kLsBK * 2) : a_lds + slot * (A_TILE * 2);
    const saddr_t cw = b_lds + slot * (W_TILE * 2);
    if constexpr (NT == 1) {
#pragma unroll
      for (int k2 = 0; k2 < kLsBK / 32; ++k2) {
        uint32_t a0[4], a1[4], bf[4];
        ldmatrix_x4(a0, ca + k2 * 64);
        ldmatrix_x4(a1, ca + k2 * 64 + 32);
        ldmatrix_x4(bf, cw + k2 * 64);             // (8 columns) x (k .. k+31): b0, b1 of two consecutive k-steps
        mma_bf16_16816(acc[0], a0, bf[0], bf[1]);
        mma_bf16_16816(acc[0], a1, bf[2], bf[3]);
      }
    } else {
#pragma unroll
      for (int ks = 0; ks < kLsBK / 16; ++ks) {
        uint32_t af[4];
        ldmatrix_x4(af, ca + ks * 32);
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {
          uint32_t bf[4];
          ldmatrix_x4(bf, cw + (np * 16 * kLsRS) * 2 + ks * 32);
          mma_bf16_16816(acc[2 * np], af, bf[0], bf[1]);
          mma_bf16_16816(acc[2 * np + 1], af, bf[2], bf[3]);
        }
      }
    }
    slot = slot + 1 == S ? 0 : slot + 1;
    fill = fill + 1 == S ? 0 : fill + 1;
  }
  pdl_launch_dependents();
  if (prof) p.prof[3] = clock64();

  if (p.splits > 1) {
    // split-K: publish this CTA's partial tile; the LAST CTA to arrive adds the partials in split order (a fixed order: the
    // result does not depend on which CTA was last) and runs the epilogue.  No float atomics -- run-to-run determinism is a
    // property this library keeps (profiles/r2_determinism.md).
    __shared__ int s_last;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int m = m0 + wm * 16 + g + half * 8;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int n = n0 + wn * WCOLS + nt * 8 + 2 * t;
        if (m < p.M && n < p.N)
          *reinterpret_cast<float2 *>(p.splitk_ws + ((size_t)sz * p.M + m) * p.N + n) = make_float2(acc[nt][half * 2], acc[nt][half * 2 + 1]);
      }
    }
    __threadfence();
    __syncthreads();
    int *cnt = p.splitk_cnt + blockIdx.y * gridDim.x + blockIdx.x;
    if (tid == 0) s_last = atomicAdd(cnt, 1) == p.splits - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (tid == 0) *cnt = 0;                       // ready for the next launch
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int m = m0 + wm * 16 + g + half * 8;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int n = n0 + wn * WCOLS + nt * 8 + 2 * t;
        float2 sum = make_float2(0.f, 0.f);
        if (m < p.M && n < p.N)
          for (int z = 0; z < p.splits; ++z) {
            const float *part = p.splitk_ws + ((size_t)z * p.M + m) * p.N + n;   // written by other CTAs: bypass L1 (ld.cg)
            sum.x += __ldcg(part);
            sum.y += __ldcg(part + 1);
          }
        acc[nt][half * 2] = sum.x;
        acc[nt][half * 2 + 1] = sum.y;
      }
    }
  }

  // ---- epilogue: bias, ReLU, residual, stores ----
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int m = m0 + wm * 16 + g + half * 8;
    if (m >= p.M) continue;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int n = n0 + wn * WCOLS + nt * 8 + 2 * t;
      if (n >= p.N) continue;
      float v0 = acc[nt][half * 2] + e_bias[nt].x, v1 = acc[nt][half * 2 + 1] + e_bias[nt].y;
      if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
      v0 += e_res[half][nt].x;
      v1 += e_res[half][nt].y;
      const size_t o = (size_t)bz * p.y_batch + (size_t)m * p.ldy + n;
      if (p.y_f32) *reinterpret_cast<float2 *>(p.y_f32 + o) = make_float2(v0, v1);
      if (p.y_bf16) *reinterpret_cast<uint32_t *>(p.y_bf16 + o) = pack_bf16x2(v0, v1);
    }
  }
  if constexpr (!PRO && BM == 32) {
    if (p.e1_g) {
      // ---- epilogue LayerNorm by the last CTA of the row block (the producer-side form of the post-norm blocks: consumers
      // then are plain bf16 GEMMs; normalising in every consumer's prologue repeated the work N / 64 times) ----
      __shared__ int s_last_row;
      __threadfence();
      __syncthreads();
      if (tid == 0) s_last_row = atomicAdd(p.rowblk_cnt + blockIdx.y, 1) == (int)gridDim.x - 1;
      __syncthreads();
      if (s_last_row) {
        __threadfence();
        if (tid == 0) p.rowblk_cnt[blockIdx.y] = 0;
        constexpr int NV = kLsMaxProK / 128, RG = 4;
        const int nv = p.N / 128;
        float4 g1[NV], b1[NV], g2[NV], b2[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i)
          if (i < nv) {
            const int c = lane * 4 + 128 * i;
            g1[i] = *reinterpret_cast<const float4 *>(p.e1_g + c); b1[i] = *reinterpret_cast<const float4 *>(p.e1_b + c);
            if (p.e2_g) { g2[i] = *reinterpret_cast<const float4 *>(p.e2_g + c); b2[i] = *reinterpret_cast<const float4 *>(p.e2_b + c); }
          }
        float4 v[RG][NV], u[RG][NV];
#pragma unroll
        for (int j = 0; j < RG; ++j) {
          const int m = m0 + warp + 8 * j;
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            v[j][i] = u[j][i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < nv && m < p.M) {
              v[j][i] = __ldcg(reinterpret_cast<const float4 *>(p.y_f32 + (size_t)m * p.ldy + lane * 4 + 128 * i));   // other CTAs' tiles: L2
              if (p.e_src1) {
                const size_t o = (size_t)m * p.N + lane * 4 + 128 * i;
                if (p.e_src1_bf16) {
                  const uint2 w2 = *reinterpret_cast<const uint2 *>(static_cast<const __nv_bfloat16 *>(p.e_src1) + o);
                  u[j][i] = make_float4(__uint_as_float(w2.x << 16), __uint_as_float(w2.x & 0xffff0000u), __uint_as_float(w2.y << 16),
                                        __uint_as_float(w2.y & 0xffff0000u));
                } else {
                  u[j][i] = *reinterpret_cast<const float4 *>(static_cast<const float *>(p.e_src1) + o);
                }
              }
            }
          }
        }
        ln_rows<RG, NV>(v, nv, p.N, g1, b1, p.eps);
        auto store = [&](float *f32, __nv_bfloat16 *b16) {
#pragma unroll
          for (int j = 0; j < RG; ++j) {
            const int m = m0 + warp + 8 * j;
            if (m >= p.M) continue;
#pragma unroll
            for (int i = 0; i < NV; ++i)
              if (i < nv) {
                const size_t o = (size_t)m * p.N + lane * 4 + 128 * i;
                if (f32) *reinterpret_cast<float4 *>(f32 + o) = v[j][i];
                if (b16) *reinterpret_cast<uint2 *>(b16 + o) = make_uint2(pack_bf16x2(v[j][i].x, v[j][i].y), pack_bf16x2(v[j][i].z, v[j][i].w));
              }
          }
        };
        store(p.e1_f32, p.e1_bf16);
        if (p.e2_g) {
#pragma unroll
          for (int j = 0; j < RG; ++j)
#pragma unroll
            for (int i = 0; i < NV; ++i) { v[j][i].x += u[j][i].x; v[j][i].y += u[j][i].y; v[j][i].z += u[j][i].z; v[j][i].w += u[j][i].w; }
          ln_rows<RG, NV>(v, nv, p.N, g2, b2, p.eps);
          store(p.e2_f32, p.e2_bf16);
        }
      }
    }
  }
  if (prof) p.prof[4] = clock64();
}

#ifndef DVIS_SIMT_EMULATION
long long *g_ls_prof = nullptr;                // device buffer of 8 stamps when DVIS_LS_PROF is set (tests/perf only)
#endif

template <int BM, bool PRO, int BN = kLsBN>
int launch_small_linear(LsParams p, int batch, cudaStream_t s) {
#ifndef DVIS_SIMT_EMULATION
  static const bool want_prof = getenv("DVIS_LS_PROF") != nullptr;
  if (want_prof && !g_ls_prof) cudaMalloc(&g_ls_prof, 8 * sizeof(long long));
  p.prof = g_ls_prof;
#endif
  // K <= 512 (8 chunks; 6 for the 64-row tiles): the ring holds all of K (+1 slot: the loop's look-ahead index); else 6 deep
  const int nk = p.K / kLsBK / p.splits, cap = BM == 32 ? kLsMaxStages : 6;
  p.stages = nk <= cap ? nk + 1 : 6;
  if (nk > cap) p.stages = BN == 32 ? kLsMaxStages : 6;     // streaming: as many bytes in flight as fit
  const size_t smem = (size_t)p.stages * BN * kLsRS * 2 +
                      (PRO ? (size_t)BM * (p.K + 8) * 2 : (size_t)p.stages * BM * kLsRS * 2);
  auto kern = small_linear_kernel<BM, PRO, BN>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  // one carve-out for every temporal-stage kernel: consecutive launches whose shared-memory needs fall into different L1 /
  // shared splits make the SMs drain and reconfigure between them (~8 us per launch in a mixed chain, chain_probe.py)
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, p.splits > 1 ? p.splits : batch);
#ifdef DVIS_SIMT_EMULATION
  kern<<<grid, kLsThreads, smem, s>>>(p);
#else
  launch_pdl<LsParams>(kern, grid, dim3(kLsThreads), smem, s, p);
#endif
  return check_launch("small_linear_kernel");
}

}  // namespace
}  // namespace dvis

using namespace dvis;

extern "C" int dvis_linear_small_ln(const void *x, int64_t ldx, const void *w, const float *bias, const float *residual, int64_t ldr,
                                    int relu, int M, int N, int K, float *y_f32, const float *ln_gamma, const float *ln_beta, float eps,
                                    const void *src1, int src1_dtype, const float *ln2_gamma, const float *ln2_beta, float *ln_f32,
                                    void *ln_bf16, float *ln2_f32, void *ln2_bf16, float *splitk_workspace, int *splitk_counters,
                                    int *rowblock_counters, void *stream) {
  DVIS_REQUIRE(x && w && y_f32 && ln_gamma && ln_beta && rowblock_counters && (ln_f32 || ln_bf16), "linear_small_ln: null pointer argument");
  DVIS_REQUIRE(M > 0 && M <= 512 && N > 0 && N <= kLsMaxProK && N % 128 == 0 && K > 0 && K % kLsBK == 0,
               "linear_small_ln: need M <= 512, N <= 512 with N %% 128 == 0, K %% 64 == 0 (M=%d N=%d K=%d)", M, N, K);
  DVIS_REQUIRE(aligned16(x) && ldx % 8 == 0 && aligned16(w) && (!residual || ldr % 2 == 0), "linear_small_ln: operand alignment");
  DVIS_REQUIRE(!ln2_gamma == !ln2_beta && (!ln2_gamma || ln2_f32 || ln2_bf16), "linear_small_ln: second LayerNorm needs gamma, beta and an output");
  DVIS_REQUIRE(!src1 || src1_dtype == DVIS_F32 || src1_dtype == DVIS_BF16, "linear_small_ln: src1 must be f32 or bf16");
  LsParams p{};
  p.x = static_cast<const __nv_bfloat16 *>(x); p.ldx = ldx; p.taps = 1;
  p.w = static_cast<const __nv_bfloat16 *>(w); p.bias = bias; p.residual = residual; p.ldr = ldr; p.relu = relu;
  p.y_f32 = y_f32; p.ldy = N; p.M = M; p.N = N; p.K = K; p.eps = eps;
  p.e1_g = ln_gamma; p.e1_b = ln_beta; p.e2_g = ln2_gamma; p.e2_b = ln2_beta; p.e_src1 = src1; p.e_src1_bf16 = src1_dtype == DVIS_BF16;
  p.e1_f32 = ln_f32; p.e1_bf16 = static_cast<__nv_bfloat16 *>(ln_bf16); p.e2_f32 = ln2_f32; p.e2_bf16 = static_cast<__nv_bfloat16 *>(ln2_bf16);
  p.rowblk_cnt = rowblock_counters;
  p.splits = 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (splitk_workspace && splitk_counters && K >= 1024 && K % (4 * kLsBK) == 0) {
    p.splits = 4;
    p.splitk_ws = splitk_workspace;
    p.splitk_cnt = splitk_counters;
  }
  return launch_small_linear<32, false, 32>(p, 1, s);
}

extern "C" int dvis_debug_linear_small_stamps(long long *host_out) {   // tests/perf: the stamps of the last profiled launch
#ifndef DVIS_SIMT_EMULATION
  if (!g_ls_prof) return fail(DVIS_ERR_INVALID, "set DVIS_LS_PROF=1 before the first dvis_linear_small call");
  return cudaMemcpy(host_out, g_ls_prof, 8 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? DVIS_OK : DVIS_ERR_CUDA;
#else
  (void)host_out;
  return DVIS_OK;
#endif
}

extern "C" int dvis_linear_small(const void *x, int64_t ldx, int64_t x_batch, int taps, int tap_pad, int tap_period, int tap_len,
                                 const float *src0, const float *ln0_gamma, const float *ln0_beta, const void *src1,
                                 int src1_dtype, const float *ln1_gamma, const float *ln1_beta, float eps, float *side0,
                                 float *side1, const void *w, int64_t w_batch, const float *bias, int64_t bias_batch,
                                 const float *residual, int64_t ldr, int relu, float *y_f32, void *y_bf16, int64_t ldy,
                                 int64_t y_batch, int batch, int M, int N, int K, float *splitk_workspace, int *splitk_counters,
                                 void *stream) {
  DVIS_REQUIRE(w && (y_f32 || y_bf16), "linear_small: null pointer argument");
  DVIS_REQUIRE((x != nullptr) != (src0 != nullptr), "linear_small: exactly one of x (bf16 operand) and src0 (prologue) must be given");
  DVIS_REQUIRE(batch > 0 && batch <= 65535 && M > 0 && N > 0 && K > 0, "linear_small: sizes must be positive");
  DVIS_REQUIRE(K % kLsBK == 0 && N % 8 == 0, "linear_small: need K %% 64 == 0 and N %% 8 == 0 (K=%d N=%d)", K, N);
  DVIS_REQUIRE(aligned16(w) && w_batch % 8 == 0, "linear_small: weights must be 16-byte aligned");
  DVIS_REQUIRE(ldy % 2 == 0 && y_batch % 2 == 0 && (!residual || ldr % 2 == 0), "linear_small: output / residual strides must be even");
  LsParams p{};
  p.w = static_cast<const __nv_bfloat16 *>(w); p.w_batch = w_batch; p.bias = bias; p.bias_batch = bias_batch;
  p.residual = residual; p.ldr = ldr; p.relu = relu; p.y_f32 = y_f32; p.y_bf16 = static_cast<__nv_bfloat16 *>(y_bf16);
  p.ldy = ldy; p.y_batch = y_batch; p.M = M; p.N = N; p.K = K;
  p.splits = 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool big = M > 512;                      // the refiner's T*Q rows: 64-row tiles halve the weight re-reads
  if (x) {
    DVIS_REQUIRE(taps >= 1 && K % taps == 0 && (K / taps) % kLsBK == 0, "linear_small: K must split into taps x (multiple of 64)");
    DVIS_REQUIRE(taps == 1 || (tap_period > 0 && tap_len > 0 && M == tap_period * tap_len && batch == 1),
                 "linear_small: conv mode needs M == tap_period * tap_len and batch == 1");
    DVIS_REQUIRE(aligned16(x) && ldx % 8 == 0 && x_batch % 8 == 0, "linear_small: x rows must be 16-byte aligned");
    p.x = static_cast<const __nv_bfloat16 *>(x); p.ldx = ldx; p.x_batch = x_batch;
    p.taps = taps; p.tap_pad = tap_pad; p.tap_period = tap_period; p.tap_len = tap_len;
    if (big) return launch_small_linear<64, false>(p, batch, s);
    // long reductions over a narrow output (FFN2: K = 2048 -> N = 512, 200 rows): 4-way split-K, every CTA's share of K fits
    // the resident ring; needs the caller's workspace (4*M*N floats) and zeroed per-tile counters
    if (splitk_workspace && splitk_counters && taps == 1 && batch == 1 && K >= 1024 && K % (4 * kLsBK) == 0 &&
        (int64_t)((N + 31) / 32) * ((M + 31) / 32) <= 2 * kNumSMs) {
      p.splits = 4;
      p.splitk_ws = splitk_workspace;
      p.splitk_cnt = splitk_counters;
      return launch_small_linear<32, false, 32>(p, batch, s);
    }
    // narrow outputs (out-proj, FFN2: N = 512): 32-column tiles double the CTA count and cut the bytes each SM must pull
    if ((int64_t)((N + 63) / 64) * ((M + 31) / 32) * batch < kNumSMs / 2) return launch_small_linear<32, false, 32>(p, batch, s);
    return launch_small_linear<32, false>(p, batch, s);
  }
  DVIS_REQUIRE(batch == 1, "linear_small: the LayerNorm prologue is not batched");
  DVIS_REQUIRE(K <= kLsMaxProK && K % 128 == 0, "linear_small: prologue needs K <= 512 and K %% 128 == 0 (K=%d)", K);
  DVIS_REQUIRE(!ln0_gamma == !ln0_beta && !ln1_gamma == !ln1_beta, "linear_small: LayerNorm needs both gamma and beta");
  DVIS_REQUIRE(!src1 || src1_dtype == DVIS_F32 || src1_dtype == DVIS_BF16, "linear_small: src1 must be f32 or bf16");
  DVIS_REQUIRE(!side0 || ln0_gamma, "linear_small: side0 is the output of LN0");
  p.src0 = src0; p.ln0_g = ln0_gamma; p.ln0_b = ln0_beta; p.src1 = src1; p.src1_bf16 = src1_dtype == DVIS_BF16;
  p.ln1_g = ln1_gamma; p.ln1_b = ln1_beta; p.eps = eps; p.side0 = side0; p.side1 = side1;
  return big ? launch_small_linear<64, true>(p, batch, s) : launch_small_linear<32, true>(p, batch, s);
}
