// Multi-head attention core on the warp tensor cores, built for the SHORT-query attentions of the temporal stage:
//   out[b, i, h, :] = softmax_j(scale * <q[b,i,h,:], k[b,j,h,:]> [masked]) @ v[b,j,h,:]
// i.e. what nn.MultiheadAttention computes between its in- and out-projections in SelfAttentionLayer / CrossAttentionLayer
// (P/mask2former_video/modeling/transformer_decoder/video_mask2former_transformer_decoder.py:46,104),
// ReferringCrossAttentionLayer (P/dvis_Plus/tracker.py:45) and, with the bit mask, the segmenter decoder's masked
// cross-attention (P/dvis_Plus/video_mask2former_transformer_decoder.py:295-315).
//
// Q = 200 rows is a LATENCY problem (one cuDNN flash call costs ~7 us here, 150 of them per clip), so the work is cut
// for the shortest dependent chain instead of for reuse: a CTA owns 16 query rows of one (batch, head) -- 13 x 8 = 104
// CTAs for one 200-query layer -- and its 8 warps split the KEYS (flash-decoding): warp w walks key blocks w, w+8, ...
// of 32 keys with an online softmax in registers (mma.sync m16n8k16, bf16 operands, fp32 accumulate), K / V blocks
// arrive through a warp-private cp.async ring (1 or 2 stages), and the 8 partial (max, sum, O) states are merged
// through shared memory at the end.  Strides are explicit so q / k / v can be slices of a packed projection output.
#include "mma.cuh"

namespace dvis {
namespace {

constexpr int kFaWarps = 8;    // 2 warps per scheduler: with 1 the kernel is bound by dependent-issue latency
constexpr int kFaRows = 16;    // query rows per warp tile

struct FlashParams {
  const __nv_bfloat16 *q, *k, *v;
  __nv_bfloat16 *o;
  int64_t q_row, q_batch, q_head, k_row, k_batch, k_head, v_row, v_batch, v_head, o_row, o_batch;   // element strides
  const uint8_t *mask;          // optional bits: (key j of row i) set -> j may NOT be attended; shared by the heads
  int64_t mask_row, mask_batch; // bytes; mask_row % 8 == 0
  int B, Lq, Lk, H;
  float scale_log2e;
  int stages;
};

// SPLIT = true  (short key sequences): 16 query rows per CTA, the 8 warps split the keys in blocks of 32 (flash-decoding),
//                warp-private cp.async ring, the 8 partial (max, sum, O) states merged through shared memory;
// SPLIT = false (long key sequences, the predictor's 920 .. 14 720-pixel memories): NW x 16 query rows per CTA, one m16 tile
//                per warp, all warps walk every 64-key block through a CTA-wide 2-stage ring.  NW = 16 (256 rows) puts ALL
//                200 queries of a (batch, head) in one CTA: K / V are streamed once per head and 16 frames x 8 heads = 128
//                CTAs are a single wave on 148 SMs (NW = 8 needed 256 CTAs = 1.7 waves and streamed K / V twice).
template <int DH, bool SPLIT, int NW = kFaWarps>
__global__ void __launch_bounds__(NW * 32) flash_attn_kernel(const FlashParams p) {
  constexpr int kFaWarps = NW;                     // warps of this instantiation (shadows the default)
  constexpr int KB = SPLIT ? 32 : 64;              // keys per block
  constexpr int RS = DH + 8;                       // padded row stride (bf16): 16-byte aligned, ldmatrix conflict-free
  constexpr int TILE = KB * RS;                    // elements of one K (or V) block
  constexpr int KSTEPS = DH / 16, DT = DH / 8, NKT = KB / 8;
  constexpr int CH = DH / 8;                       // 16-byte pieces per row
  constexpr int LDT = SPLIT ? 32 : kFaWarps * 32;  // threads copying one block
  constexpr int RPP = LDT / CH;                    // rows copied per pass
  constexpr int PASSES = (KB + RPP - 1) / RPP;     // passes per block (the last pass may be partial: 16 warps, 64 keys)
  extern __shared__ uint4 fa_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y;
  const int row0 = SPLIT ? blockIdx.x * kFaRows : (blockIdx.x * kFaWarps + warp) * kFaRows;   // this WARP's first query row
  __nv_bfloat16 *ring = reinterpret_cast<__nv_bfloat16 *>(fa_smem) + (SPLIT ? (size_t)warp * p.stages * 2 * TILE : 0);
  const __nv_bfloat16 *kb = p.k + (size_t)b * p.k_batch + (size_t)h * p.k_head;
  const __nv_bfloat16 *vb = p.v + (size_t)b * p.v_batch + (size_t)h * p.v_head;
  const int nblocks = (p.Lk + KB - 1) / KB;
  const int blk0 = SPLIT ? warp : 0, blk_step = SPLIT ? kFaWarps : 1;

  // per-thread copy assignment, computed once: row (ld / CH) + RPP * pass, piece (ld % CH)
  const int ld = SPLIT ? lane : (int)threadIdx.x, lr = ld / CH, lc = (ld % CH) * 8;
  const saddr_t k_dst = saddr(ring) + (lr * RS + lc) * 2;
  auto issue = [&](int blk, int stage) {           // this warp's (SPLIT) / CTA's copy of key block `blk` into ring slot `stage`
    const int key0 = blk * KB + lr;
#pragma unroll
    for (int i = 0; i < PASSES; ++i) {
      if (lr + i * RPP >= KB) break;               // more copying threads than 16-byte pieces in a block
      const int key = key0 + i * RPP;
      const bool ok = key < p.Lk;                  // rows past Lk are zero-filled so that 0-probability x V stays 0
      const size_t kr = ok ? (size_t)key : 0;
      const saddr_t d = k_dst + (stage * 2 * TILE + i * RPP * RS) * 2;
      cp_async_16(d, kb + kr * p.k_row + lc, ok ? 16 : 0);
      cp_async_16(d + TILE * 2, vb + kr * p.v_row + lc, ok ? 16 : 0);
    }
    cp_async_commit();
  };

  pdl_wait();
  if (blk0 < nblocks) issue(blk0, 0);

  // Q fragments of rows (g, g+8) straight from global memory
  uint32_t qa[KSTEPS][4];
  {
    const int r0 = row0 + g, r1 = row0 + g + 8;
    const __nv_bfloat16 *q0 = p.q + (size_t)b * p.q_batch + (size_t)h * p.q_head + (size_t)min(r0, p.Lq - 1) * p.q_row + 2 * t;
    const __nv_bfloat16 *q1 = p.q + (size_t)b * p.q_batch + (size_t)h * p.q_head + (size_t)min(r1, p.Lq - 1) * p.q_row + 2 * t;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
      qa[ks][0] = *reinterpret_cast<const uint32_t *>(q0 + ks * 16);
      qa[ks][1] = *reinterpret_cast<const uint32_t *>(q1 + ks * 16);
      qa[ks][2] = *reinterpret_cast<const uint32_t *>(q0 + ks * 16 + 8);
      qa[ks][3] = *reinterpret_cast<const uint32_t *>(q1 + ks * 16 + 8);
    }
  }

  float o[DT][4];
#pragma unroll
  for (int i = 0; i < DT; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
  const uint8_t *mk0 = nullptr, *mk1 = nullptr;
  if (p.mask) {
    mk0 = p.mask + (size_t)b * p.mask_batch + (size_t)min(row0 + g, p.Lq - 1) * p.mask_row;
    mk1 = p.mask + (size_t)b * p.mask_batch + (size_t)min(row0 + g + 8, p.Lq - 1) * p.mask_row;
  }
  // this lane's ldmatrix addresses inside a ring slot: K as the col-major B operand of S = Q K^T, V (transposed load) of O = P V
  const saddr_t k_lds = saddr(ring) + (((lane >> 4) * 8 + (lane & 7)) * RS + ((lane >> 3) & 1) * 8) * 2;
  const saddr_t v_lds = saddr(ring) + (TILE + (((lane >> 3) & 1) * 8 + (lane & 7)) * RS + (lane >> 4) * 8) * 2;

  // the mask words of a block are fetched one block ahead: their L2 round trip must not sit between S and the softmax
  uint64_t nb0 = 0, nb1 = 0;
  if (p.mask && blk0 < nblocks) {
    nb0 = *reinterpret_cast<const uint64_t *>(mk0 + (((blk0 * KB) >> 6) << 3));
    nb1 = *reinterpret_cast<const uint64_t *>(mk1 + (((blk0 * KB) >> 6) << 3));
  }
  int it = 0;
  for (int blk = blk0; blk < nblocks; blk += blk_step, ++it) {
    const int stage = p.stages == 2 ? (it & 1) : 0;
    const bool more = blk + blk_step < nblocks;
    if (p.stages == 2 && more) {
      issue(blk + blk_step, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    if constexpr (SPLIT) __syncwarp(); else __syncthreads();
    const saddr_t sk = k_lds + stage * (2 * TILE * 2), sv = v_lds + stage * (2 * TILE * 2);
    const int key0 = blk * KB;
    const int nkeys = min(KB, p.Lk - key0);
    const int nkt = (nkeys + 7) >> 3;              // 8-key tiles that hold at least one valid key

    // ---- S = Q K^T ----
    float s[NKT][4];
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
    for (int np = 0; np < NKT / 2; ++np) {         // pairs of 8-key tiles
      if (2 * np < nkt) {
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
          uint32_t bf[4];
          ldmatrix_x4(bf, sk + (np * 16 * RS) * 2 + ks * 32);
          mma_bf16_16816(s[2 * np], qa[ks], bf[0], bf[1]);
          mma_bf16_16816(s[2 * np + 1], qa[ks], bf[2], bf[3]);
        }
      }
    }
    // ---- scale, mask, online softmax ----
    // 64-bit words hold 64 keys; a 32-key block reads its half.  Keys past Lk count as masked: fold them into the words
    // once per block instead of comparing every element
    uint64_t bits0 = nb0 >> (key0 & 63), bits1 = nb1 >> (key0 & 63);
    if (nkeys < KB) {
      const uint64_t tail = ~uint64_t(0) << nkeys;
      bits0 |= tail;
      bits1 |= tail;
    }
    if (p.mask && more) {
      const int nk0 = (blk + blk_step) * KB;
      nb0 = *reinterpret_cast<const uint64_t *>(mk0 + ((nk0 >> 6) << 3));
      nb1 = *reinterpret_cast<const uint64_t *>(mk1 + ((nk0 >> 6) << 3));
    }
    const uint32_t lo0 = uint32_t(bits0 >> (2 * t)), hi0 = uint32_t(bits0 >> (32 + 2 * t));   // this lane's columns 2t, 2t+1 (+8 nt)
    const uint32_t lo1 = uint32_t(bits1 >> (2 * t)), hi1 = uint32_t(bits1 >> (32 + 2 * t));
    float mx0 = mrow[0], mx1 = mrow[1];
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt) {
      const uint32_t w0 = (nt < 4 ? lo0 : hi0) >> ((nt & 3) * 8), w1 = (nt < 4 ? lo1 : hi1) >> ((nt & 3) * 8);
      s[nt][0] = (w0 & 1) ? -INFINITY : s[nt][0] * p.scale_log2e;
      s[nt][1] = (w0 & 2) ? -INFINITY : s[nt][1] * p.scale_log2e;
      s[nt][2] = (w1 & 1) ? -INFINITY : s[nt][2] * p.scale_log2e;
      s[nt][3] = (w1 & 2) ? -INFINITY : s[nt][3] * p.scale_log2e;
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float use0 = mx0 == -INFINITY ? 0.f : mx0, use1 = mx1 == -INFINITY ? 0.f : mx1;   // rows with nothing open yet
    const float al0 = exp2f(mrow[0] - use0), al1 = exp2f(mrow[1] - use1);
    mrow[0] = mx0;
    mrow[1] = mx1;
    lrow[0] *= al0;
    lrow[1] *= al1;
#pragma unroll
    for (int i = 0; i < DT; ++i) {
      o[i][0] *= al0; o[i][1] *= al0; o[i][2] *= al1; o[i][3] *= al1;
    }
    uint32_t pa[NKT / 2][4];                       // P as A fragments: k-step kk covers keys kk*16 .. kk*16+15
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt) {
      const float p0 = exp2f(s[nt][0] - use0), p1 = exp2f(s[nt][1] - use0);
      const float p2 = exp2f(s[nt][2] - use1), p3 = exp2f(s[nt][3] - use1);
      lrow[0] += p0 + p1;
      lrow[1] += p2 + p3;
      pa[nt >> 1][(nt & 1) * 2] = pack_bf16x2(p0, p1);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < NKT / 2; ++kk) {
      if (2 * kk < nkt) {
#pragma unroll
        for (int dp = 0; dp < DT / 2; ++dp) {      // pairs of 8-wide output column tiles
          uint32_t bf[4];
          ldmatrix_x4_trans(bf, sv + (kk * 16 * RS) * 2 + dp * 32);
          mma_bf16_16816(o[2 * dp], pa[kk], bf[0], bf[1]);
          mma_bf16_16816(o[2 * dp + 1], pa[kk], bf[2], bf[3]);
        }
      }
    }
    if constexpr (SPLIT) __syncwarp(); else __syncthreads();   // everyone is done with this ring slot before it is refilled
  }
  pdl_launch_dependents();

  lrow[0] += __shfl_xor_sync(0xffffffffu, lrow[0], 1);
  lrow[0] += __shfl_xor_sync(0xffffffffu, lrow[0], 2);
  lrow[1] += __shfl_xor_sync(0xffffffffu, lrow[1], 1);
  lrow[1] += __shfl_xor_sync(0xffffffffu, lrow[1], 2);
  if constexpr (!SPLIT) {
    // each warp owns its 16 rows outright: normalise and store from the fragments
    const float i0 = lrow[0] > 0.f ? 1.f / lrow[0] : 0.f, i1 = lrow[1] > 0.f ? 1.f / lrow[1] : 0.f;
    __nv_bfloat16 *ob = p.o + (size_t)b * p.o_batch + (size_t)h * DH + 2 * t;
#pragma unroll
    for (int i = 0; i < DT; ++i) {
      if (row0 + g < p.Lq) *reinterpret_cast<uint32_t *>(ob + (size_t)(row0 + g) * p.o_row + i * 8) = pack_bf16x2(o[i][0] * i0, o[i][1] * i0);
      if (row0 + g + 8 < p.Lq)
        *reinterpret_cast<uint32_t *>(ob + (size_t)(row0 + g + 8) * p.o_row + i * 8) = pack_bf16x2(o[i][2] * i1, o[i][3] * i1);
    }
    return;
  }
  // ---- SPLIT: merge the 8 warps' partial states ----
  constexpr int OS = DH + 4;                       // fp32 row stride of the merge buffer
  float *mo = reinterpret_cast<float *>(ring);     // this warp's own ring memory: [16][OS] then m[16], l[16]
  float *mm = mo + kFaRows * OS, *ml = mm + kFaRows;
#pragma unroll
  for (int i = 0; i < DT; ++i) {
    *reinterpret_cast<float2 *>(mo + g * OS + i * 8 + 2 * t) = make_float2(o[i][0], o[i][1]);
    *reinterpret_cast<float2 *>(mo + (g + 8) * OS + i * 8 + 2 * t) = make_float2(o[i][2], o[i][3]);
  }
  if (t == 0) {
    mm[g] = mrow[0]; mm[g + 8] = mrow[1];
    ml[g] = lrow[0]; ml[g + 8] = lrow[1];
  }
  __syncthreads();
  const size_t wstride = (size_t)p.stages * 2 * TILE * sizeof(__nv_bfloat16) / sizeof(float);   // floats between warps' regions
  const float *base = reinterpret_cast<const float *>(fa_smem);
  for (int e = threadIdx.x; e < kFaRows * (DH / 2); e += kFaWarps * 32) {
    const int r = e / (DH / 2), c = (e - r * (DH / 2)) * 2;
    if (blockIdx.x * kFaRows + r >= p.Lq) continue;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < kFaWarps; ++w) M = fmaxf(M, base[w * wstride + kFaRows * OS + r]);
    float L = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int w = 0; w < kFaWarps; ++w) {
      const float *wb = base + w * wstride;
      const float mw = wb[kFaRows * OS + r];
      const float f = mw == -INFINITY ? 0.f : exp2f(mw - M);
      L += wb[kFaRows * OS + kFaRows + r] * f;
      a0 += wb[r * OS + c] * f;
      a1 += wb[r * OS + c + 1] * f;
    }
    const float inv = L > 0.f ? 1.f / L : 0.f;   // a row with every key masked (the mask producer never emits one) -> 0
    __nv_bfloat16 *op = p.o + (size_t)b * p.o_batch + (size_t)(blockIdx.x * kFaRows + r) * p.o_row + (size_t)h * DH + c;
    *reinterpret_cast<uint32_t *>(op) = pack_bf16x2(a0 * inv, a1 * inv);
  }
}

template <int DH, bool SPLIT, int NW = kFaWarps>
int launch_flash(const FlashParams &p, cudaStream_t s) {
  const int kb = SPLIT ? 32 : 64;
  const size_t ring = (size_t)(SPLIT ? NW : 1) * p.stages * 2 * kb * (DH + 8) * sizeof(__nv_bfloat16);
  auto kern = flash_attn_kernel<DH, SPLIT, NW>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(ring));
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);   // see small_linear.cu
  const int rows = SPLIT ? kFaRows : kFaRows * NW;
  dim3 grid((p.Lq + rows - 1) / rows, p.H, p.B);
#ifdef DVIS_SIMT_EMULATION
  kern<<<grid, NW * 32, ring, s>>>(p);
#else
  launch_pdl<FlashParams>(kern, grid, dim3(NW * 32), ring, s, p);
#endif
  return check_launch("flash_attn_kernel");
}

}  // namespace
}  // namespace dvis

using namespace dvis;

extern "C" int dvis_flash_attn(const void *q, int64_t q_row, int64_t q_batch, int64_t q_head, const void *k, int64_t k_row,
                               int64_t k_batch, int64_t k_head, const void *v, int64_t v_row, int64_t v_batch, int64_t v_head,
                               void *out, int64_t o_row, int64_t o_batch, const void *mask_bits, int64_t mask_row_bytes,
                               int64_t mask_batch_bytes, int B, int Lq, int Lk, int H, int Dh, float scale, void *stream) {
  DVIS_REQUIRE(q && k && v && out, "flash_attn: null pointer argument");
  DVIS_REQUIRE(B > 0 && Lq > 0 && Lk > 0 && H > 0, "flash_attn: sizes must be positive");
  DVIS_REQUIRE(B <= 65535 && H <= 65535, "flash_attn: batch / heads too large for the grid");
  if (Dh != 32 && Dh != 64) return fail(DVIS_ERR_UNSUPPORTED, "flash_attn: head dim %d (built: 32, 64)", Dh);
  const auto m8 = [](int64_t s) { return (s & 7) == 0; };
  DVIS_REQUIRE(aligned16(k) && aligned16(v) && m8(k_row) && m8(k_batch) && m8(k_head) && m8(v_row) && m8(v_batch) && m8(v_head),
               "flash_attn: k / v rows must be 16-byte aligned (pointer and strides multiples of 8 elements)");
  const auto even = [](int64_t s) { return (s & 1) == 0; };
  DVIS_REQUIRE((reinterpret_cast<uintptr_t>(q) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0 && even(q_row) &&
                   even(q_batch) && even(q_head) && even(o_row) && even(o_batch),
               "flash_attn: q / out must be 4-byte aligned (pointer and strides even)");
  if (mask_bits) {
    DVIS_REQUIRE((reinterpret_cast<uintptr_t>(mask_bits) & 7) == 0 && m8(mask_row_bytes) && m8(mask_batch_bytes) &&
                     mask_row_bytes >= ((int64_t)(Lk + 63) / 64) * 8,
                 "flash_attn: mask rows must be 8-byte aligned and hold ceil(Lk / 64) * 8 bytes");
  }
  FlashParams p{static_cast<const __nv_bfloat16 *>(q), static_cast<const __nv_bfloat16 *>(k), static_cast<const __nv_bfloat16 *>(v),
                static_cast<__nv_bfloat16 *>(out), q_row, q_batch, q_head, k_row, k_batch, k_head, v_row, v_batch, v_head, o_row,
                o_batch, static_cast<const uint8_t *>(mask_bits), mask_row_bytes, mask_batch_bytes, B, Lq, Lk, H,
                scale * 1.4426950408889634f, 1};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // Tiling, from measurements on the B200 (tests/perf/flash_long_memory_probe.py, flash_variant_probe.py;
  // profiles/r2_flash_long_memory_probe.json):
  //  * long memories (the predictor's masked cross-attention over 920 / 3 680 / 14 720 pixels): 64-ROW tiles (4 warps, every warp
  //    walks every key block, K / V streamed once per CTA) from 48 (batch, head) pairs on -- 425 us at 16 frames x 14 720 keys vs
  //    469 us for 128-row tiles (2 CTAs per head: 256 CTAs leave the 148 SMs unevenly loaded) and 430 us for cuDNN with a dense
  //    bias; below that (4 and 2 frames per rank: 4 / 8 GPUs) the KEY-SPLIT variant (16-row CTAs, the 8 warps split the keys):
  //    141 vs 182 us at 4 frames, 93 vs 182 us at 2;
  //  * short memories: key split, except big batches of 200-query problems, where 16-row CTAs would re-stream K / V 13x per head
  //    (refiner: 1 664 CTAs, 30 us; cuDNN 8 us): 128-row tiles.
  const int64_t split_ctas = (int64_t)((Lq + kFaRows - 1) / kFaRows) * H * B;
  const int64_t bh = (int64_t)H * B;
  const bool long_mem = Lk > 512;
  const char *force = getenv("DVIS_FLASH_VARIANT");   // tests / experiments: 1 = 128-row, 2 = 64-row, 4 = 32-row tiles, 3 = key split
  const int forced = force ? atoi(force) : 0;
  if (forced == 2 || (forced == 0 && long_mem && bh >= 48)) {
    p.stages = 2;
    return Dh == 32 ? launch_flash<32, false, 4>(p, s) : launch_flash<64, false, 4>(p, s);
  }
  if (forced == 4) {
    p.stages = 2;
    return Dh == 32 ? launch_flash<32, false, 2>(p, s) : launch_flash<64, false, 2>(p, s);
  }
  if (forced == 1 || (forced == 0 && !long_mem && Lq > 64 && split_ctas > 2 * kNumSMs)) {
    p.stages = 2;
    // (16 warps = all 200 queries of a head in one CTA, K / V streamed once: measured SLOWER -- 500 us vs 465 us at 14 720 keys)
    return Dh == 32 ? launch_flash<32, false>(p, s) : launch_flash<64, false>(p, s);
  }
  p.stages = (Lk + 31) / 32 > kFaWarps ? 2 : 1;    // a warp with more than one 32-key block prefetches the next one
  return Dh == 32 ? launch_flash<32, true>(p, s) : launch_flash<64, true>(p, s);
}
