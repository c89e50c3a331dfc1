// Token-wise linear layers of the MSDeformAttn encoder on the 5th-gen tensor cores, with the epilogues the path needs:
//
//   y[r, :] = x[r, :] . W^T + bias          x (rows, K) bf16, W (N, K) bf16 (nn.Linear layout = K-major, what UMMA wants), N <= 256
//
//   DVIS_LINEAR_PLAIN   bf16 (rows, N) row-major, optional ReLU
//   DVIS_LINEAR_HEADS   the value projection of MSDeformAttn (OPS/modules/ms_deform_attn.py:98-101): bf16, written HEAD-MAJOR
//                       (batch, N/32, S, 32) -- row r is token s = r % S of batch item r / S -- with masked rows zeroed
//                       (py:99-100).  This is the layout csrc/msda_forward.cu's head-major gather reads: the two x-adjacent
//                       bilinear corners are one 128-byte line.  No transposition pass: the epilogue writes it directly.
//   DVIS_LINEAR_ADD_LN  the attention output projection followed by the post-norm residual block
//                       (py:118 + msdeformattn.py:118-119): LayerNorm(residual + y) * gamma + beta over the N columns, written
//                       as f32 (the residual stream), bf16 (what the next GEMM reads) and bf16 (+ pos) (the next layer's query):
//                       replaces one cuBLAS GEMM + one add_layernorm pass (the projection's output never reaches HBM).
//
// Shape of the problem: K = N = 256, rows = frames x 19 320 tokens: HBM-bound like the mask GEMM (csrc/mask_gemm.cu), whose
// skeleton this reuses: persistent CTAs, warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = epilogue (two groups of 4, one per accumulator); W resident in shared
// memory (128 KB), x streamed in 128-row x 64-column SWIZZLE_128B blocks, the 128 x N fp32 accumulator double-buffered in TMEM.
// Epilogue: TMEM lane = row, so a thread owns one row of the tile and all N columns of it -- the LayerNorm statistics need NO
// cross-thread reduction (one pass adds bias + residual, writes the row back with tcgen05.st and accumulates shifted sums; a
// second pass over TMEM normalises, so the residual is read once).  Global accesses go through a warp-private 32 x 32 transposition tile so that
// every load / store instruction covers whole 128-byte (f32) or 64-byte (bf16) row segments.
#include <algorithm>

#include "common.cuh"
#include "tc05.cuh"

namespace dvis {
namespace {

using namespace tc05;

constexpr int kTileM = 128;
constexpr int kBlockK = 64;
constexpr int kStageBytes = kTileM * 128;
constexpr int kMaxStages = 6;
constexpr int kEpiGroups = 2;                     // one group of 4 epilogue warps per TMEM accumulator: two tiles drain concurrently
constexpr int kEpiWarps = 4 * kEpiGroups;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kAccCols = 256;
constexpr int kXposeBytes = 32 * 32 * 4;          // one 32 x 32 f32 transposition tile per epilogue warp (XOR-swizzled rows)
constexpr int kSmemLimit = 227 * 1024;            // opt-in dynamic shared memory per CTA on sm_100

enum { kPlain = 0, kHeads = 1, kAddLn = 2 };

struct LinearTcParams {
  int rows, N, KB, stages, total_tiles;
  const float *bias;            // (N,) f32 or null
  int relu;
  // kPlain / kHeads
  __nv_bfloat16 *y;
  int64_t ldy;                  // kPlain: elements between rows
  int S, heads;                 // kHeads
  const uint8_t *row_mask;      // kHeads: 1 = zero the row (input_padding_mask), or null
  // kAddLn
  const float *residual;        // (rows, N) f32
  const float *gamma, *beta;    // (N,) f32
  float eps;
  float *out_f32;               // each optional
  __nv_bfloat16 *out_lp, *out_lp_pos;
  const float *pos;             // (pos_rows, N) f32
  int pos_rows;
};

struct __align__(8) Barriers {
  uint64_t full[kMaxStages], empty[kMaxStages];
  uint64_t b_full;
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const LinearTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by OFFSET arithmetic on the __shared__ array: a round trip through uintptr_t makes the compiler forget the
  // address space and every staging access becomes a generic LD / ST (ncu: 28 % of the stall samples of the first version)
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int b_block_bytes = p.N * 128;                          // one W k-block: N rows x 128 B
  uint8_t *sB = smem;
  uint8_t *sA = smem + p.KB * b_block_bytes;                    // 1024-aligned: N % 8 == 0
  float *sXpose = reinterpret_cast<float *>(sA + p.stages * kStageBytes);          // warp-private transposition tiles (4 KB each)
  float *sBias = sXpose + kEpiWarps * (kXposeBytes / 4);                            // N floats (kPlain / kHeads only)
  Barriers *bars = reinterpret_cast<Barriers *>(sBias + (MODE == kAddLn ? 0 : 256));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_x);
    prefetch_tensormap(&tmap_w);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    mbar_init(&bars->b_full, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&bars->acc_full[a], 1); mbar_init(&bars->acc_empty[a], 4); }   // 4 warps of the owning group
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
  if (MODE != kAddLn)
    for (int c = threadIdx.x; c < 256; c += kThreads) sBias[c] = (p.bias && c < p.N) ? p.bias[c] : 0.f;
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem_base = bars->tmem_base;

  const int tile_begin = int((int64_t)p.total_tiles * blockIdx.x / gridDim.x);
  const int tile_end = int((int64_t)p.total_tiles * (blockIdx.x + 1) / gridDim.x);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && tile_begin < tile_end) {
      mbar_arrive_expect_tx(&bars->b_full, uint32_t(p.KB * b_block_bytes));
      for (int kb = 0; kb < p.KB; ++kb) tma_load_3d(sB + kb * b_block_bytes, &tmap_w, &bars->b_full, kb * kBlockK, 0, 0);
      int stage = 0, phase = 0;
      for (int t = tile_begin; t < tile_end; ++t) {
        for (int kb = 0; kb < p.KB; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&bars->full[stage], kStageBytes);
          tma_load_3d(sA + stage * kStageBytes, &tmap_x, &bars->full[stage], kb * kBlockK, t * kTileM, 0);   // rows past the end: zeros
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && tile_begin < tile_end) {
      const uint32_t idesc = make_idesc(kTileM, p.N, /*BF16*/ 1);
      int stage = 0, phase = 0, n_tile = 0;
      mbar_wait(&bars->b_full, 0);
      for (int t = tile_begin; t < tile_end; ++t, ++n_tile) {
        const int acc = n_tile & 1;
        mbar_wait(&bars->acc_empty[acc], ((n_tile >> 1) & 1) ^ 1);
        fence_after_thread_sync();
        const uint32_t d_tmem = tmem_base + acc * kAccCols;
        for (int kb = 0; kb < p.KB; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          fence_after_thread_sync();
          const uint32_t a_addr = smem_u32(sA + stage * kStageBytes);
          const uint32_t b_addr = smem_u32(sB + kb * b_block_bytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            mma_bf16_ss(d_tmem, make_desc_k_sw128(a_addr + k * 32), make_desc_k_sw128(b_addr + k * 32), idesc, uint32_t(kb | k));
          mma_commit(&bars->empty[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        mma_commit(&bars->acc_full[acc]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // Only one epilogue warp per scheduler would run at a fraction of the issue rate (every dependent latency exposed), so the
    // two accumulators are drained by two independent groups of 4 warps: group g takes the tiles with n_tile % 2 == g.
    const int quarter = warp & 3;                          // TMEM lanes [32*quarter, +32) belong to this warp
    const int group = (warp - 2) >> 2;
    float *xp = sXpose + (warp - 2) * (kXposeBytes / 4);   // warp-private transposition tile
    for (int n_tile = group, t = tile_begin + group; t < tile_end; t += kEpiGroups, n_tile += kEpiGroups) {
      const int acc = n_tile & 1;
      mbar_wait(&bars->acc_full[acc], (n_tile >> 1) & 1);
      fence_after_thread_sync();
      const uint32_t taddr = tmem_base + acc * kAccCols + (uint32_t(quarter * 32) << 16);
      const int64_t row0 = (int64_t)t * kTileM + quarter * 32;       // first of this warp's 32 rows

      if constexpr (MODE == kPlain || MODE == kHeads) {
        // reader mapping for a 32 x 32 bf16 block: lane -> (row = it * 8 + lane / 4, 8 columns at (lane % 4) * 8)
        const int rr = lane >> 2, c8 = lane & 3;
        uint4 *xp16 = reinterpret_cast<uint4 *>(xp);                  // rows of 80 bytes (5 x uint4; 4 used)
        // element offsets of this lane's 4 rows (it = 0..3), once per tile; -1 = row past the end
        int64_t base[4];
        bool zero[4];
        {
          int64_t n0 = 0;
          int s0 = 0;
          if constexpr (MODE == kHeads) { n0 = row0 / p.S; s0 = int(row0 - n0 * p.S); }   // (batch item, token) of the warp's first row
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int row = it * 8 + rr;
            const int64_t gr = row0 + row;
            zero[it] = false;
            if (gr >= p.rows) { base[it] = -1; continue; }
            if constexpr (MODE == kHeads) {
              int64_t n = n0;
              int s = s0 + row;
              while (s >= p.S) { s -= p.S; ++n; }
              base[it] = ((n * p.heads * p.S + s) << 5) + c8 * 8;
              zero[it] = p.row_mask && p.row_mask[gr];
            } else {
              base[it] = gr * p.ldy + c8 * 8;
            }
          }
        }
        const int64_t chunk_step = MODE == kHeads ? ((int64_t)p.S << 5) : 32;            // one head slab / 32 columns
        for (int c0 = 0; c0 < p.N; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c0, r);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 bb = *reinterpret_cast<const float2 *>(sBias + c0 + 2 * i);
            float a = __uint_as_float(r[2 * i]) + bb.x, b = __uint_as_float(r[2 * i + 1]) + bb.y;
            if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
            pk[i] = pack_bf16x2(a, b);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) xp16[lane * 5 + k] = make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
          __syncwarp();
          const int64_t coff = (c0 >> 5) * chunk_step;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            if (base[it] >= 0) {
              uint4 v = xp16[(it * 8 + rr) * 5 + c8];
              if (zero[it]) v = make_uint4(0u, 0u, 0u, 0u);
              *reinterpret_cast<uint4 *>(p.y + base[it] + coff) = v;
            }
          }
          __syncwarp();
        }
      } else {
        // reader mapping for a 32 x 32 f32 block: lane -> (row = it * 4 + lane / 8, 4 columns at (lane % 8) * 4).  The tile is
        // stored as rows of 8 float4 with the float4 index XOR-ed by (row & 7): both the reader mapping (8 lanes = one row) and
        // the owner mapping (8 lanes = 8 consecutive rows, same float4) touch 8 distinct 16-byte bank groups.
        const int rr = lane >> 3, c4 = lane & 7;
        float4 *xq = reinterpret_cast<float4 *>(xp);
        const float invN = 1.f / float(p.N);
        const int row0i = int(row0);
        const int pr0 = p.out_lp_pos ? int(row0 % p.pos_rows) : 0;
        int roff[8];                                       // element offset of (row, c4 * 4) for this lane's 8 reader rows; -1 past the end
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int gr = row0i + it * 4 + rr;
          roff[it] = gr < p.rows ? gr * p.N + c4 * 4 : -1;
        }
        auto load_chunk = [&](float4 (&buf)[8], int c0) {
#pragma unroll
          for (int it = 0; it < 8; ++it)
            buf[it] = roff[it] >= 0 ? __ldg(reinterpret_cast<const float4 *>(p.residual + roff[it] + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        // pass 1: v = acc + bias + residual, written back to TMEM, with the row's shifted sums (shift = the row's first value:
        // var = E[(v - c)^2] - (E[v - c])^2 is then as stable as the two-pass form for any mean).  The residual block is read in
        // the coalesced reader layout TWO chunks ahead of its use (whole lines per instruction: the ~28 KB of L1 left beside the
        // shared memory cannot hold partially consumed lines) and handed to the row owners through the transposition tile.
        float4 pa[8], pb[8];
        load_chunk(pa, 0);
        if (p.N > 32) load_chunk(pb, 32);
        float shift = 0.f, s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
        auto pass1_chunk = [&](float4 (&buf)[8], int c0) {
          const float4 b4 = p.bias ? __ldg(reinterpret_cast<const float4 *>(p.bias + c0 + c4 * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int it = 0; it < 8; ++it)
            xq[(it * 4 + rr) * 8 + (c4 ^ ((it * 4 + rr) & 7))] =
                make_float4(buf[it].x + b4.x, buf[it].y + b4.y, buf[it].z + b4.z, buf[it].w + b4.w);
          if (c0 + 64 < p.N) load_chunk(buf, c0 + 64);
          __syncwarp();
          uint32_t r[32];
          tmem_ld_32x32(taddr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 a = xq[lane * 8 + (k ^ (lane & 7))];
            const float v0 = __uint_as_float(r[4 * k]) + a.x, v1 = __uint_as_float(r[4 * k + 1]) + a.y;
            const float v2 = __uint_as_float(r[4 * k + 2]) + a.z, v3 = __uint_as_float(r[4 * k + 3]) + a.w;
            if (c0 == 0 && k == 0) shift = v0;
            const float d0 = v0 - shift, d1 = v1 - shift, d2 = v2 - shift, d3 = v3 - shift;
            s0 += d0 + d1; s1 += d2 + d3;
            q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q0 = fmaf(d2, d2, q0); q1 = fmaf(d3, d3, q1);
            r[4 * k] = __float_as_uint(v0); r[4 * k + 1] = __float_as_uint(v1);
            r[4 * k + 2] = __float_as_uint(v2); r[4 * k + 3] = __float_as_uint(v3);
          }
          tmem_st_32x32(taddr + c0, r);
          __syncwarp();
        };
        for (int c0 = 0; c0 < p.N; c0 += 64) {
          pass1_chunk(pa, c0);
          if (c0 + 32 < p.N) pass1_chunk(pb, c0 + 32);
        }
        tmem_st_wait();
        const float md = (s0 + s1) * invN;                              // mean - shift
        const float mean = shift + md;
        const float rstd = rsqrtf(fmaxf((q0 + q1) * invN - md * md, 0.f) + p.eps);
        // pass 2: normalise, transpose, scale / shift per column, write
        for (int c0 = 0; c0 < p.N; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 8; ++k)
            xq[lane * 8 + (k ^ (lane & 7))] =
                make_float4((__uint_as_float(r[4 * k]) - mean) * rstd, (__uint_as_float(r[4 * k + 1]) - mean) * rstd,
                            (__uint_as_float(r[4 * k + 2]) - mean) * rstd, (__uint_as_float(r[4 * k + 3]) - mean) * rstd);
          __syncwarp();
          const float4 g4 = __ldg(reinterpret_cast<const float4 *>(p.gamma + c0 + c4 * 4));
          const float4 e4 = __ldg(reinterpret_cast<const float4 *>(p.beta + c0 + c4 * 4));
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            if (roff[it] >= 0) {
              const int row = it * 4 + rr;
              const float4 nv = xq[row * 8 + (c4 ^ (row & 7))];
              const float4 y = make_float4(fmaf(nv.x, g4.x, e4.x), fmaf(nv.y, g4.y, e4.y), fmaf(nv.z, g4.z, e4.z), fmaf(nv.w, g4.w, e4.w));
              const int o = roff[it] + c0;                                 // rows * N < 2^31 (checked by the host)
              if (p.out_f32) *reinterpret_cast<float4 *>(p.out_f32 + o) = y;
              if (p.out_lp) *reinterpret_cast<uint2 *>(p.out_lp + o) = make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
              if (p.out_lp_pos) {
                int pr = pr0 + row;
                while (pr >= p.pos_rows) pr -= p.pos_rows;
                const float4 ps = __ldg(reinterpret_cast<const float4 *>(p.pos + (int64_t)pr * p.N + c0 + c4 * 4));
                *reinterpret_cast<uint2 *>(p.out_lp_pos + o) = make_uint2(pack_bf16x2(y.x + ps.x, y.y + ps.y), pack_bf16x2(y.z + ps.z, y.w + ps.w));
              }
            }
          }
          __syncwarp();
        }
      }
      fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty[acc]);
    }
  }

  fence_before_thread_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int encode_2d(CUtensorMap *map, const void *base, uint64_t inner, uint64_t rows, uint64_t row_stride_elems, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return fail(DVIS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t dims[3] = {inner, rows, 1};
  const cuuint64_t strides[2] = {row_stride_elems * 2, row_stride_elems * rows * 2};
  const cuuint32_t box[3] = {uint32_t(kBlockK), box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DVIS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", int(r));
  return DVIS_OK;
}

template <int MODE>
int launch_linear_tc(const void *x, int64_t ldx, const void *w, LinearTcParams p, int K, cudaStream_t s) {
  p.KB = K / kBlockK;
  p.total_tiles = (p.rows + kTileM - 1) / kTileM;
  const int b_bytes = p.KB * p.N * 128;
  const int fixed = 1024 + b_bytes + kEpiWarps * kXposeBytes + (MODE == kAddLn ? 0 : 256 * 4) + int(sizeof(Barriers));
  p.stages = std::min(kMaxStages, (kSmemLimit - fixed) / kStageBytes);
  if (p.stages < 2) return fail(DVIS_ERR_UNSUPPORTED, "linear_tc: N=%d, K=%d do not fit in shared memory", p.N, K);
  const size_t smem = size_t(fixed) + size_t(p.stages) * kStageBytes;
  CUtensorMap tm_x, tm_w;
  if (int rc = encode_2d(&tm_x, x, K, p.rows, ldx, kTileM)) return rc;
  if (int rc = encode_2d(&tm_w, w, K, p.N, K, p.N)) return rc;
  const int grid = std::min(p.total_tiles, kNumSMs);
  cudaFuncSetAttribute(linear_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  linear_tc_kernel<MODE><<<grid, kThreads, smem, s>>>(tm_x, tm_w, p);
  return check_launch("linear_tc_kernel");
}

int validate(const void *x, int64_t ldx, const void *w, int rows, int N, int K) {
  DVIS_REQUIRE(x && w, "linear_tc: null pointer argument");
  DVIS_REQUIRE(rows > 0 && N > 0 && K > 0, "linear_tc: sizes must be positive");
  DVIS_REQUIRE(N % 32 == 0 && N <= 256, "linear_tc: N must be a multiple of 32 and <= 256 (got %d)", N);
  DVIS_REQUIRE(K % kBlockK == 0 && K <= 512, "linear_tc: K must be a multiple of 64 and <= 512 (got %d)", K);
  DVIS_REQUIRE(ldx >= K && ldx % 8 == 0, "linear_tc: x row stride must be >= K and a multiple of 8 elements");
  DVIS_REQUIRE(aligned16(x) && aligned16(w), "linear_tc: x / w must be 16-byte aligned");
  return DVIS_OK;
}

}  // namespace
}  // namespace dvis

using namespace dvis;

extern "C" int dvis_linear_tc(const void *x, int64_t ldx, const void *w, const float *bias, int relu, int rows, int N, int K,
                              void *y, int64_t ldy, void *stream) {
  if (int rc = validate(x, ldx, w, rows, N, K)) return rc;
  DVIS_REQUIRE(y && aligned16(y) && ldy >= N && ldy % 8 == 0, "linear_tc: y must be 16-byte aligned with a row stride >= N, multiple of 8");
  LinearTcParams p{};
  p.rows = rows; p.N = N; p.bias = bias; p.relu = relu; p.y = static_cast<__nv_bfloat16 *>(y); p.ldy = ldy;
  return launch_linear_tc<kPlain>(x, ldx, w, p, K, static_cast<cudaStream_t>(stream));
}

extern "C" int dvis_linear_tc_heads(const void *x, int64_t ldx, const void *w, const float *bias, int batch, int S, int N, int K,
                                    const uint8_t *row_mask, void *value_hm, void *stream) {
  DVIS_REQUIRE(batch > 0 && S > 0 && (int64_t)batch * S < (int64_t(1) << 31), "linear_tc_heads: extent out of range");
  if (int rc = validate(x, ldx, w, batch * S, N, K)) return rc;
  DVIS_REQUIRE(value_hm && aligned16(value_hm), "linear_tc_heads: value_hm must be 16-byte aligned");
  LinearTcParams p{};
  p.rows = batch * S; p.N = N; p.bias = bias; p.y = static_cast<__nv_bfloat16 *>(value_hm); p.S = S; p.heads = N / 32;
  p.row_mask = row_mask;
  return launch_linear_tc<kHeads>(x, ldx, w, p, K, static_cast<cudaStream_t>(stream));
}

extern "C" int dvis_linear_tc_add_ln(const void *x, int64_t ldx, const void *w, const float *bias, const float *residual,
                                     const float *gamma, const float *beta, float eps, int rows, int N, int K, const float *pos,
                                     int pos_rows, float *out_f32, void *out_lp, void *out_lp_pos, void *stream) {
  if (int rc = validate(x, ldx, w, rows, N, K)) return rc;
  DVIS_REQUIRE(residual && gamma && beta && aligned16(residual), "linear_tc_add_ln: residual / gamma / beta must be given (16-byte aligned)");
  DVIS_REQUIRE(out_f32 || out_lp || out_lp_pos, "linear_tc_add_ln: no output requested");
  DVIS_REQUIRE((int64_t)rows * N < (int64_t(1) << 31), "linear_tc_add_ln: rows * N must be < 2^31");
  DVIS_REQUIRE((!out_f32 || aligned16(out_f32)) && (!out_lp || aligned16(out_lp)) && (!out_lp_pos || aligned16(out_lp_pos)),
               "linear_tc_add_ln: outputs must be 16-byte aligned");
  DVIS_REQUIRE(!out_lp_pos || (pos && pos_rows > 0 && aligned16(pos)), "linear_tc_add_ln: out_lp_pos needs pos");
  LinearTcParams p{};
  p.rows = rows; p.N = N; p.bias = bias; p.residual = residual; p.gamma = gamma; p.beta = beta; p.eps = eps;
  p.out_f32 = out_f32; p.out_lp = static_cast<__nv_bfloat16 *>(out_lp); p.out_lp_pos = static_cast<__nv_bfloat16 *>(out_lp_pos);
  p.pos = pos; p.pos_rows = pos_rows > 0 ? pos_rows : 1;
  return launch_linear_tc<kAddLn>(x, ldx, w, p, K, static_cast<cudaStream_t>(stream));
}
