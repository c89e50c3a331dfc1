// Mask-logit GEMM on the 5th-gen tensor cores:  out[b, q, p] = sum_c emb[b, q, c] * feat[b, p, c]
//
// Replaces torch.einsum("bqc,bchw->bqhw") (P/dvis_Plus/video_mask2former_transformer_decoder.py:363) and the
// tracker / refiner "lbtqc,btchw->lbqthw" (P/dvis_Plus/tracker.py:379, refiner.py:185-189).
//
// Shape of the problem: K = C = 256 is tiny, so the op is HBM-bound (read C*HW, write Q*HW): the design goal is
// to stream `feat` exactly once at full TMA bandwidth and write `out` with fully coalesced stores.
//   * MMA tile: M = 128 pixels (A = feat tile, K-major), N = Q rounded up to 16 (B = emb, K-major, resident in smem
//     for the whole batch element), K = C in 64-element (128-byte, SWIZZLE_128B) blocks.
//   * D lives in TMEM with lane = pixel, column = query; two accumulator stages (2 x 256 columns) so the epilogue
//     of tile i overlaps the MMAs of tile i+1.
//   * Epilogue: tcgen05.ld gives every thread one pixel's logits for 32 queries; they are staged as a
//     [32 queries][128 pixels] smem tile (conflict-free: lane = pixel) and written with one TMA store per chunk,
//     double-buffered, so the (B, Q, HW) output is produced with full-line writes and no per-element predicates.
//   * Persistent CTAs (one per SM), warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc),
//     warps 2..9 = epilogue (two groups of four, one per accumulator stage; a warp per TMEM lane quarter).
#include <algorithm>

#include "common.cuh"
#include "tc05.cuh"

namespace dvis {
namespace {

using namespace tc05;

constexpr int kTileM = 128;
constexpr int kBlockK = 64;                       // bf16 elements per 128-byte swizzle row
constexpr int kStageBytes = kTileM * 128;         // one A k-block: 128 rows x 128 B
constexpr int kMaxStages = 8;
// Epilogue groups of 4 warps, one per TMEM accumulator stage (round 2: a single epilogue warp per scheduler exposes every TMEM /
// shared-memory latency).  Two groups for 2-byte outputs and the bit-mask epilogue (T = 16, Q = 200, bf16: 155.6 -> 148.4 us);
// fp32 outputs keep one group: their 16 KB staging tiles would take the shared memory of the A pipeline (measured 219 -> 230 us).
template <typename TO, bool kBits>
constexpr int epi_groups() { return (kBits || sizeof(TO) == 2) ? 2 : 1; }
constexpr int kThreads = 64 + 2 * 128;            // launch bound; a one-group kernel is launched with 192 threads
constexpr int kAccCols = 256;                     // TMEM columns per accumulator stage
constexpr int kEpiCols = 32;                      // queries per epilogue chunk (one tcgen05.ld.32x32b.x32)
constexpr int kEpiThreads = 128;

struct MaskGemmParams {
  void *out;
  int B, Q, Qpad, KB;   // KB = C / 64
  int64_t HW;
  int tiles_per_batch, total_tiles, stages;
  int *row_open;        // kBias epilogue: row_open[b*Q + q] = 1 if some pixel of the row has logit >= 0
  int64_t out_batch;    // elements between batch items of `out` (Q*HW when dense; larger for a query slice)
  int64_t out_row;      // elements between query rows of `out` (HW when dense; B*HW for the query-major clip layout (Q, B, HW))
  int epi_bufs;         // output staging tiles in flight (2..4): depth of the TMA-store pipeline
  uint8_t *bits;        // kBits epilogue: (B*Q) rows of bits_row bytes, bit (p % 8) of byte (p / 8) set <=> logit[p] < 0
  int64_t bits_row;
  int row_batch;        // rows (queries) per batch item in row_open / bits (= Q unless this launch covers a query slice)
};

struct __align__(8) Barriers {
  uint64_t full[kMaxStages], empty[kMaxStages];
  uint64_t b_full, b_empty;
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

template <typename TO>
__device__ __forceinline__ TO cvt_logit(uint32_t bits);
template <>
__device__ __forceinline__ float cvt_logit<float>(uint32_t bits) { return __uint_as_float(bits); }
template <>
__device__ __forceinline__ __nv_bfloat16 cvt_logit<__nv_bfloat16>(uint32_t bits) {
  return __float2bfloat16_rn(__uint_as_float(bits));
}

// kTmaStore = false: direct (still coalesced) global stores for outputs whose row pitch is not a multiple of 16 bytes
// kBias = true: the epilogue writes the additive attention bias of the masked-attention decoder instead of the logits:
// -inf where sigmoid(logit) < 0.5 <=> logit < 0, else 0 (decoder.py:370-371), and records per (b, q) row whether any pixel
// stays open so that fully masked rows can be reset afterwards (decoder.py:297).
// kBits = true: the epilogue writes ONE BIT per (query, pixel) -- set where the query may NOT attend the pixel -- instead of
// a 16- or 32-bit bias: what csrc/flash_attn.cu consumes (16x fewer bytes than the bf16 bias, L2-resident; nothing else of the
// 9 intermediate mask-head calls ever reaches HBM).  A warp's 32 lanes are 32 consecutive pixels, so one ballot per query
// column is the packed word; lane i keeps column i's word and the 32 words of a chunk go out as 32 4-byte stores.
// kTf32 = true: emb / feat are fp32 in memory and multiplied as TF32 (kind::tf32; a 128-byte swizzle row holds 32 elements): the
// fp32-mode mask head (decoder.py:363 without autocast) at 1e-3 of the output scale instead of bf16's 1e-2.
template <typename TO, bool kTmaStore, bool kBias = false, bool kBits = false, bool kTf32 = false>
__global__ void __launch_bounds__(kThreads, 1)
mask_gemm_kernel(const __grid_constant__ CUtensorMap tmap_feat, const __grid_constant__ CUtensorMap tmap_emb,
                 const __grid_constant__ CUtensorMap tmap_out, const MaskGemmParams p) {
  constexpr int kEpiGroups = epi_groups<TO, kBits>();
  extern __shared__ uint8_t smem_raw[];
  // aligned by offset arithmetic on the __shared__ array (a uintptr_t round trip would turn the staging stores into generic ST)
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int b_block_bytes = p.Qpad * 128;                       // one emb k-block
  uint8_t *sB = smem;
  uint8_t *sA = smem + p.KB * b_block_bytes;                    // 1024-aligned because Qpad % 8 == 0
  // epi_bufs output staging tiles [kEpiCols queries][128 pixels] for the TMA store (row-major, no swizzle)
  TO *sOut = reinterpret_cast<TO *>(sA + p.stages * kStageBytes);      // kEpiGroups x epi_bufs staging tiles
  Barriers *bars = reinterpret_cast<Barriers *>(reinterpret_cast<uint8_t *>(sOut) +
                                                (kBits ? 0 : kEpiGroups * p.epi_bufs * kEpiCols * kTileM * sizeof(TO)));   // bits: no staging

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kBlkElems = kTf32 ? 32 : 64;                   // elements per 128-byte k-block row

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_feat);
    prefetch_tensormap(&tmap_emb);
    prefetch_tensormap(&tmap_out);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    mbar_init(&bars->b_full, 1);
    mbar_init(&bars->b_empty, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&bars->acc_full[a], 1); mbar_init(&bars->acc_empty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem_base = bars->tmem_base;

  // contiguous, balanced slice of the (batch, pixel-tile) list for this CTA
  const int tile_begin = int((int64_t)p.total_tiles * blockIdx.x / gridDim.x);
  const int tile_end = int((int64_t)p.total_tiles * (blockIdx.x + 1) / gridDim.x);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0, phase = 0, cur_b = -1, n_bload = 0;
      for (int t = tile_begin; t < tile_end; ++t) {
        const int b = t / p.tiles_per_batch, tile = t - b * p.tiles_per_batch;
        if (b != cur_b) {
          if (n_bload > 0) mbar_wait(&bars->b_empty, (n_bload - 1) & 1);   // MMAs that read the old emb are done
          mbar_arrive_expect_tx(&bars->b_full, uint32_t(p.KB * b_block_bytes));
          for (int kb = 0; kb < p.KB; ++kb) tma_load_3d(sB + kb * b_block_bytes, &tmap_emb, &bars->b_full, kb * kBlkElems, 0, b);
          cur_b = b;
          ++n_bload;
        }
        for (int kb = 0; kb < p.KB; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&bars->full[stage], kStageBytes);
          tma_load_3d(sA + stage * kStageBytes, &tmap_feat, &bars->full[stage], kb * kBlkElems, tile * kTileM, b);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(kTileM, p.Qpad, kTf32 ? /*TF32*/ 2 : /*BF16*/ 1);
      int stage = 0, phase = 0, cur_b = -1, n_bload = 0, n_tile = 0;
      for (int t = tile_begin; t < tile_end; ++t, ++n_tile) {
        const int b = t / p.tiles_per_batch;
        if (b != cur_b) {
          mbar_wait(&bars->b_full, n_bload & 1);
          cur_b = b;
          ++n_bload;
        }
        const int acc = n_tile & 1;
        mbar_wait(&bars->acc_empty[acc], ((n_tile >> 1) & 1) ^ 1);     // epilogue drained this accumulator
        fence_after_thread_sync();
        const uint32_t d_tmem = tmem_base + acc * kAccCols;
        for (int kb = 0; kb < p.KB; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          fence_after_thread_sync();
          const uint32_t a_addr = smem_u32(sA + stage * kStageBytes);
          const uint32_t b_addr = smem_u32(sB + kb * b_block_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {                                 // UMMA_K = 16 bf16 / 8 tf32 = 32 bytes inside the swizzle row
            if constexpr (kTf32)
              mma_tf32_ss(d_tmem, make_desc_k_sw128(a_addr + k * 32), make_desc_k_sw128(b_addr + k * 32), idesc, uint32_t(kb | k));
            else
              mma_bf16_ss(d_tmem, make_desc_k_sw128(a_addr + k * 32), make_desc_k_sw128(b_addr + k * 32), idesc, uint32_t(kb | k));
          }
          mma_commit(&bars->empty[stage]);                               // frees the A stage when these MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        mma_commit(&bars->acc_full[acc]);
        const bool last_of_batch = (t + 1 < tile_end) && ((t + 1) / p.tiles_per_batch != b);
        if (last_of_batch) mma_commit(&bars->b_empty);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    // TMEM -> registers -> smem tile [query][pixel] -> TMA store.  Thread = pixel, so for a fixed query the warp's
    // 32 lanes write 32 consecutive pixels of the staging row (conflict-free) and the TMA engine does the
    // (Q, HW)-strided global writes, clipping partial tiles in both pixel and query direction.
    const int quarter = warp & 3;                     // TMEM lanes [32*quarter, 32*quarter+32) are this warp's
    const int group = (warp - 2) >> 2;                // group g drains accumulator g: the tiles with n_tile % 2 == g
    const bool issuer = (((warp - 2) & 3) == 0 && lane == 0);
    const int px = quarter * 32 + lane;
    TO *sOutG = sOut + group * p.epi_bufs * (kEpiCols * kTileM);
    int n_chunk = 0;
    for (int n_tile = group, t = tile_begin + group; t < tile_end; t += kEpiGroups, n_tile += kEpiGroups) {
      const int b = t / p.tiles_per_batch, tile = t - b * p.tiles_per_batch;
      const int acc = n_tile & 1;
      mbar_wait(&bars->acc_full[acc], (n_tile >> 1) & 1);
      fence_after_thread_sync();
      const uint32_t taddr = tmem_base + acc * kAccCols + (uint32_t(quarter * 32) << 16);
      for (int c0 = 0; c0 < p.Q; c0 += kEpiCols, ++n_chunk) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c0, r);
        tmem_ld_wait();
        if constexpr (kBits) {
          const bool pix_ok = (int64_t)tile * kTileM + px < p.HW;
          uint32_t mine = 0;
#pragma unroll
          for (int i = 0; i < kEpiCols; ++i) {
            const bool open = pix_ok && !(__uint_as_float(r[i]) < 0.f);
            const unsigned any = __ballot_sync(0xffffffffu, open);
            if (any && lane == 0 && c0 + i < p.Q) p.row_open[b * p.row_batch + c0 + i] = 1;
            if (lane == i) mine = ~any;
          }
          const int64_t off = (int64_t)tile * (kTileM / 8) + quarter * 4;
          if (c0 + lane < p.Q && off + 4 <= p.bits_row)
            *reinterpret_cast<uint32_t *>(p.bits + ((int64_t)b * p.row_batch + c0 + lane) * p.bits_row + off) = mine;
          continue;
        }
        if constexpr (kBias) {
          const bool pix_ok = (int64_t)tile * kTileM + px < p.HW;
#pragma unroll
          for (int i = 0; i < kEpiCols; ++i) {
            const bool open = pix_ok && !(__uint_as_float(r[i]) < 0.f);
            const unsigned any = __ballot_sync(0xffffffffu, open);
            if (any && lane == 0 && c0 + i < p.Q) p.row_open[b * p.row_batch + c0 + i] = 1;
            r[i] = open ? 0u : 0xff800000u;                 // 0.0f / -inf
          }
        }
        if constexpr (kTmaStore) {
          TO *buf = sOutG + (n_chunk % p.epi_bufs) * (kEpiCols * kTileM);
          if (issuer) {                                  // the store that last read this buffer (epi_bufs chunks ago) is done
            if (p.epi_bufs == 4) tma_store_wait_read<3>();
            else if (p.epi_bufs == 3) tma_store_wait_read<2>();
            else if (p.epi_bufs == 2) tma_store_wait_read<1>();
            else tma_store_wait_read<0>();
          }
          named_barrier_sync(1 + 2 * group, kEpiThreads);
#pragma unroll
          for (int i = 0; i < kEpiCols; ++i) buf[i * kTileM + px] = cvt_logit<TO>(r[i]);
          fence_proxy_async();                           // generic-proxy smem writes -> visible to the TMA engine
          named_barrier_sync(2 + 2 * group, kEpiThreads);
          if (issuer) {
            tma_store_3d(&tmap_out, buf, tile * kTileM, c0, b);
            tma_store_commit();
          }
        } else {
          const int64_t pix = (int64_t)tile * kTileM + px;
          if (pix < p.HW) {
            TO *dst = static_cast<TO *>(p.out) + (int64_t)b * p.out_batch + (int64_t)c0 * p.out_row + pix;
            const int nq = min(kEpiCols, p.Q - c0);
            for (int i = 0; i < nq; ++i) dst[(int64_t)i * p.out_row] = cvt_logit<TO>(r[i]);
          }
        }
      }
      fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty[acc]);
    }
    if (kTmaStore && issuer) tma_store_wait_read<0>();
  }

  fence_before_thread_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// 3-D map over a dense (batch, rows, inner) tensor; box = (box_inner, box_rows, 1)
int encode_map(CUtensorMap *map, const void *base, CUtensorMapDataType dt, int esize, uint64_t inner, uint64_t rows,
               uint64_t batch, uint32_t box_inner, uint32_t box_rows, CUtensorMapSwizzle swz, uint64_t batch_stride = 0,
               uint64_t row_stride = 0) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return fail(DVIS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t dims[3] = {inner, rows, batch};
  const cuuint64_t strides[2] = {(row_stride ? row_stride : inner) * esize, (batch_stride ? batch_stride : inner * rows) * esize};   // bytes, dims 1..2
  const cuuint32_t box[3] = {box_inner, box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, dt, 3, const_cast<void *>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DVIS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", int(r));
  return DVIS_OK;
}

}  // namespace
}  // namespace dvis

using namespace dvis;

namespace dvis {
namespace {
// rows whose every position is masked attend everywhere: reset them to 0 (decoder.py:297)
template <typename TB>
__global__ void __launch_bounds__(256) reset_closed_rows_kernel(const int *__restrict__ row_open, TB *__restrict__ bias, int64_t HW) {
  const int64_t row = blockIdx.x;
  if (row_open[row]) return;
  TB *b = bias + row * HW;
  for (int64_t i = threadIdx.x; i < HW; i += blockDim.x) b[i] = TB(0.f);
}
__global__ void __launch_bounds__(128) reset_closed_bit_rows_kernel(const int *__restrict__ row_open, uint8_t *__restrict__ bits,
                                                                   int64_t row_bytes) {
  const int64_t row = blockIdx.x;
  if (row_open[row]) return;
  uint32_t *w = reinterpret_cast<uint32_t *>(bits + row * row_bytes);
  for (int64_t i = threadIdx.x; i < row_bytes / 4; i += blockDim.x) w[i] = 0u;
}
}  // namespace
int mask_gemm_launch(const void *emb, const void *feat, int B, int Q, int C, int64_t HW, void *out, int out_dtype,
                     int *row_open, void *stream, int64_t emb_batch = 0, int64_t out_batch = 0, int64_t bits_row = 0,
                     int row_batch = 0, bool tf32 = false, int64_t out_row = 0);
}  // namespace dvis

extern "C" int dvis_mask_logits(const void *emb, const void *feat, int B, int Q, int C, int64_t HW, void *out,
                                int out_dtype, void *stream) {
  return mask_gemm_launch(emb, feat, B, Q, C, HW, out, out_dtype, nullptr, stream);
}

extern "C" int dvis_mask_logits_strided(const void *emb, int64_t emb_batch_stride, const void *feat, int B, int Q, int C,
                                        int64_t HW, void *out, int64_t out_batch_stride, int out_dtype, void *stream) {
  DVIS_REQUIRE(emb_batch_stride >= (int64_t)Q * C && out_batch_stride >= (int64_t)Q * HW, "mask_logits_strided: batch strides too small");
  DVIS_REQUIRE(emb_batch_stride % 8 == 0, "mask_logits_strided: emb batch stride must be a multiple of 8 elements (16 bytes)");
  return mask_gemm_launch(emb, feat, B, Q, C, HW, out, out_dtype, nullptr, stream, emb_batch_stride, out_batch_stride);
}

// The mask logits of a CLIP in the layout the meta-architecture keeps them in, (Q, T, H, W) = "b q t h w" with b = 1
// (P/dvis_Plus/refiner.py:185-189 "lbtqc,btchw->lbqthw"): frame t is a batch item of the GEMM whose rows are written T*HW apart --
// no transposition pass over the (T, Q, HW) result (377 MB at T = 16, Q = 200, 720p, bf16).
extern "C" int dvis_mask_logits_clip(const void *emb, const void *feat, int T, int Q, int C, int64_t HW, void *out, int out_dtype,
                                     void *stream) {
  DVIS_REQUIRE(Q <= 256, "mask_logits_clip: Q must be <= 256 (got %d)", Q);
  return mask_gemm_launch(emb, feat, T, Q, C, HW, out, out_dtype, nullptr, stream, 0, HW, 0, 0, false, (int64_t)T * HW);
}

extern "C" int dvis_mask_attn_bias(const void *emb, const void *feat, int B, int Q, int C, int64_t HW, void *bias,
                                   int bias_dtype, int *row_open_workspace, void *stream) {
  DVIS_REQUIRE(row_open_workspace, "mask_attn_bias: null workspace");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaMemsetAsync(row_open_workspace, 0, sizeof(int) * size_t(B) * Q, s);
  if (int rc = mask_gemm_launch(emb, feat, B, Q, C, HW, bias, bias_dtype, row_open_workspace, stream)) return rc;
  if (bias_dtype == DVIS_F32)
    reset_closed_rows_kernel<float><<<B * Q, 256, 0, s>>>(row_open_workspace, static_cast<float *>(bias), HW);
  else
    reset_closed_rows_kernel<__nv_bfloat16><<<B * Q, 256, 0, s>>>(row_open_workspace, static_cast<__nv_bfloat16 *>(bias), HW);
  return check_launch("reset_closed_rows_kernel");
}

extern "C" int dvis_mask_attn_bits(const void *emb, const void *feat, int B, int Q, int C, int64_t HW, void *bits,
                                   int64_t bits_row_bytes, int *row_open_workspace, void *stream) {
  DVIS_REQUIRE(row_open_workspace && bits, "mask_attn_bits: null pointer argument");
  DVIS_REQUIRE(bits_row_bytes % 8 == 0 && bits_row_bytes >= ((HW + 63) / 64) * 8 && (reinterpret_cast<uintptr_t>(bits) & 7) == 0,
               "mask_attn_bits: rows must be 8-byte aligned and hold ceil(HW / 64) * 8 bytes");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaMemsetAsync(row_open_workspace, 0, sizeof(int) * size_t(B) * Q, s);
  for (int q0 = 0; q0 < Q; q0 += 256) {             // the GEMM holds at most 256 queries in TMEM: slices of the query dim
    const int nq = std::min(256, Q - q0);
    if (int rc = mask_gemm_launch(static_cast<const __nv_bfloat16 *>(emb) + (size_t)q0 * C, feat, B, nq, C, HW,
                                  static_cast<uint8_t *>(bits) + (size_t)q0 * bits_row_bytes, DVIS_F32, row_open_workspace + q0,
                                  stream, (int64_t)Q * C, 0, bits_row_bytes, Q))
      return rc;
  }
  reset_closed_bit_rows_kernel<<<B * Q, 128, 0, s>>>(row_open_workspace, static_cast<uint8_t *>(bits), bits_row_bytes);
  return check_launch("reset_closed_bit_rows_kernel");
}

// fp32 operands multiplied as TF32.  emb (Q x C fp32 = 1 KB per query) stays resident in shared memory, so the queries are
// processed in slices of <= 128 (feat is re-read once per slice: 2 passes at Q = 200).
constexpr int kTf32QuerySlice = 128;
extern "C" int dvis_mask_logits_tf32(const void *emb, const void *feat, int B, int Q, int C, int64_t HW, void *out, void *stream) {
  DVIS_REQUIRE(emb && feat && out && Q > 0, "mask_logits_tf32: null pointer argument");
  for (int q0 = 0; q0 < Q; q0 += kTf32QuerySlice) {
    const int nq = std::min(kTf32QuerySlice, Q - q0);
    if (int rc = mask_gemm_launch(static_cast<const float *>(emb) + (size_t)q0 * C, feat, B, nq, C, HW,
                                  static_cast<float *>(out) + (size_t)q0 * HW, DVIS_F32, nullptr, stream, (int64_t)Q * C,
                                  (int64_t)Q * HW, 0, 0, true))
      return rc;
  }
  return DVIS_OK;
}

extern "C" int dvis_mask_attn_bias_tf32(const void *emb, const void *feat, int B, int Q, int C, int64_t HW, void *bias,
                                        int *row_open_workspace, void *stream) {
  DVIS_REQUIRE(emb && feat && bias && row_open_workspace && Q > 0, "mask_attn_bias_tf32: null pointer argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaMemsetAsync(row_open_workspace, 0, sizeof(int) * size_t(B) * Q, s);
  for (int q0 = 0; q0 < Q; q0 += kTf32QuerySlice) {
    const int nq = std::min(kTf32QuerySlice, Q - q0);
    if (int rc = mask_gemm_launch(static_cast<const float *>(emb) + (size_t)q0 * C, feat, B, nq, C, HW,
                                  static_cast<float *>(bias) + (size_t)q0 * HW, DVIS_F32, row_open_workspace + q0, stream,
                                  (int64_t)Q * C, (int64_t)Q * HW, 0, Q, true))
      return rc;
  }
  reset_closed_rows_kernel<float><<<B * Q, 256, 0, s>>>(row_open_workspace, static_cast<float *>(bias), HW);
  return check_launch("reset_closed_rows_kernel");
}

int dvis::mask_gemm_launch(const void *emb, const void *feat, int B, int Q, int C, int64_t HW, void *out, int out_dtype,
                           int *row_open, void *stream, int64_t emb_batch, int64_t out_batch, int64_t bits_row, int row_batch,
                           bool tf32, int64_t out_row) {
  DVIS_REQUIRE(emb && feat && out, "mask_logits: null pointer argument");
  DVIS_REQUIRE(B > 0 && Q > 0 && C > 0 && HW > 0, "mask_logits: sizes must be positive");
  const int blk = tf32 ? 32 : kBlockK;                         // operand elements per 128-byte k-block row
  DVIS_REQUIRE(C % kBlockK == 0 && C <= 512, "mask_logits: C must be a multiple of 64 and <= 512 (got %d)", C);
  DVIS_REQUIRE(!tf32 || (out_dtype == DVIS_F32 && !bits_row), "mask_logits: tf32 operands are built for f32 outputs");
  DVIS_REQUIRE(Q <= 256, "mask_logits: Q must be <= 256 (got %d); split the queries", Q);
  DVIS_REQUIRE(aligned16(emb) && aligned16(feat), "mask_logits: emb / feat must be 16-byte aligned");
  DVIS_REQUIRE(HW < (int64_t(1) << 31) && B < 65536, "mask_logits: extent too large");
  if (out_dtype != DVIS_F32 && out_dtype != DVIS_BF16)
    return fail(DVIS_ERR_UNSUPPORTED, "mask_logits: out_dtype %d (f32 or bf16)", out_dtype);

  MaskGemmParams p{};
  p.out = out; p.B = B; p.Q = Q; p.Qpad = (Q + 15) & ~15; p.KB = C / blk; p.HW = HW;
  p.row_open = row_open;
  p.bits = bits_row ? static_cast<uint8_t *>(out) : nullptr;
  p.bits_row = bits_row;
  p.row_batch = row_batch ? row_batch : Q;
  p.out_batch = out_batch ? out_batch : (int64_t)Q * HW;
  p.out_row = out_row ? out_row : HW;
  p.tiles_per_batch = int((HW + kTileM - 1) / kTileM);
  p.total_tiles = p.tiles_per_batch * B;
  const int b_bytes = p.KB * p.Qpad * 128;
  const int esize = out_dtype == DVIS_F32 ? 4 : 2;
  const bool tma_store = !bits_row && aligned16(out) && (p.out_row * esize) % 16 == 0 && (p.out_batch * esize) % 16 == 0;   // TMA: 16-byte pitches
  // deepest store pipeline (up to 4 staging tiles per epilogue group) that still leaves >= 4 A stages
  const int groups = (bits_row || esize == 2) ? 2 : 1;          // = epi_groups<TO, kBits>() of the kernel launched below
  const int threads = 64 + 128 * groups;
  int stage_out_bytes = 0, budget = 0;
  for (p.epi_bufs = 4; p.epi_bufs >= 1; --p.epi_bufs) {
    stage_out_bytes = bits_row ? 0 : groups * p.epi_bufs * kEpiCols * kTileM * esize;
    budget = 225 * 1024 - b_bytes - 1024 - stage_out_bytes - int(sizeof(Barriers));
    if (budget / kStageBytes >= 4 || p.epi_bufs == 1) break;
  }
  p.stages = std::min(kMaxStages, budget / kStageBytes);
  if (p.stages < 2) return fail(DVIS_ERR_UNSUPPORTED, "mask_logits: Q=%d, C=%d do not fit in shared memory", Q, C);
  const size_t smem = 1024 + b_bytes + size_t(p.stages) * kStageBytes + stage_out_bytes + sizeof(Barriers);

  CUtensorMap tm_feat, tm_emb, tm_out;
  const CUtensorMapDataType in_dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const int in_es = tf32 ? 4 : 2;
  if (int rc = encode_map(&tm_feat, feat, in_dt, in_es, C, HW, B, blk, kTileM, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  if (int rc = encode_map(&tm_emb, emb, in_dt, in_es, C, Q, B, blk, p.Qpad, CU_TENSOR_MAP_SWIZZLE_128B, emb_batch)) return rc;
  if (tma_store) {
    if (int rc = encode_map(&tm_out, out, out_dtype == DVIS_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                            esize, HW, Q, B, kTileM, kEpiCols, CU_TENSOR_MAP_SWIZZLE_NONE, p.out_batch, p.out_row)) return rc;
  } else {
    tm_out = tm_emb;   // unused by the direct-store variant; any valid map
  }

  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = std::min(p.total_tiles, kNumSMs);
#define DVIS_LAUNCH(TO, TMA, BIAS)                                                                                   \
  do {                                                                                                               \
    cudaFuncSetAttribute(mask_gemm_kernel<TO, TMA, BIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));    \
    mask_gemm_kernel<TO, TMA, BIAS><<<grid, threads, smem, s>>>(tm_feat, tm_emb, tm_out, p);                         \
  } while (0)
#define DVIS_LAUNCH_T(TO)                                                                            \
  do {                                                                                               \
    if (row_open) { if (tma_store) DVIS_LAUNCH(TO, true, true); else DVIS_LAUNCH(TO, false, true); } \
    else { if (tma_store) DVIS_LAUNCH(TO, true, false); else DVIS_LAUNCH(TO, false, false); }        \
  } while (0)
#define DVIS_LAUNCH_TF32(TMA, BIAS)                                                                                              \
  do {                                                                                                                         \
    cudaFuncSetAttribute(mask_gemm_kernel<float, TMA, BIAS, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)); \
    mask_gemm_kernel<float, TMA, BIAS, false, true><<<grid, threads, smem, s>>>(tm_feat, tm_emb, tm_out, p);                   \
  } while (0)
  if (tf32) {
    if (row_open) { if (tma_store) DVIS_LAUNCH_TF32(true, true); else DVIS_LAUNCH_TF32(false, true); }
    else { if (tma_store) DVIS_LAUNCH_TF32(true, false); else DVIS_LAUNCH_TF32(false, false); }
  } else if (bits_row) {
    cudaFuncSetAttribute(mask_gemm_kernel<float, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    mask_gemm_kernel<float, false, false, true><<<grid, threads, smem, s>>>(tm_feat, tm_emb, tm_out, p);
  } else if (out_dtype == DVIS_F32) DVIS_LAUNCH_T(float); else DVIS_LAUNCH_T(__nv_bfloat16);
#undef DVIS_LAUNCH_TF32
#undef DVIS_LAUNCH_T
#undef DVIS_LAUNCH
  return check_launch("mask_gemm_kernel");
}
