// Warp-level tensor-core primitives for the SMALL-M kernels of the temporal stage (csrc/flash_attn.cu, csrc/small_linear.cu):
// mma.sync.m16n8k16 (bf16 x bf16 -> f32), ldmatrix, cp.async.  200-row problems are latency-bound: a tcgen05 pipeline (TMEM
// allocation, 128-row tiles, an MMA-issuer warp) would cost more than it saves there, so these kernels keep the accumulators
// in registers.  The large GEMMs (csrc/mask_gemm.cu, csrc/linear_tc05.cu) use tcgen05 / TMEM / TMA instead.
//
// Fragment layouts (PTX ISA, mma.m16n8k16 with .bf16), g = lane / 4, t = lane % 4:
//   A (16x16, row):  a0 = A[g][2t..2t+1]   a1 = A[g+8][2t..2t+1]   a2 = A[g][2t+8..2t+9]   a3 = A[g+8][2t+8..2t+9]
//   B (16x8,  col):  b0 = B[2t..2t+1][g]   b1 = B[2t+8..2t+9][g]
//   C (16x8):        c0 = C[g][2t]  c1 = C[g][2t+1]  c2 = C[g+8][2t]  c3 = C[g+8][2t+1]
// Under DVIS_SIMT_EMULATION (tests/simt, CPU test infrastructure only) the same functions are implemented on the emulator's
// warp exchange so the kernels' fragment bookkeeping is checked on the host.
#pragma once
#include "common.cuh"

namespace dvis {

#ifndef DVIS_SIMT_EMULATION

__device__ __forceinline__ uint32_t smem_addr_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// four 8x8 b16 matrices; lane l supplies the address of row (l % 8) of matrix (l / 8); r[j] = this lane's pair of matrix j
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void *row_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_addr_u32(row_ptr)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void *row_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_addr_u32(row_ptr)));
}

// Shared-memory addresses as 32-bit values, computed once and advanced by byte offsets: with one warp per scheduler these
// kernels are bound by INSTRUCTION COUNT (profiles/r2_temporal_kernels.md), so the hot loops must not re-derive
// generic -> shared conversions or 64-bit addresses.
typedef uint32_t saddr_t;
__device__ __forceinline__ saddr_t saddr(const void *p) { return smem_addr_u32(p); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], saddr_t a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], saddr_t a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void cp_async_16(saddr_t dst, const void *gmem_src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_16(saddr_t dst, const void *gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
}

// 16-byte asynchronous global -> shared copy (L2 only: the operands are read once per CTA); src_bytes < 16 zero-fills
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src, int src_bytes = 16) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr_u32(smem_dst)), "l"(gmem_src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// programmatic dependent launch (griddepcontrol): a kernel launched with the programmatic-stream-serialization attribute
// may start while its predecessor drains; everything before pdl_wait() must not read the predecessor's outputs.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// launch with (or, when dvis_set_pdl(0), without) the programmatic-stream-serialization attribute
extern bool g_pdl;
template <typename P>
inline void launch_pdl(void (*kern)(P), dim3 grid, dim3 block, size_t smem, cudaStream_t s, const P &params) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, params);
}

#else  // ---------------- CPU emulation (tests/simt/simt_shim.h) ----------------

inline void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  const uint32_t mine[6] = {a[0], a[1], a[2], a[3], b0, b1};
  uint32_t all[32][6];
  simt::warp_gather(mine, all);
  const int lane = simt::t_linear & 31, g = lane >> 2, t = lane & 3;
  auto bf = [](uint32_t w, int half) { return __uint_as_float(half ? (w & 0xffff0000u) : (w << 16)); };
  auto A = [&](int r, int k) {   // element A[r][k] from the lane that holds it
    const int src = (r & 7) * 4 + ((k & 7) >> 1), reg = (r >> 3) + 2 * (k >> 3);
    return bf(all[src][reg], k & 1);
  };
  auto B = [&](int k, int n) {
    const int src = n * 4 + ((k & 7) >> 1), reg = 4 + (k >> 3);
    return bf(all[src][reg], k & 1);
  };
  for (int i = 0; i < 4; ++i) {
    const int r = g + 8 * (i >> 1), n = 2 * t + (i & 1);
    float acc = c[i];
    for (int k = 0; k < 16; ++k) acc += A(r, k) * B(k, n);
    c[i] = acc;
  }
}

inline void ldmatrix_generic(uint32_t (&r)[4], const void *row_ptr, bool trans) {
  const uint64_t mine[1] = {reinterpret_cast<uint64_t>(row_ptr)};
  uint64_t all[32][1];
  simt::warp_gather(mine, all);
  const int lane = simt::t_linear & 31;
  for (int j = 0; j < 4; ++j) {
    auto elem = [&](int row, int col) { return reinterpret_cast<const uint16_t *>(all[8 * j + row][0])[col]; };
    uint16_t lo, hi;
    if (!trans) { lo = elem(lane >> 2, 2 * (lane & 3)); hi = elem(lane >> 2, 2 * (lane & 3) + 1); }
    else { lo = elem(2 * (lane & 3), lane >> 2); hi = elem(2 * (lane & 3) + 1, lane >> 2); }
    r[j] = uint32_t(lo) | (uint32_t(hi) << 16);
  }
  simt::warp_barrier().arrive_and_wait();   // warp-synchronous like the instruction: nobody overwrites the rows early
}
inline void ldmatrix_x4(uint32_t (&r)[4], const void *row_ptr) { ldmatrix_generic(r, row_ptr, false); }
inline void ldmatrix_x4_trans(uint32_t (&r)[4], const void *row_ptr) { ldmatrix_generic(r, row_ptr, true); }
// "shared addresses" are plain host pointers with byte arithmetic
struct saddr_t {
  char *p;
  saddr_t operator+(long o) const { return saddr_t{p + o}; }
  saddr_t &operator+=(long o) { p += o; return *this; }
};
inline saddr_t saddr(const void *p) { return saddr_t{static_cast<char *>(const_cast<void *>(p))}; }
inline void ldmatrix_x4(uint32_t (&r)[4], saddr_t a) { ldmatrix_generic(r, a.p, false); }
inline void ldmatrix_x4_trans(uint32_t (&r)[4], saddr_t a) { ldmatrix_generic(r, a.p, true); }

inline void cp_async_16(void *smem_dst, const void *gmem_src, int src_bytes = 16) {
  std::memset(smem_dst, 0, 16);
  std::memcpy(smem_dst, gmem_src, size_t(src_bytes));
}
inline void cp_async_16(saddr_t dst, const void *gmem_src, int src_bytes = 16) { cp_async_16(static_cast<void *>(dst.p), gmem_src, src_bytes); }
inline void cp_async_commit() {}
template <int N>
inline void cp_async_wait() {}
inline void pdl_wait() {}
inline void pdl_launch_dependents() {}

#endif

}  // namespace dvis
