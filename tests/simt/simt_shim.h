// TEST INFRASTRUCTURE ONLY -- a minimal SIMT emulator so that the CPU test-suite can execute the CUDA-core kernels of
// libdvis_b200 (csrc/postproc.cu, csrc/lap.cu: no tensor cores, no TMA) from their ORIGINAL sources without a GPU.
//
// tests/simt/simt_build.py turns a .cu file into a g++ translation unit (kernel launches `k<<<g, b, s, st>>>(args)` become
// SIMT_LAUNCH(...), `extern __shared__` arrays become pointers into a per-block buffer) and includes this header instead
// of <cuda_runtime.h> / <cuda_bf16.h>.  Every CUDA thread of a block runs as an OS thread; blocks run one after the other.
// __syncthreads() and the warp collectives (__shfl_xor_sync, __ballot_sync, __any_sync) are rendezvous points
// (std::barrier; threads that return early drop out, as on the device).  What this checks: launch geometry, index
// arithmetic, guards, shared-memory protocols and the host-side argument handling of the C-ABI entry points -- everything
// except timing and the hardware's own floating-point contraction.  It is never linked into, or loaded by, the product.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <barrier>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct float2 { float x, y; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct SimtIdx { unsigned x, y, z; };
inline thread_local SimtIdx threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

// ---- bf16 ------------------------------------------------------------------------------------------------------------
struct __nv_bfloat16 {
  uint16_t bits;
  __nv_bfloat16() = default;
  explicit __nv_bfloat16(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) { bits = uint16_t((u >> 16) | 0x40); return; }   // NaN
    u += 0x7fffu + ((u >> 16) & 1u);                                                      // round to nearest even
    bits = uint16_t(u >> 16);
  }
  operator float() const {
    uint32_t u = uint32_t(bits) << 16;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
  }
};
struct __nv_bfloat162 { __nv_bfloat16 x, y; };
inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) { return __nv_bfloat162{__nv_bfloat16(a), __nv_bfloat16(b)}; }
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float __bfloat162float(__nv_bfloat16 b) { return float(b); }
inline float __expf(float x) { return std::exp(x); }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
// packed f32x2 fused multiply-add (fma.rn.f32x2): two independent fmaf
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)}; }
// fp16 pairs (csrc/msda_pair.cu)
struct __half2 { _Float16 x, y; };
inline __half2 __floats2half2_rn(float a, float b) { return __half2{(_Float16)a, (_Float16)b}; }
inline float __low2float(__half2 h) { return float(h.x); }
inline float __high2float(__half2 h) { return float(h.y); }
#define __align__(n) alignas(n)
inline long long clock64() { return 0; }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
template <typename T> inline T __ldg(const T *p) { return *p; }
template <typename T> inline T __ldcg(const T *p) { return *p; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }

// ---- runtime API surface the entry points touch --------------------------------------------------------------------------
typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0, cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9,
       cudaSharedmemCarveoutMaxShared = 100 };
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t) { return "simt"; }
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

// ---- execution model ---------------------------------------------------------------------------------------------------------
namespace simt {

struct Block {
  int threads;
  std::unique_ptr<std::barrier<>> bar;                   // __syncthreads
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar; // warp collectives
  std::vector<uint64_t> slot;                            // one 8-byte exchange slot per thread
  std::vector<uint64_t> gather;                          // 64 bytes per thread: warp_gather (mma / ldmatrix emulation)
  std::vector<char> smem;                                // dynamic shared memory
};
inline Block *g_block = nullptr;
inline thread_local int t_linear = 0;                    // linear thread id within the block

// Race shaker: with jitter on, a pseudo-random subset of the threads is delayed right after every block barrier, so code
// that relies on "the other warps will have read this by now" (no barrier between a read and a later write) misbehaves
// reproducibly instead of by scheduling luck.  This is how the p[j1] read-after-barrier race in csrc/lap.cu was found.
inline int g_jitter = 0;
inline thread_local uint32_t t_rng = 0;
inline void jitter_point() {
  if (!g_jitter) return;
  t_rng = t_rng * 1664525u + 1013904223u + uint32_t(t_linear) * 2654435761u;
  if (((t_rng >> 16) % uint32_t(g_jitter)) == 0) std::this_thread::sleep_for(std::chrono::microseconds(200));
}

inline void *dyn_smem() { return g_block->smem.data(); }
inline void syncthreads() { g_block->bar->arrive_and_wait(); jitter_point(); }
inline std::barrier<> &warp_barrier() { return *g_block->warp_bar[t_linear >> 5]; }
inline int warp_width() { return std::min(32, g_block->threads - (t_linear & ~31)); }

template <typename T>
inline T exchange(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "exchange slot");
  uint64_t raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  g_block->slot[t_linear] = raw;
  warp_barrier().arrive_and_wait();
  const int base = t_linear & ~31;
  uint64_t got = (src_lane >= 0 && src_lane < warp_width()) ? g_block->slot[base + src_lane] : raw;
  warp_barrier().arrive_and_wait();
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}

// every lane of a (full) warp publishes N values and receives all 32 lanes' values: the building block of the
// mma.sync / ldmatrix emulation in csrc/mma.cuh
template <typename T, int N>
inline void warp_gather(const T (&mine)[N], T (&all)[32][N]) {
  static_assert(sizeof(T) * N <= 64, "gather slot");
  std::memcpy(&g_block->gather[size_t(t_linear) * 8], mine, sizeof(T) * N);
  warp_barrier().arrive_and_wait();
  const int base = t_linear & ~31;
  for (int l = 0; l < 32; ++l) std::memcpy(all[l], &g_block->gather[size_t(base + l) * 8], sizeof(T) * N);
  warp_barrier().arrive_and_wait();
}

template <typename F>
inline void launch(dim3 grid, dim3 block, size_t smem, F &&body) {
  const int nthr = int(block.x * block.y * block.z);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        Block b;
        b.threads = nthr;
        b.bar = std::make_unique<std::barrier<>>(nthr);
        for (int w = 0; w < (nthr + 31) / 32; ++w) b.warp_bar.push_back(std::make_unique<std::barrier<>>(std::min(32, nthr - 32 * w)));
        b.slot.assign(nthr, 0);
        b.gather.assign(size_t(nthr) * 8, 0);
        b.smem.assign(smem + 16, 0);
        g_block = &b;
        std::vector<std::thread> pool;
        pool.reserve(nthr);
        for (int t = 0; t < nthr; ++t)
          pool.emplace_back([&, t] {
            t_linear = t;
            threadIdx = SimtIdx{unsigned(t) % block.x, (unsigned(t) / block.x) % block.y, unsigned(t) / (block.x * block.y)};
            blockIdx = SimtIdx{bx, by, bz};
            blockDim = block;
            gridDim = grid;
            body();
            b.bar->arrive_and_drop();                      // a thread that has returned no longer takes part
            b.warp_bar[t >> 5]->arrive_and_drop();
          });
        for (auto &th : pool) th.join();
        g_block = nullptr;
      }
}

}  // namespace simt

inline void __syncthreads() { simt::syncthreads(); }
inline void __syncwarp(unsigned = 0xffffffffu) { simt::warp_barrier().arrive_and_wait(); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int lane_mask) { return simt::exchange(v, (simt::t_linear & 31) ^ lane_mask); }
template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return simt::exchange(v, src); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned delta) { return simt::exchange(v, (simt::t_linear & 31) + int(delta)); }
inline unsigned __ballot_sync(unsigned, bool pred) {
  simt::g_block->slot[simt::t_linear] = pred ? 1 : 0;
  simt::warp_barrier().arrive_and_wait();
  const int base = simt::t_linear & ~31;
  unsigned m = 0;
  for (int l = 0; l < simt::warp_width(); ++l) m |= unsigned(simt::g_block->slot[base + l] & 1) << l;
  simt::warp_barrier().arrive_and_wait();
  return m;
}
inline bool __any_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) != 0; }
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
namespace simt {
template <typename F, typename U>
inline F atomic_add_fp(F *p, F v) {          // compare-and-swap loop on the value's bit pattern
  static_assert(sizeof(F) == sizeof(U), "bit pattern type");
  U *up = reinterpret_cast<U *>(p);
  U old = __atomic_load_n(up, __ATOMIC_RELAXED), desired;
  F cur;
  do {
    std::memcpy(&cur, &old, sizeof(F));
    const F sum = cur + v;
    std::memcpy(&desired, &sum, sizeof(F));
  } while (!__atomic_compare_exchange_n(up, &old, desired, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  return cur;
}
}  // namespace simt
inline float atomicAdd(float *p, float v) { return simt::atomic_add_fp<float, uint32_t>(p, v); }
inline double atomicAdd(double *p, double v) { return simt::atomic_add_fp<double, uint64_t>(p, v); }
inline float4 atomicAdd(float4 *p, float4 v) {   // 16-byte vector atomic (red.global.add.v4.f32): four independent adds
  return float4{atomicAdd(&p->x, v.x), atomicAdd(&p->y, v.y), atomicAdd(&p->z, v.z), atomicAdd(&p->w, v.w)};
}

#define SIMT_LAUNCH(kernel, grid, block, smem, stream, ...) \
  simt::launch(dim3(grid), dim3(block), size_t(smem), [&] { kernel(__VA_ARGS__); })
