// Shared host/device helpers for libdvis_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdlib>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "dvis_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libdvis_b200 targets sm_100a (B200) only"
#endif

namespace dvis {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; used for grid sizing heuristics only

// ---- error reporting ---------------------------------------------------------------------------
char *last_error_buffer();  // thread-local, defined in api.cu

inline int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

inline int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(DVIS_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return DVIS_OK;
}

#define DVIS_REQUIRE(cond, ...) \
  do {                          \
    if (!(cond)) return ::dvis::fail(DVIS_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// Experiment switch DVIS_SMEM_CARVEOUT = 0..100 (unset: the driver's choice): preferred shared-memory carve-out applied to a kernel
// before its launch.  Motivation (tests/perf/interference_probe.py): next to the per-frame stage's big kernels, a tiny kernel of
// the temporal stage is nearly free when it needs no shared memory (0.01 - 0.2 us of the big stream lost per tiny kernel) and costs
// 0.7 - 2 us when it needs a large shared-memory configuration the SM is not in.
inline int smem_carveout_pct() {
  static const int v = getenv("DVIS_SMEM_CARVEOUT") ? atoi(getenv("DVIS_SMEM_CARVEOUT")) : -1;
  return v;
}
template <typename K>
inline void prefer_carveout(K kern) {
#ifndef DVIS_SIMT_EMULATION
  const int v = smem_carveout_pct();
  if (v >= 0) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, v);
#endif
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- device helpers ----------------------------------------------------------------------------
template <typename T>
struct Vec16;  // 16-byte vector of T
template <>
struct Vec16<float> {
  static constexpr int N = 4;
  float4 v;
  __device__ __forceinline__ float get(int i) const { return (&v.x)[i]; }
};
template <>
struct Vec16<double> {
  static constexpr int N = 2;
  double2 v;
  __device__ __forceinline__ double get(int i) const { return (&v.x)[i]; }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  uint4 v;
  __device__ __forceinline__ float get(int i) const {
    const uint32_t w = (&v.x)[i >> 1];
    return __uint_as_float((i & 1) ? (w & 0xffff0000u) : (w << 16));
  }
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&t);
}

}  // namespace dvis
