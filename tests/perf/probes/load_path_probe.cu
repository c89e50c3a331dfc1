// Probe (not product): how fast can ONE CTA pull a stream of small tiles from L2 into shared memory on B200, by method?
// Each CTA reads `nchunks` chunks of `rows` x 128 bytes (row pitch `pitch` bytes in global memory, 144-byte pitch in smem)
// with 3 chunks in flight, and reports clock cycles per chunk.  Build: nvcc -arch=sm_100a -O3 -o load_path_probe load_path_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <stdint.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int METHOD>
__global__ void __launch_bounds__(128) probe(const char *src, int rows, int pitch, int nchunks, long long *out) {
  extern __shared__ __align__(128) char smem[];
  __shared__ uint64_t bars[4];
  const int tid = threadIdx.x;
  const int chunk_smem = rows * 144;
  const char *base = src + (size_t)blockIdx.x * nchunks * 128;   // different columns per CTA, same rows
  if (METHOD >= 4 && tid == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int kc) {
    char *dst = smem + (kc & 3) * chunk_smem;
    const char *g = base + (size_t)kc * 128;
    if (METHOD <= 2) {
      for (int c = tid; c < rows * 8; c += 128) {
        const int r = c >> 3, cc = c & 7;
        const char *gp = g + (size_t)r * pitch + cc * 16;
        if (METHOD == 0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s32(dst + r * 144 + cc * 16)), "l"(gp), "r"(16) : "memory");
        if (METHOD == 1) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(dst + r * 144 + cc * 16)), "l"(gp) : "memory");
        if (METHOD == 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s32(dst + r * 144 + cc * 16)), "l"(gp) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    } else if (METHOD == 3) {
      for (int c = tid; c < rows * 8; c += 128) {
        const int r = c >> 3, cc = c & 7;
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(g + (size_t)r * pitch + cc * 16));
        *reinterpret_cast<uint4 *>(dst + r * 144 + cc * 16) = v;
      }
    } else if (METHOD == 4) {       // one 128-byte bulk copy per row, issued by the lanes of warp 0
      if (tid < 32) {
        if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bars[kc & 3])), "r"(rows * 128) : "memory");
        __syncwarp();
        for (int r = tid; r < rows; r += 32)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst + r * 144)),
                       "l"(g + (size_t)r * pitch), "r"(128), "r"(s32(&bars[kc & 3]))
                       : "memory");
      }
    } else if (METHOD == 5) {       // ONE bulk copy per chunk: the tile is contiguous in global memory (pre-tiled weights)
      if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bars[kc & 3])), "r"(rows * 128) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
                     "l"(src + ((size_t)blockIdx.x * nchunks + kc) * rows * 128), "r"(rows * 128), "r"(s32(&bars[kc & 3]))
                     : "memory");
      }
    }
  };
  auto wait = [&](int kc) {
    if (METHOD <= 2) { asm volatile("cp.async.wait_group 2;" ::: "memory"); }
    else if (METHOD >= 4) {
      const uint32_t parity = (kc >> 2) & 1;
      asm volatile("{\n .reg .pred P1;\nW: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n @P1 bra D;\n bra W;\nD:\n}" ::"r"(s32(&bars[kc & 3])), "r"(parity) : "memory");
    }
  };
  for (int s = 0; s < 3; ++s) issue(s);
  __syncthreads();
  const long long t0 = clock64();
  unsigned acc = 0;
  for (int kc = 0; kc < nchunks; ++kc) {
    wait(kc);
    __syncthreads();
    if (kc + 3 < nchunks) issue(kc + 3);
    else if (METHOD <= 2) asm volatile("cp.async.commit_group;" ::: "memory");
    acc += *reinterpret_cast<unsigned *>(smem + (kc & 3) * chunk_smem + (tid % rows) * 144 + (tid & 7) * 16);
  }
  const long long t1 = clock64();
  if (tid == 0) out[blockIdx.x] = t1 - t0 + (acc == 0x12345678u);
}

template <int METHOD>
void run(const char *name, const char *src, int rows, int pitch, int nchunks, int ctas, long long *dout) {
  const int smem = 4 * rows * 144;
  cudaFuncSetAttribute(probe<METHOD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int it = 0; it < 3; ++it) probe<METHOD><<<ctas, 128, smem>>>(src, rows, pitch, nchunks, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-44s ERROR %s\n", name, cudaGetErrorString(e)); return; }
  long long h[1024];
  cudaMemcpy(h, dout, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
  double mx = 0, sum = 0;
  for (int i = 0; i < ctas; ++i) { sum += h[i]; if (h[i] > mx) mx = h[i]; }
  printf("%-44s rows=%3d ctas=%4d chunks=%3d: %7.1f cycles/chunk mean, %7.1f max  (%.1f B/clk/CTA)\n", name, rows, ctas, nchunks,
         sum / ctas / nchunks, mx / nchunks, rows * 128.0 / (sum / ctas / nchunks));
}

int main() {
  const size_t bytes = 256u << 20;
  char *src;
  long long *dout;
  cudaMalloc(&src, bytes);
  cudaMemset(src, 1, bytes);
  cudaMalloc(&dout, sizeof(long long) * 1024);
  for (int ctas : {56, 148, 444}) {
    for (int rows : {96, 64}) {
      const int pitch = 4096, nchunks = 32;     // K = 2048 bf16 rows
      run<0>("cp.async.cg 16B + src-size (current)", src, rows, pitch, nchunks, ctas, dout);
      run<1>("cp.async.cg 16B", src, rows, pitch, nchunks, ctas, dout);
      run<2>("cp.async.ca 16B", src, rows, pitch, nchunks, ctas, dout);
      run<3>("ld.global.nc.v4 + st.shared.v4", src, rows, pitch, nchunks, ctas, dout);
      run<4>("cp.async.bulk 128 B per row", src, rows, pitch, nchunks, ctas, dout);
      run<5>("cp.async.bulk one copy per chunk (pre-tiled)", src, rows, pitch, nchunks, ctas, dout);
    }
  }
  return 0;
}
