// ORACLE / BASELINE INFRASTRUCTURE ONLY -- never linked into the product library.
//
// C-ABI shim around the reference's own CUDA kernels, compiled from the sources where they lie under
// /root/reference (the header is #included by path; nothing is copied).  The reference's host launcher
// file (OPS/src/cuda/ms_deform_attn_cuda.cu) does not compile against torch 2.11 (`value.type()` at :69,:139),
// so this shim calls the header's launchers ms_deformable_im2col_cuda / ms_deformable_col2im_cuda
// (OPS/src/cuda/ms_deform_im2col_cuda.cuh:928-959, :961-1331) directly with the same arguments the reference
// host code passes (ms_deform_attn_cuda.cu:70-79, :140-153), one call per im2col_step chunk.
// Output: oracle/_ref/libref_msda.so (git-ignored; travels to the GPU box).  Used as "the reference CUDA
// kernel on the same B200" in tests/test_vs_reference_cuda_gpu.py and profiles/microbench.py.
#include REF_CUH_PATH

extern "C" int ref_msda_forward_f32(const float *value, const int64_t *shapes, const int64_t *lsi, const float *loc,
                                    const float *attn, int batch, int S, int M, int D, int L, int Lq, int P,
                                    int im2col_step, float *out, void *stream) {
  const int step = batch < im2col_step ? batch : im2col_step;
  if (batch % step) return 1;
  const long pv = (long)S * M * D, pl = (long)Lq * M * L * P * 2, pa = (long)Lq * M * L * P, po = (long)Lq * M * D;
  for (int n = 0; n < batch / step; ++n)
    ms_deformable_im2col_cuda<float>((cudaStream_t)stream, value + n * step * pv, shapes, lsi, loc + n * step * pl,
                                     attn + n * step * pa, step, S, M, D, L, Lq, P, out + n * step * po);
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

extern "C" int ref_msda_backward_f32(const float *value, const int64_t *shapes, const int64_t *lsi, const float *loc,
                                     const float *attn, const float *gout, int batch, int S, int M, int D, int L, int Lq,
                                     int P, int im2col_step, float *gvalue, float *gloc, float *gattn, void *stream) {
  const int step = batch < im2col_step ? batch : im2col_step;
  if (batch % step) return 1;
  const long pv = (long)S * M * D, pl = (long)Lq * M * L * P * 2, pa = (long)Lq * M * L * P, po = (long)Lq * M * D;
  for (int n = 0; n < batch / step; ++n)
    ms_deformable_col2im_cuda<float>((cudaStream_t)stream, gout + n * step * po, value + n * step * pv, shapes, lsi,
                                     loc + n * step * pl, attn + n * step * pa, step, S, M, D, L, Lq, P,
                                     gvalue + n * step * pv, gloc + n * step * pl, gattn + n * step * pa);
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}
