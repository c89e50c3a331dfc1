"""TEST INFRASTRUCTURE ONLY.  Turn CUDA-core kernels of dvis_plus_b200/csrc into a g++ build on top of the SIMT emulator
(tests/simt/simt_shim.h): tests/simt/_build/libdvis_simt.so exports the SAME C-ABI entry points as libdvis_b200.so for the
translated files, operating on host memory.

Source transforms (textual, the kernels themselves are untouched):
  * project headers (`#include "x.cuh"`) are inlined once; <cuda_runtime.h> / <cuda_bf16.h> are replaced by the shim
  * `kernel<<<grid, block, smem, stream>>>(args);`   ->  SIMT_LAUNCH((kernel), grid, block, smem, stream, args);
  * `extern __shared__ T name[];`                    ->  T *name = reinterpret_cast<T *>(simt::dyn_smem());
"""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "dvis_plus_b200", "csrc")
INCLUDE = os.path.join(ROOT, "include")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libdvis_simt_%s.so" % os.environ["SIMT_SANITIZE"] if os.environ.get("SIMT_SANITIZE") else "libdvis_simt.so")
# every CUDA-core kernel file of the library; csrc/mask_gemm.cu (tcgen05 / TMEM / TMA) and csrc/api.cu (driver entry
# points) have no CPU meaning and stay out
FILES = ["postproc.cu", "lap.cu", "msda_forward.cu", "msda_backward.cu", "layernorm.cu", "groupnorm.cu", "mask_aux.cu",
         "msda_pair.cu", "flash_attn.cu", "small_linear.cu"]


# What csrc/api.cu provides in the real library, plus TEST DOUBLES (plain loops, not emulation) for the entry points of
# csrc/mask_gemm.cu -- the tcgen05 / TMEM / TMA kernel has no CPU meaning; the doubles let module-level tests run the
# product's host code end to end on the emulated device (tests/simt/emulated_device.py).
GLUE = r'''
#include "simt_shim.h"
#include "dvis_b200.h"
namespace dvis { char *last_error_buffer() { static thread_local char buf[512] = {0}; return buf; } }
extern "C" const char *dvis_last_error(void) { return dvis::last_error_buffer(); }
extern "C" int dvis_abi_version(void) { return DVIS_B200_ABI_VERSION; }
extern "C" void simt_set_jitter(int one_in) { simt::g_jitter = one_in; }
extern "C" int dvis_set_pdl(int) { return 0; }

namespace {
template <typename TO, typename TI = __nv_bfloat16>
void mask_gemm_double(const TI *emb, int64_t emb_batch, const TI *feat, int B, int Q, int C, int64_t HW,
                      TO *out, int64_t out_batch, bool bias) {
  for (int b = 0; b < B; ++b)
    for (int q = 0; q < Q; ++q) {
      bool any_open = false;
      TO *row = out + b * out_batch + (int64_t)q * HW;
      for (int64_t p = 0; p < HW; ++p) {
        float acc = 0.f;
        for (int c = 0; c < C; ++c) acc += float(emb[b * emb_batch + (int64_t)q * C + c]) * float(feat[((int64_t)b * HW + p) * C + c]);
        if (bias) { const bool open = !(acc < 0.f); any_open |= open; acc = open ? 0.f : -INFINITY; }
        row[p] = TO(acc);
      }
      if (bias && !any_open) for (int64_t p = 0; p < HW; ++p) row[p] = TO(0.f);
    }
}
}  // namespace
extern "C" int dvis_mask_logits_strided(const void *emb, int64_t emb_batch_stride, const void *feat, int B, int Q, int C, int64_t HW,
                                        void *out, int64_t out_batch_stride, int out_dtype, void *) {
  const auto *e = static_cast<const __nv_bfloat16 *>(emb), *f = static_cast<const __nv_bfloat16 *>(feat);
  if (out_dtype == DVIS_F32) mask_gemm_double(e, emb_batch_stride, f, B, Q, C, HW, static_cast<float *>(out), out_batch_stride, false);
  else mask_gemm_double(e, emb_batch_stride, f, B, Q, C, HW, static_cast<__nv_bfloat16 *>(out), out_batch_stride, false);
  return 0;
}
extern "C" int dvis_mask_logits(const void *emb, const void *feat, int B, int Q, int C, int64_t HW, void *out, int out_dtype, void *s) {
  return dvis_mask_logits_strided(emb, (int64_t)Q * C, feat, B, Q, C, HW, out, (int64_t)Q * HW, out_dtype, s);
}
extern "C" int dvis_mask_attn_bits(const void *emb, const void *feat, int B, int Q, int C, int64_t HW, void *bits, int64_t row_bytes,
                                   int *, void *) {
  const auto *e = static_cast<const __nv_bfloat16 *>(emb), *f = static_cast<const __nv_bfloat16 *>(feat);
  auto *out = static_cast<uint8_t *>(bits);
  for (int b = 0; b < B; ++b)
    for (int q = 0; q < Q; ++q) {
      uint8_t *row = out + ((int64_t)b * Q + q) * row_bytes;
      std::memset(row, 0xff, row_bytes);                         // bytes past the row's pixels are unspecified in the product
      bool any_open = false;
      for (int64_t p = 0; p < HW; ++p) {
        float acc = 0.f;
        for (int c = 0; c < C; ++c) acc += float(e[((int64_t)b * Q + q) * C + c]) * float(f[((int64_t)b * HW + p) * C + c]);
        const bool open = !(acc < 0.f);
        any_open |= open;
        if (open) row[p >> 3] &= uint8_t(~(1u << (p & 7)));
      }
      if (!any_open) std::memset(row, 0, row_bytes);
    }
  return 0;
}
extern "C" int dvis_mask_logits_clip(const void *emb, const void *feat, int T, int Q, int C, int64_t HW, void *out, int out_dtype, void *) {
  const auto *e = static_cast<const __nv_bfloat16 *>(emb), *f = static_cast<const __nv_bfloat16 *>(feat);
  for (int t = 0; t < T; ++t)
    for (int q = 0; q < Q; ++q)
      for (int64_t p = 0; p < HW; ++p) {
        float acc = 0.f;
        for (int c = 0; c < C; ++c) acc += float(e[((int64_t)t * Q + q) * C + c]) * float(f[((int64_t)t * HW + p) * C + c]);
        const int64_t o = ((int64_t)q * T + t) * HW + p;
        if (out_dtype == DVIS_F32) static_cast<float *>(out)[o] = acc; else static_cast<__nv_bfloat16 *>(out)[o] = __nv_bfloat16(acc);
      }
  return 0;
}
// fp32-operand (TF32 on the device) variants: exact fp32 here
extern "C" int dvis_mask_logits_tf32(const void *emb, const void *feat, int B, int Q, int C, int64_t HW, void *out, void *) {
  mask_gemm_double<float, float>(static_cast<const float *>(emb), (int64_t)Q * C, static_cast<const float *>(feat), B, Q, C, HW,
                                 static_cast<float *>(out), (int64_t)Q * HW, false);
  return 0;
}
extern "C" int dvis_mask_attn_bias_tf32(const void *emb, const void *feat, int B, int Q, int C, int64_t HW, void *bias, int *, void *) {
  mask_gemm_double<float, float>(static_cast<const float *>(emb), (int64_t)Q * C, static_cast<const float *>(feat), B, Q, C, HW,
                                 static_cast<float *>(bias), (int64_t)Q * HW, true);
  return 0;
}
// doubles of csrc/linear_tc.cu (tcgen05): plain loops with an fp32 accumulator
namespace {
void linear_row_double(const __nv_bfloat16 *x, const __nv_bfloat16 *w, const float *bias, int N, int K, float *y) {
  for (int n = 0; n < N; ++n) {
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc += float(x[k]) * float(w[(int64_t)n * K + k]);
    y[n] = acc + (bias ? bias[n] : 0.f);
  }
}
}  // namespace
extern "C" int dvis_linear_tc(const void *x, int64_t ldx, const void *w, const float *bias, int relu, int rows, int N, int K, void *y,
                              int64_t ldy, void *) {
  std::vector<float> row(N);
  for (int64_t r = 0; r < rows; ++r) {
    linear_row_double(static_cast<const __nv_bfloat16 *>(x) + r * ldx, static_cast<const __nv_bfloat16 *>(w), bias, N, K, row.data());
    for (int n = 0; n < N; ++n) static_cast<__nv_bfloat16 *>(y)[r * ldy + n] = __nv_bfloat16(relu ? std::max(row[n], 0.f) : row[n]);
  }
  return 0;
}
extern "C" int dvis_linear_tc_heads(const void *x, int64_t ldx, const void *w, const float *bias, int batch, int S, int N, int K,
                                    const uint8_t *row_mask, void *value_hm, void *) {
  std::vector<float> row(N);
  const int heads = N / 32;
  for (int64_t r = 0; r < (int64_t)batch * S; ++r) {
    linear_row_double(static_cast<const __nv_bfloat16 *>(x) + r * ldx, static_cast<const __nv_bfloat16 *>(w), bias, N, K, row.data());
    const int64_t n = r / S, s = r % S;
    for (int c = 0; c < N; ++c)
      static_cast<__nv_bfloat16 *>(value_hm)[((n * heads + c / 32) * S + s) * 32 + c % 32] =
          __nv_bfloat16((row_mask && row_mask[r]) ? 0.f : row[c]);
  }
  return 0;
}
extern "C" int dvis_linear_tc_add_ln(const void *x, int64_t ldx, const void *w, const float *bias, const float *residual,
                                     const float *gamma, const float *beta, float eps, int rows, int N, int K, const float *pos,
                                     int pos_rows, float *out_f32, void *out_lp, void *out_lp_pos, void *) {
  std::vector<float> row(N);
  for (int64_t r = 0; r < rows; ++r) {
    linear_row_double(static_cast<const __nv_bfloat16 *>(x) + r * ldx, static_cast<const __nv_bfloat16 *>(w), bias, N, K, row.data());
    float mean = 0.f, var = 0.f;
    for (int n = 0; n < N; ++n) { row[n] += residual[r * N + n]; mean += row[n]; }
    mean /= N;
    for (int n = 0; n < N; ++n) var += (row[n] - mean) * (row[n] - mean);
    const float rstd = 1.f / std::sqrt(var / N + eps);
    for (int n = 0; n < N; ++n) {
      const float y = (row[n] - mean) * rstd * gamma[n] + beta[n];
      if (out_f32) out_f32[r * N + n] = y;
      if (out_lp) static_cast<__nv_bfloat16 *>(out_lp)[r * N + n] = __nv_bfloat16(y);
      if (out_lp_pos) static_cast<__nv_bfloat16 *>(out_lp_pos)[r * N + n] = __nv_bfloat16(y + pos[(r % pos_rows) * N + n]);
    }
  }
  return 0;
}
extern "C" int dvis_mask_attn_bias(const void *emb, const void *feat, int B, int Q, int C, int64_t HW, void *bias, int bias_dtype, int *,
                                   void *) {
  const auto *e = static_cast<const __nv_bfloat16 *>(emb), *f = static_cast<const __nv_bfloat16 *>(feat);
  if (bias_dtype == DVIS_F32) mask_gemm_double(e, (int64_t)Q * C, f, B, Q, C, HW, static_cast<float *>(bias), (int64_t)Q * HW, true);
  else mask_gemm_double(e, (int64_t)Q * C, f, B, Q, C, HW, static_cast<__nv_bfloat16 *>(bias), (int64_t)Q * HW, true);
  return 0;
}
'''


def _split_top_level(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "(<[{":
            depth += 1
        elif ch in ")>]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def _inline(path, seen):
    out = []
    for line in open(path).read().splitlines():
        m = re.match(r'\s*#include\s+"([\w.]+)"', line)
        if m:
            name = m.group(1)
            for d in (CSRC, INCLUDE):
                if os.path.exists(os.path.join(d, name)):
                    if name not in seen:
                        seen.add(name)
                        out.append(f"// ---- inlined {name}")
                        out.extend(_inline(os.path.join(d, name), seen))
                    break
            else:
                raise RuntimeError(f"{path}: cannot resolve include {name}")
            continue
        if re.match(r"\s*#include\s+<cuda(_runtime|_bf16|_fp16)\.h>", line):
            continue
        if re.match(r"\s*#pragma\s+once", line):
            continue
        out.append(line)
    return out


def translate(cu):
    src = "\n".join(_inline(os.path.join(CSRC, cu), set()))
    src = re.sub(r"extern\s+__shared__\s+([\w ]+?)\s+(\w+)\s*\[\s*\]\s*;",
                 r"\1 *\2 = reinterpret_cast<\1 *>(simt::dyn_smem());", src)

    def launch(m):
        cfg = _split_top_level(m.group(2))
        assert len(cfg) == 4, f"{cu}: launch configuration must have 4 entries: {m.group(0)}"
        return f"SIMT_LAUNCH(({m.group(1)}), {cfg[0]}, {cfg[1]}, {cfg[2]}, {cfg[3]}, {m.group(3)});"

    src, n = re.subn(r"([A-Za-z_][\w:]*(?:<[^<>;()]*>)?)\s*<<<(.*?)>>>\s*\((.*?)\)\s*;", launch, src, flags=re.S)
    assert n > 0, f"{cu}: no kernel launch found"
    return f'#include "simt_shim.h"\n#define DVIS_SIMT_EMULATION 1\n{src}\n', n


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))] + \
           [os.path.join(INCLUDE, "dvis_b200.h"), os.path.join(HERE, "simt_shim.h"), os.path.abspath(__file__)]
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(d) for d in deps):
        return OUT
    units = []
    for cu in FILES:
        text, _ = translate(cu)
        path = os.path.join(OUT_DIR, cu[:-3] + "_simt.cpp")
        with open(path, "w") as fh:
            fh.write(text)
        units.append(path)
    glue = os.path.join(OUT_DIR, "glue_simt.cpp")
    with open(glue, "w") as fh:
        fh.write(GLUE)
    flags = ["-O1", "-std=c++20", "-fPIC", "-pthread", "-ffp-contract=off", "-Wno-unknown-pragmas", "-Wno-attributes",
             f"-I{HERE}", f"-I{INCLUDE}"]
    sanitize = os.environ.get("SIMT_SANITIZE")        # "address": out-of-bounds checks on every emulated load / store
    if sanitize:                                      # (run the tests with LD_PRELOAD=$(gcc -print-file-name=libasan.so))
        flags += [f"-fsanitize={sanitize}", "-fno-omit-frame-pointer", "-g"]

    def compile_one(src):
        obj = src[:-4] + (".%s.o" % sanitize if sanitize else ".o")
        r = subprocess.run(["g++"] + flags + ["-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed on {src}:\n{r.stdout[-3000:]}\n{r.stderr[-6000:]}")
        return obj

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(compile_one, units + [glue]))
    r = subprocess.run(["g++", "-shared", "-pthread"] + ([f"-fsanitize={sanitize}"] if sanitize else []) + objs + ["-o", OUT],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout[-3000:]}\n{r.stderr[-6000:]}")
    return OUT


if __name__ == "__main__":
    print(build(force=True))
