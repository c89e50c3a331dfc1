"""Build oracle/_ref/libref_msda.so from the reference's CUDA header in /root/reference (this container only).

ORACLE / BASELINE INFRASTRUCTURE ONLY.  The GPU box has no /root/reference: it uses the prebuilt .so that
travels with the snapshot (oracle/_ref/ is git-ignored but not gpurun-ignored).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CUH = "/root/reference/DVIS_Plus/mask2former/modeling/pixel_decoder/ops/src/cuda/ms_deform_im2col_cuda.cuh"
OUT_DIR = os.path.join(os.path.dirname(HERE), "_ref")
OUT = os.path.join(OUT_DIR, "libref_msda.so")


def build(force=False):
    if not os.path.exists(REF_CUH):
        return OUT if os.path.exists(OUT) else None
    if os.path.exists(OUT) and not force:
        return OUT
    import torch
    from torch.utils.cpp_extension import include_paths
    os.makedirs(OUT_DIR, exist_ok=True)
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC", "-w", f'-DREF_CUH_PATH="{REF_CUH}"', "-DCUDA_HAS_FP16=1",
           "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    cmd += [f"-I{p}" for p in include_paths("cuda")]
    cmd += [os.path.join(HERE, "shim.cu"), "-o", OUT, f"-L{tlib}", "-lc10", "-lc10_cuda", "-ltorch_cpu",
            "-Xlinker", f"-rpath={tlib}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("reference CUDA kernel build failed")
    return OUT


if __name__ == "__main__":
    print(build(force=True))
