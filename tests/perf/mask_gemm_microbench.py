"""Micro-benchmark of the tcgen05 mask-logit GEMM vs torch.einsum (cuBLAS) on one B200.
Writes gpurun_out/mask_gemm_microbench.json.  Algorithmic bytes = (C*HW + Q*HW + Q*C) * sizeof per frame."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from dvis_plus_b200 import ops  # noqa: E402
from msda_microbench import timeit, PEAK_GBS  # noqa: E402


def main():
    H, W, C = 184, 320, 256
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    res = []
    for B, Q in ((1, 100), (1, 200), (16, 200)):
        emb = torch.randn(B, Q, C, device="cuda").bfloat16()
        feat = torch.randn(B, C, H, W, device="cuda").to(torch.bfloat16, memory_format=torch.channels_last)
        feat_nchw = feat.contiguous()
        feat32 = feat_nchw.float()
        emb32 = emb.float()
        feat32_cl = feat32.contiguous(memory_format=torch.channels_last)
        variants = {
            "ours_tcgen05_out_bf16": (lambda: ops.mask_logits(emb, feat, torch.bfloat16), 2, 2),
            "ours_tcgen05_out_f32": (lambda: ops.mask_logits(emb, feat, torch.float32, operand_dtype=torch.bfloat16), 2, 4),
            "ours_tcgen05_tf32_operands_out_f32": (lambda: ops.mask_logits(emb32, feat32_cl, torch.float32, operand_dtype=torch.float32), 4, 4),
            "cublas_einsum_bf16_nchw": (lambda: torch.einsum("bqc,bchw->bqhw", emb, feat_nchw), 2, 2),
            "cublas_einsum_fp32_nchw": (lambda: torch.einsum("bqc,bchw->bqhw", emb32, feat32), 4, 4),
        }
        ref = torch.einsum("bqc,bchw->bqhw", emb32, feat32)
        for name, (fn, in_b, out_b) in variants.items():
            err = (fn().float() - ref).abs().max().item() / ref.abs().max().item()
            med, best = timeit(fn, flush=flush)
            bytes_ = B * (C * H * W * in_b + Q * H * W * out_b + Q * C * in_b)
            flops = 2.0 * B * Q * C * H * W
            res.append(dict(kernel=name, frames=B, Q=Q, us_median=round(med, 2), us_min=round(best, 2),
                            algorithmic_MB=round(bytes_ / 1e6, 2), achieved_GBs=round(bytes_ / med / 1e3, 1),
                            frac_of_measured_hbm=round(bytes_ / med / 1e3 / PEAK_GBS, 4),
                            tflops=round(flops / med / 1e6, 1), rel_err_vs_fp32=err))
            print(res[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "mask_gemm_microbench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
