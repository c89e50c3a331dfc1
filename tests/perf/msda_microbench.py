"""Micro-benchmark of the MSDA forward kernels on one B200 (run under gpurun; writes gpurun_out/msda_microbench.json).

Two sampling regimes at the BASELINE 720p size:
  encoder-like : reference points = pixel centres, offsets = the module's initial ring bias (+ small noise)
  uniform      : uniformly random locations (OPS/test.py:37) -- worst-case gather
Compared on the same inputs: our kernel without / with the tiled locality schedule, the fused variant
(f32 and bf16 value), and the reference's own CUDA kernel built for sm_100a (oracle/_ref), if present.
"""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dvis_plus_b200 import ops  # noqa: E402
from dvis_plus_b200.locality import tiled_item_order  # noqa: E402
from oracle import ref_cuda_binding as refcuda  # noqa: E402

PEAK_GBS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, iters=50, warmup=10, flush=None):   # 10 warm-up + >= 50 timed iterations, median and min
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def make_inputs(regime, N, shapes, M=8, D=32, P=4, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    value = torch.randn(N, S, M, D, device="cuda", generator=g)
    logits = torch.randn(N, S, M, L * P, device="cuda", generator=g)
    if regime == "uniform":
        loc = torch.rand(N, S, M, L, P, 2, device="cuda", generator=g)
        ref = torch.zeros(N, S, L, 2, device="cuda")
        sh = torch.tensor([[w, h] for h, w in shapes], device="cuda", dtype=torch.float32)
        offsets = loc * sh[None, None, None, :, None, :]
    else:
        pts = []
        for h, w in shapes:
            ys = (torch.arange(h, device="cuda") + 0.5) / h
            xs = (torch.arange(w, device="cuda") + 0.5) / w
            gy, gx = torch.meshgrid(ys, xs, indexing="ij")
            pts.append(torch.stack((gx.reshape(-1), gy.reshape(-1)), -1))
        ref = torch.cat(pts, 0)[None, :, None, :].expand(N, -1, L, -1).contiguous()
        th = torch.arange(M, dtype=torch.float32, device="cuda") * (2 * math.pi / M)
        grid = torch.stack([th.cos(), th.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(M, 1, 1, 2).repeat(1, L, P, 1)
        for i in range(P):
            grid[:, :, i, :] *= i + 1
        offsets = grid[None, None].expand(N, S, -1, -1, -1, -1) + 0.5 * torch.randn(N, S, M, L, P, 2, device="cuda", generator=g)
        sh = torch.tensor([[w, h] for h, w in shapes], device="cuda", dtype=torch.float32)
        loc = ref[:, :, None, :, None, :] + offsets / sh[None, None, None, :, None, :]
    attn = logits.softmax(-1).view(N, S, M, L, P)
    return value, loc.contiguous(), attn.contiguous(), offsets.contiguous(), logits, ref


def main():
    shapes = [(92, 160), (46, 80), (23, 40)]
    sh_t = torch.tensor(shapes, device="cuda")
    lsi = torch.cat((sh_t.new_zeros((1,)), sh_t.prod(1).cumsum(0)[:-1]))
    M, D, L, P = 8, 32, 3, 4
    S = sum(h * w for h, w in shapes)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    results = []
    for N in (1, 16):
        alg_bytes = 4 * N * (S * M * D + S * M * L * P * 3 + S * M * D)
        for regime in ("encoder-like", "uniform"):
            value, loc, attn, offsets, logits, ref = make_inputs(regime, N, shapes)
            order = tiled_item_order(shapes, M, "cuda")
            offs_flat = offsets.view(N, S, -1)
            lg_flat = logits.view(N, S, -1)
            vb = value.bfloat16()
            ob16, lb16 = offs_flat.bfloat16(), lg_flat.bfloat16()
            variants = {
                "ours_plain": lambda: ops.ms_deform_attn_forward(value, sh_t, lsi, loc, attn, 128),
                "ours_plain_tiled": lambda: ops.ms_deform_attn_forward(value, sh_t, lsi, loc, attn, 128, item_order=order),
                "ours_fused_f32_tiled": lambda: ops.msda_fused_forward(value, sh_t, lsi, offs_flat, lg_flat, ref, M, L, P, item_order=order),
                "ours_fused_bf16_tiled": lambda: ops.msda_fused_forward(vb, sh_t, lsi, offs_flat, lg_flat, ref, M, L, P, item_order=order),
                # the encoder's production form: value, offsets, logits and output all bf16
                "ours_fused_bf16_params_bf16_tiled": lambda: ops.msda_fused_forward(vb, sh_t, lsi, ob16, lb16, ref, M, L, P, item_order=order),
            }
            if refcuda.available():
                variants["reference_cuda_sm100"] = lambda: refcuda.forward(value, sh_t, lsi, loc, attn)
            base = None
            for name, fn in variants.items():
                out = fn().float()
                if base is None:
                    base = out
                err = (out - base).abs().max().item()
                med, best = timeit(fn, flush=flush)
                results.append(dict(kernel=name, regime=regime, frames=N, us_median=round(med, 2), us_min=round(best, 2),
                                    algorithmic_MB=round(alg_bytes / 1e6, 2), achieved_GBs=round(alg_bytes / med / 1e3, 1),
                                    frac_of_measured_hbm=round(alg_bytes / med / 1e3 / PEAK_GBS, 4), max_abs_diff_vs_plain=err))
                print(results[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "msda_microbench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
