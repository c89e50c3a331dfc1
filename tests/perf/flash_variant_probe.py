"""flash_attn tilings (DVIS_FLASH_VARIANT: 0 auto, 1 = 128-row tiles, 2 = 64-row tiles, 3 = key split) vs cuDNN SDPA at the
batched 200-query shapes of the predictor's self-attention (dh 32) and the refiner's object / cross attention (dh 64), inside a
CUDA graph of 20 dependent calls (what the pipeline replays)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F

from dvis_plus_b200 import ops


def graph_us(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 / n * 1e3


res = {}
for B in (16, 8, 2):
    for dh in (32, 64):
        H = 8
        qkv = torch.randn(B, 200, 3, H, dh, device="cuda").bfloat16()
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
        row = {"sdpa_us": round(graph_us(lambda: F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))), 2)}
        for var in (0, 1, 2, 3):
            os.environ["DVIS_FLASH_VARIANT"] = str(var)
            row[f"flash_variant_{var}_us"] = round(graph_us(lambda: ops.flash_attn(q, k, v, dh ** -0.5)), 2)
        os.environ["DVIS_FLASH_VARIANT"] = "0"
        res[f"B{B}_dh{dh}"] = row
        print(B, dh, row, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "flash_variant_probe.json"), "w"), indent=1)
