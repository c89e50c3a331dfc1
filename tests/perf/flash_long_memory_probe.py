"""Masked cross-attention of the predictor (16 frames x 8 heads x 200 queries over 920 / 3 680 / 14 720 pixels, bit mask):
flash_attn tilings (DVIS_FLASH_VARIANT 0 auto, 1 = 128-row tiles, 2 = 64-row tiles, 3 = key split), us per call."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

from dvis_plus_b200 import ops


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


res = {}
B, H, dh, Q = int(sys.argv[1]) if len(sys.argv) > 1 else 16, 8, 32, 200
for Lk in (920, 3680, 14720):
    q = torch.randn(B, Q, H, dh, device="cuda").bfloat16()
    k = torch.randn(B, Lk, H, dh, device="cuda").bfloat16()
    v = torch.randn(B, Lk, H, dh, device="cuda").bfloat16()
    nbytes = (Lk + 63) // 64 * 8
    bits = torch.randint(0, 256, (B, Q, nbytes), device="cuda", dtype=torch.uint8)
    row = {}
    for var in (0, 1, 2, 3, 4):
        os.environ["DVIS_FLASH_VARIANT"] = str(var)
        row[f"variant_{var}_us"] = round(timeit(lambda: ops.flash_attn(q, k, v, dh ** -0.5, mask_bits=bits)), 1)
    os.environ["DVIS_FLASH_VARIANT"] = "0"
    res[f"B{B}_Lk{Lk}"] = row
    print(B, Lk, row, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"flash_long_memory_probe_B{B}.json"), "w"), indent=1)
