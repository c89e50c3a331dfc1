"""ncu target: one launch each of the temporal-stage kernels at tracker / refiner shapes (after a warm-up launch).
    ncu --set full --import-source on -k regex:'small_linear|flash_attn' -o gpurun_out/ncu_temporal python tests/perf/temporal_profile_target.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dvis_plus_b200 import ops  # noqa: E402

dev = "cuda"
for M, N, K in ((200, 512, 2048), (200, 1536, 512), (3200, 1536, 512)):
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    for _ in range(2):
        ops.linear_small(w, b, x=x)
src0 = torch.randn(200, 512, device=dev)
g = torch.ones(512, device=dev)
w = torch.randn(1536, 512, device=dev).to(torch.bfloat16)
for _ in range(2):
    ops.linear_small(w, None, src0=src0, ln0=(g, g), src1=src0, ln1=(g, g), want_side0=True, want_side1=True)
for B in (1, 16):
    q = torch.randn(B, 200, 8, 64, device=dev).to(torch.bfloat16)
    for _ in range(2):
        ops.flash_attn(q, q, q, 1 / math.sqrt(64))
torch.cuda.synchronize()
