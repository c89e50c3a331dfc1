"""Tiny driver for `ncu`: a few launches of the mask GEMM at the 720p size."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dvis_plus_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
Q, C, H, W = 200, 256, 184, 320
emb = torch.randn(B, Q, C, device="cuda").bfloat16()
feat = torch.randn(B, C, H, W, device="cuda").to(torch.bfloat16, memory_format=torch.channels_last)
for _ in range(3):
    ops.mask_logits(emb, feat, torch.bfloat16)
    ops.mask_logits(emb, feat, torch.float32)
torch.cuda.synchronize()
