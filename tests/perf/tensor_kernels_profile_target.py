"""ncu target: one launch of each tcgen05 kernel of round 2 at the bench size (16 frames 720p, Q=200): mask GEMM (bf16 out, the
clip layout, the bit-mask epilogue, TF32 operands) and the head-major value projection."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dvis_plus_b200 import ops  # noqa: E402

T, Q, C, H, W = 16, 200, 256, 184, 320
g = torch.Generator(device="cuda").manual_seed(0)
emb = torch.randn(T, Q, C, generator=g, device="cuda").bfloat16()
feat = torch.randn(T, C, H, W, generator=g, device="cuda").to(torch.bfloat16, memory_format=torch.channels_last)
lvl = torch.randn(T, C, 92, 160, generator=g, device="cuda").to(torch.bfloat16, memory_format=torch.channels_last)
x = torch.randn(T, 19320, C, generator=g, device="cuda").bfloat16()
w = (torch.randn(C, C, generator=g, device="cuda") / 16).bfloat16()
b = torch.randn(C, generator=g, device="cuda")
for _ in range(2):
    ops.mask_logits(emb, feat, torch.bfloat16)
    ops.mask_logits_clip(emb, feat, torch.bfloat16)
    ops.mask_attn_bits(emb, lvl)
    ops.mask_logits(emb.float(), feat.float().contiguous(memory_format=torch.channels_last), torch.float32, operand_dtype=torch.float32)
    ops.linear_tc_heads(x, w, b)
torch.cuda.synchronize()
