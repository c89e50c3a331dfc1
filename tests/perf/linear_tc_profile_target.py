"""ncu target: one launch of each csrc/linear_tc.cu mode at the bench shape (16 x 19 320 tokens x 256)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from dvis_plus_b200 import ops

T, S, C = 16, 19320, 256
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(T, S, C, generator=g, device="cuda").bfloat16()
w = (torch.randn(C, C, generator=g, device="cuda") / 16).bfloat16()
b = torch.randn(C, generator=g, device="cuda")
res = torch.randn(T, S, C, generator=g, device="cuda")
gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
for _ in range(2):
    ops.linear_tc_add_layernorm(x, w, b, res, gamma, beta, 1e-5)
    ops.linear_tc_heads(x, w, b)
torch.cuda.synchronize()
