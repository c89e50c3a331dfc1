"""Tiny driver for `ncu`: launches the bf16 fused MSDA variants a few times at the 720p size (N frames, chosen regime)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from msda_microbench import make_inputs  # noqa: E402
from dvis_plus_b200 import ops  # noqa: E402
from dvis_plus_b200.locality import tiled_item_order  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
regime = sys.argv[2] if len(sys.argv) > 2 else "encoder-like"
shapes = [(92, 160), (46, 80), (23, 40)]
sh_t = torch.tensor(shapes, device="cuda")
lsi = torch.cat((sh_t.new_zeros((1,)), sh_t.prod(1).cumsum(0)[:-1]))
M, L, P = 8, 3, 4
S = sum(h * w for h, w in shapes)
value, loc, attn, offsets, logits, ref = make_inputs(regime, N, shapes)
order = tiled_item_order(shapes, M, "cuda")
vb, ob, lb = value.bfloat16(), offsets.view(N, S, -1).bfloat16(), logits.view(N, S, -1).bfloat16()
for _ in range(3):
    ops.msda_fused_forward(vb, sh_t, lsi, ob, lb, ref, M, L, P, item_order=order)
    ops.msda_pair_forward(vb, sh_t, lsi, ob, lb, ref, M, L, P, item_order=order)
torch.cuda.synchronize()
