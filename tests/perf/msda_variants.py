"""Time the fused MSDA kernel (bf16 and f32 value) at T=16, 720p for the tuning variant selected by DVIS_MSDA_VARIANT."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from msda_microbench import make_inputs, timeit  # noqa: E402
from dvis_plus_b200 import ops  # noqa: E402
from dvis_plus_b200.locality import tiled_item_order  # noqa: E402

shapes = [(92, 160), (46, 80), (23, 40)]
sh_t = torch.tensor(shapes, device="cuda")
lsi = torch.cat((sh_t.new_zeros((1,)), sh_t.prod(1).cumsum(0)[:-1]))
M, L, P, N = 8, 3, 4, 16
S = sum(h * w for h, w in shapes)
value, loc, attn, offsets, logits, ref = make_inputs("encoder-like", N, shapes)
order = tiled_item_order(shapes, M, "cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
vb, ob, lb = value.bfloat16(), offsets.view(N, S, -1).bfloat16(), logits.view(N, S, -1).bfloat16()
of, lf = offsets.view(N, S, -1), logits.view(N, S, -1)
for name, fn in (("pair_bf16", lambda: ops.msda_pair_forward(vb, sh_t, lsi, ob, lb, ref, M, L, P, item_order=order)),
                 ("fused_bf16", lambda: ops.msda_fused_forward(vb, sh_t, lsi, ob, lb, ref, M, L, P, item_order=order)),
                 ("fused_f32", lambda: ops.msda_fused_forward(value, sh_t, lsi, of, lf, ref, M, L, P, item_order=order)),
                 ("plain_f32", lambda: ops.ms_deform_attn_forward(value, sh_t, lsi, loc, attn, 128, item_order=order))):
    med, best = timeit(fn, flush=flush)
    print(os.environ.get("DVIS_MSDA_VARIANT", "0"), name, f"{med:.1f} us median, {best:.1f} us min ({med / N:.1f} us/frame)", flush=True)
