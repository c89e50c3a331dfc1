"""Head-major MSDA gather at the bench shape under DVIS_MSDA_HM_VARIANT (one process per variant: the switch is read once)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from dvis_plus_b200 import ops
from dvis_plus_b200.locality import tiled_item_order

T = int(sys.argv[1]) if len(sys.argv) > 1 else 16
shapes = ((92, 160), (46, 80), (23, 40))
M, D, L, P = 8, 32, 3, 4
S = sum(h * w for h, w in shapes)
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
value_hm = torch.randn(T, M, S, D, generator=g, device=dev).bfloat16()
sh = torch.as_tensor(shapes, dtype=torch.long, device=dev)
lsi = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
fused = (torch.randn(T, S, M * L * P * 3, generator=g, device=dev) * 1.5).bfloat16()
n_off = M * L * P * 2
ys, xs = [], []
for h, w_ in shapes:
    yy, xx = torch.meshgrid((torch.arange(h, device=dev) + 0.5) / h, (torch.arange(w_, device=dev) + 0.5) / w_, indexing="ij")
    ys.append(yy.reshape(-1)); xs.append(xx.reshape(-1))
ref = torch.stack([torch.cat(xs), torch.cat(ys)], -1)[None, :, None, :].expand(T, S, L, 2).contiguous()
order = tiled_item_order(shapes, M, torch.device(dev))
fn = lambda: ops.msda_fused_forward_hm(value_hm, sh, lsi, fused[..., :n_off], fused[..., n_off:], ref, L, P, item_order=order)
for _ in range(5):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    fn()
e1.record()
torch.cuda.synchronize()
print(json.dumps({"variant": os.environ.get("DVIS_MSDA_HM_VARIANT", "0"), "T": T, "us": round(e0.elapsed_time(e1) / 20 * 1e3, 1)}))
