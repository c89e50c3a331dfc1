"""60-second device check of the kernels added after the round's last full GPU run: the two-level strip walker
(two-resize VIS masks), the bit-packed variants and the rectangular Hungarian kernel -- against torch on the same GPU /
SciPy, plus a timing of the two-resize kernel at the 480p -> 720p size.  Writes gpurun_out/quick_check.json as it goes."""
import json
import os
import sys
import time

t0 = time.time()
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dvis_plus_b200 import ops  # noqa: E402

res = {"import_s": round(time.time() - t0, 1)}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)


def dump():
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "quick_check.json"), "w"), indent=1)
    print(res, flush=True)


def chain(m, first, img, out):
    x = F.interpolate(m, size=first, mode="bilinear", align_corners=False)[:, :, :img[0], :img[1]]
    return F.interpolate(x, size=out, mode="bilinear", align_corners=False) > 0


g = torch.Generator(device="cuda").manual_seed(0)
T, (h, w), first, img, out = 16, (120, 216), (480, 864), (480, 854), (720, 1280)
coarse = torch.randn(40, T, h // 8, w // 8, device="cuda", generator=g) * 6
logits = F.interpolate(coarse, size=(h, w), mode="bicubic", align_corners=False) - 1.0
sel = torch.randperm(40, device="cuda", generator=g)[:10]
ours = ops.vis_masks(logits, sel, first, img, out)
ref = chain(logits[sel], first, img, out)
res["two_stage_mismatch_vs_torch_cuda"] = (ours != ref).float().mean().item()
dump()
packed = ops.vis_masks(logits, sel, first, img, out, packed=True)
res["two_stage_packed_equals_bytes"] = bool(torch.equal(ops.unpack_masks(packed.cpu(), out[1]), ours.cpu()))
ident = ops.vis_masks(logits, sel, first, first, first)
res["strip_packed_equals_bytes"] = bool(torch.equal(ops.unpack_masks(ops.vis_masks(logits, sel, first, first, first, packed=True).cpu(), first[1]),
                                                    ident.cpu()))
dump()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for name, fn in (("two_stage_us", lambda: ops.vis_masks(logits, sel, first, img, out)),
                 ("two_stage_packed_us", lambda: ops.vis_masks(logits, sel, first, img, out, packed=True)),
                 ("torch_sequence_us", lambda: chain(logits[sel], first, img, out))):
    ts = []
    for i in range(8):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    res[name] = round(sorted(ts[2:])[len(ts[2:]) // 2], 1)
    dump()
import numpy as np
from scipy.optimize import linear_sum_assignment
ok = True
for rows, cols in ((30, 50), (50, 30), (200, 200)):
    c = torch.rand(rows, cols, generator=torch.Generator().manual_seed(rows))
    got = ops.lap_rect(c.cuda()).cpu().numpy()
    r, cc = linear_sum_assignment(c.numpy())
    want = np.full(rows, -1, dtype=np.int64); want[r] = cc
    ok &= bool(np.array_equal(got, want))
res["lap_rect_equals_scipy"] = ok
res["total_s"] = round(time.time() - t0, 1)
dump()
