"""Micro-benchmark of the fused post-processing kernels (csrc/postproc.cu) vs the reference's torch sequence
(F.interpolate -> crop -> F.interpolate -> threshold, on the same B200) at the BASELINE metric size: 720p, T=16,
max_num=10 instances kept.  Writes gpurun_out/postproc_microbench.json.

Algorithmic bytes (vis masks) = selected stride-4 logits read once (n*T*h*w*sizeof) + bool result written once
(n*T*Ho*Wo); the torch sequence additionally writes and re-reads n*T*H1*W1 and n*T*Ho*Wo fp32 intermediates."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from dvis_plus_b200 import ops  # noqa: E402
from msda_microbench import timeit, PEAK_GBS  # noqa: E402


def torch_chain(m, first, img, out):
    x = F.interpolate(m, size=first, mode="bilinear", align_corners=False)[:, :, :img[0], :img[1]]
    return F.interpolate(x, size=out, mode="bilinear", align_corners=False) > 0


def main():
    quick = "--quick" in sys.argv
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    res = []
    cases = [("720p_identity", 16, (184, 320), (736, 1280), (720, 1280), (720, 1280)),
             ("480p_to_720p", 16, (120, 216), (480, 864), (480, 854), (720, 1280))]
    for name, T, (h, w), first, img, out in cases:
        Q, n = 200, 10
        g = torch.Generator(device="cuda").manual_seed(0)
        coarse = torch.randn(Q, T, h // 8, w // 8, device="cuda", generator=g) * 6
        logits = F.interpolate(coarse, size=(h, w), mode="bicubic", align_corners=False) - 1.0
        sel = torch.randperm(Q, device="cuda", generator=g)[:n]
        for dt in (torch.float32, torch.bfloat16):
            m = logits.to(dt)
            ref = torch_chain(m[sel].float(), first, img, out)
            ours = ops.vis_masks(m, sel, first, img, out)
            mismatch = (ours != ref).float().mean().item()
            med, best = timeit(lambda: ops.vis_masks(m, sel, first, img, out), iters=10 if quick else 30, flush=flush)
            alg = n * T * (h * w * m.element_size() + out[0] * out[1])
            row = dict(kernel="dvis_vis_masks", case=name, logits=str(dt).split(".")[-1], us_median=round(med, 1), us_min=round(best, 1),
                       algorithmic_MB=round(alg / 1e6, 2), achieved_GBs=round(alg / med / 1e3, 1),
                       frac_of_measured_hbm=round(alg / med / 1e3 / PEAK_GBS, 4), mismatch_vs_torch_cuda=mismatch)
            if dt == torch.float32:
                tmed, _ = timeit(lambda: torch_chain(m[sel], first, img, out), iters=5 if quick else 20, flush=flush)
                row["torch_sequence_us_median"] = round(tmed, 1)
                row["speedup_vs_torch_sequence"] = round(tmed / med, 2)
            # the shared-memory tile variant (DVIS_VIS_MASKS_TILED=1), same call
            os.environ["DVIS_VIS_MASKS_TILED"] = "1"
            same = bool(torch.equal(ops.vis_masks(m, sel, first, img, out), ours))
            tmed, _ = timeit(lambda: ops.vis_masks(m, sel, first, img, out), iters=10 if quick else 30, flush=flush)
            pmed, _ = timeit(lambda: ops.vis_masks(m, sel, first, img, out, packed=True), iters=10 if quick else 30, flush=flush)
            os.environ["DVIS_VIS_MASKS_TILED"] = "0"
            row.update(tiled_us_median=round(tmed, 1), tiled_packed_us_median=round(pmed, 1), tiled_equals_default=same)
            res.append(row)
            print(row, flush=True)
    if not quick:
        # vss / vps at a reduced frame count (they are O(Q) per pixel)
        T, (h, w), first, img, out = 4, (184, 320), (736, 1280), (720, 1280), (720, 1280)
        g = torch.Generator(device="cuda").manual_seed(1)
        m = torch.randn(100, T, h, w, device="cuda", generator=g)
        cls = ops.class_scores(torch.randn(100, 125, device="cuda", generator=g))
        med, _ = timeit(lambda: ops.vss_argmax(m, cls[:, :-1], first, img, out), iters=5, warmup=2, flush=flush)
        res.append(dict(kernel="dvis_vss_argmax", Q=100, K=124, frames=T, us_median=round(med, 1),
                        gflops=round(2.0 * 100 * 124 * T * out[0] * out[1] / med / 1e3, 1)))
        print(res[-1], flush=True)
        keep = torch.arange(0, 100, 3, device="cuda")
        sc = torch.rand(keep.numel(), device="cuda", generator=g)
        med, _ = timeit(lambda: ops.vps_argmax(m, keep, sc, first, img, out), iters=5, warmup=2, flush=flush)
        res.append(dict(kernel="dvis_vps_argmax", n_keep=int(keep.numel()), frames=T, us_median=round(med, 1)))
        print(res[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "postproc_microbench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
