"""Stage-by-stage timing of the offline clip pipeline at the BASELINE size on one B200 (writes
gpurun_out/pipeline_breakdown.json).  T=16 frames 720p, Q=200, Swin-L channel widths."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bench import build_models, synthetic_features  # noqa: E402
from dvis_plus_b200 import _lib  # noqa: E402
from dvis_plus_b200.modules.precision import set_precision  # noqa: E402


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


@torch.no_grad()
def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    set_precision("bf16")
    runner = build_models("cuda", queries=200)
    feats = {k: v.cuda() for k, v in synthetic_features(T, "swinl").items()}
    res = {}
    for it in range(3):
        torch.cuda.synchronize()
        c0 = _lib.launch_count
        w0 = time.perf_counter()
        e = [ev()]
        mf, _, ms = runner.pixel_decoder.forward_features(feats); e.append(ev())
        seg = runner.predictor(ms, mf); e.append(ev())
        block = runner.pack_queries(seg)
        fe, fn, _ = runner.unpack_queries(block, seg["pred_embds"].shape[1]); e.append(ev())
        tr = runner.tracker(fe, None, resume=False, frame_embeds_no_norm=fn, with_masks=False); e.append(ev())
        outs = runner.refiner.refine(tr["pred_embds"], fn); e.append(ev())
        masks = runner.refiner.predict_masks(outs, mf[None]); e.append(ev())
        torch.cuda.synchronize()
        wall = (time.perf_counter() - w0) * 1e3
        names = ["pixel_decoder", "predictor", "pack", "tracker", "refiner_layers", "final_masks"]
        res = {n: round(e[i].elapsed_time(e[i + 1]), 3) for i, n in enumerate(names)}
        res["total_ms_device"] = round(e[0].elapsed_time(e[-1]), 3)
        res["wall_ms"] = round(wall, 3)
        res["our_kernel_launches"] = _lib.launch_count - c0
        print(it, res, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "pipeline_breakdown.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
