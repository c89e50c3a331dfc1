"""Stage A (pixel decoder + predictor) of ONE rank at N GPUs = T/N frames, as a CUDA graph replayed back to back on one B200:
the floor of the per-clip step at N GPUs, next to its kernel-launch count (tests the 'launch-bound at 2 frames per rank' claim of
DESIGN section 6).  Writes gpurun_out/per_rank_stage_probe.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bench import build_models, synthetic_features  # noqa: E402
from dvis_plus_b200.modules.precision import set_precision  # noqa: E402


@torch.no_grad()
def main():
    set_precision("bf16")
    runner = build_models("cuda", queries=200)
    out = {}
    for frames in (16, 8, 4, 2, 1):
        feats = {k: v.cuda() for k, v in synthetic_features(frames, "swinl").items()}
        for _ in range(3):
            runner.segment_stage(feats)
        torch.cuda.synchronize()
        st = torch.cuda.Stream()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            runner.segment_stage(feats)
        torch.cuda.synchronize()
        nodes = None
        try:
            nodes = len(g.raw_cuda_graph().get_nodes()) if hasattr(g, "raw_cuda_graph") else None
        except Exception:   # noqa: BLE001
            pass
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        out[f"frames_{frames}"] = {"stage_a_graph_ms": round(e0.elapsed_time(e1) / 20, 3), "ms_per_frame": round(e0.elapsed_time(e1) / 20 / frames, 3),
                                   "graph_nodes": nodes}
        print(frames, out[f"frames_{frames}"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "per_rank_stage_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
