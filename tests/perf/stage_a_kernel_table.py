"""Kernel table of stage A alone (pixel decoder + predictor) for a given number of frames: which launches make up its
launch-bound floor.  Usage: python tests/perf/stage_a_kernel_table.py [frames]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bench import build_models, synthetic_features  # noqa: E402
from dvis_plus_b200.modules.precision import set_precision  # noqa: E402


@torch.no_grad()
def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    set_precision("bf16")
    runner = build_models("cuda", queries=200)
    feats = {k: v.cuda() for k, v in synthetic_features(frames, "swinl").items()}
    for _ in range(3):
        runner.segment_stage(feats)
    torch.cuda.synchronize()
    for name, fn in (("pixel_decoder", lambda: runner.pixel_decoder.forward_features(feats)), ("stage_a", lambda: runner.segment_stage(feats))):
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            fn()
            torch.cuda.synchronize()
        rows = [(e.count, e.device_time_total, e.key) for e in prof.key_averages()
                if e.device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA]
        rows.sort(reverse=True)
        print(f"== {name}: {sum(r[0] for r in rows)} launches, {sum(r[1] for r in rows) / 1e3:.2f} ms kernel time ({frames} frames)")
        for n, t, k in rows[:40]:
            print(f"{n:5d} {t / 1e3:8.3f} ms  {k[:120]}")


if __name__ == "__main__":
    main()
