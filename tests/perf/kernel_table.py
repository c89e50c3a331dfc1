"""Per-kernel time table of one clip step (torch.profiler / CUPTI, in situ -- not ncu): writes
gpurun_out/kernel_table.txt.  Usage: python tests/perf/kernel_table.py [T]"""
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
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    set_precision("bf16")
    runner = build_models("cuda", queries=200)
    feats = {k: v.cuda() for k, v in synthetic_features(T, "swinl").items()}
    for _ in range(3):
        runner(feats)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        runner(feats)
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        if e.device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA:
            rows.append((e.device_time_total, e.count, e.key))
    rows.sort(reverse=True)
    total = sum(r[0] for r in rows)
    lines = [f"total kernel time {total / 1e3:.2f} ms over {sum(r[1] for r in rows)} launches (T={T})"]
    for t, n, k in rows[:60]:
        lines.append(f"{t / 1e3:9.3f} ms {100 * t / total:5.1f}% {n:5d}  {k[:150]}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "kernel_table.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:45]))


if __name__ == "__main__":
    main()
