"""Per-kernel GPU durations (CUPTI via torch.profiler) of one tracker window, fused vs library path: T=16, Q=200, hidden 512.
Kernel-level times are independent of the host launch overhead of the eager loop."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dvis_plus_b200.modules.precision import set_precision  # noqa: E402


@torch.no_grad()
def main():
    set_precision("bf16")
    dev = "cuda"
    runner = bench.build_models(dev, queries=200)
    T, Q = 16, 200
    base = torch.randn(1, 512, 1, Q, device=dev)
    fe = base + 0.3 * torch.randn(1, 512, T, Q, device=dev)
    fn = fe + 0.1 * torch.randn(1, 512, T, Q, device=dev)
    trk = runner.tracker
    trk.use_cuda_graph = False
    for fused in (True, False):
        trk.use_fused_kernels = fused
        for _ in range(3):
            trk(fe, None, resume=False, frame_embeds_no_norm=fn, with_masks=False)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            trk(fe, None, resume=False, frame_embeds_no_norm=fn, with_masks=False)
            torch.cuda.synchronize()
        rows = [(e.device_time_total, e.count, e.key) for e in prof.key_averages() if e.device_time_total > 0]
        rows.sort(reverse=True)
        total = sum(r[0] for r in rows)
        print("=== %s: %.2f ms of kernel time over %d launches" % ("fused" if fused else "library", total / 1e3, sum(r[1] for r in rows)))
        for t, n, k in rows[:14]:
            print("%9.1f us  %5d x %6.2f us  %s" % (t, n, t / n, k[:110]))


if __name__ == "__main__":
    main()
