"""Which torch (aten) ops of one clip step still cost device time, with shapes and the Python line that issued them
(torch.profiler, record_shapes + with_stack).  Writes gpurun_out/aten_op_table.txt."""
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
    set_precision("bf16")
    runner = build_models("cuda", queries=200)
    feats = {k: v.cuda() for k, v in synthetic_features(16, "swinl").items()}
    for _ in range(3):
        runner(feats)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=True) as prof:
        runner(feats)
        torch.cuda.synchronize()
    rows = []
    for e in prof.events():                       # CPU-side ops with the kernels they launched
        ks = getattr(e, "kernels", None) or []
        for k in ks:
            if k.duration > 8 and any(t in k.name for t in ("elementwise", "CatArray", "Memcpy", "nchwToNhwc", "nhwcToNchw", "copy")):
                rows.append((k.duration, e.name, str(e.input_shapes)[:110], k.name[:60]))
    rows.sort(reverse=True)
    lines = [f"{t:8.1f} us  {n[:26]:26s} {sh:110s} {kn}" for t, n, sh, kn in rows[:70]]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "aten_op_table.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:45]))


if __name__ == "__main__":
    main()
