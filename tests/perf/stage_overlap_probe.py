"""Why is the pipelined step (19.2 ms) 4 ms above stage A alone?  Times, on one B200 at the bench size (T=16, 720p, Q=200):
stage-A graph alone back to back, stage-B graph alone back to back, the pipelined submit loop, and -- inside the pipelined
loop -- the wall duration of each clip's stage B (events on the temporal stream) -- for the library temporal stage and for the
fused-kernel one (fewer launches).  Writes gpurun_out/stage_overlap_probe.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bench import build_models, synthetic_features  # noqa: E402
from dvis_plus_b200.modules.precision import set_precision  # noqa: E402
from dvis_plus_b200.pipeline import GraphedClipRunner  # noqa: E402


def loop_ms(fn, n=12, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


@torch.no_grad()
def measure(fused):
    runner = build_models("cuda", queries=200)
    runner.tracker.use_fused_kernels = fused
    runner.refiner.use_fused_kernels = fused
    feats = {k: v.cuda() for k, v in synthetic_features(16, "swinl").items()}
    g = GraphedClipRunner(runner, feats, depth=3, d2h_stream=True)
    out = {}
    s0 = g.slots[0]

    def a_only():
        with torch.cuda.stream(g.stream_a):
            s0["ga"].replay()
        torch.cuda.current_stream().wait_stream(g.stream_a)

    def b_only():
        with torch.cuda.stream(g.stream_b):
            s0["gb"].replay()
        torch.cuda.current_stream().wait_stream(g.stream_b)

    out["stage_a_graph_alone_ms"] = round(loop_ms(a_only), 3)
    out["stage_b_graph_alone_ms"] = round(loop_ms(b_only), 3)

    def pipelined(runner_g, n=15, warm=6):
        for _ in range(warm):
            runner_g.submit()
        runner_g.wait_all()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            runner_g.submit()
        runner_g.wait_all()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    out["pipelined_ms_per_clip"] = round(pipelined(g), 3)
    g2 = GraphedClipRunner(runner, feats, depth=3, d2h_stream=True, eager_b=True)
    out["pipelined_ms_per_clip_stage_b_issued_eagerly"] = round(pipelined(g2), 3)
    del g2
    if not fused:
        from dvis_plus_b200.partition import sm_partition_streams
        for n_small in (8, 16, 24, 32, 48):
            try:
                sa, sb, info = sm_partition_streams(n_small)
                g3 = GraphedClipRunner(runner, feats, depth=3, d2h_stream=True, stream_a=sa, stream_b=sb)
                out[f"pipelined_ms_per_clip_sm_partition_{info['sms_big']}+{info['sms_small']}"] = round(pipelined(g3), 3)
                print(n_small, info["sms_big"], info["sms_small"], out[f"pipelined_ms_per_clip_sm_partition_{info['sms_big']}+{info['sms_small']}"], flush=True)
                del g3
            except Exception as e:   # noqa: BLE001
                out[f"sm_partition_{n_small}_error"] = repr(e)[:300]
                print("partition", n_small, "failed:", repr(e)[:300], flush=True)
    # stage-B wall duration inside the pipelined loop: events recorded on the temporal stream around the graph replay
    starts, ends = [], []
    orig = [s["gb"] for s in g.slots]

    class Timed:
        def __init__(self, graph):
            self.graph = graph

        def replay(self):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()                       # current stream = stream_b inside submit()
            self.graph.replay()
            e1.record()
            starts.append(e0); ends.append(e1)
    for s in g.slots:
        s["gb"] = Timed(s["gb"])
    for _ in range(12):
        g.submit()
    g.wait_all()
    torch.cuda.synchronize()
    d = [a.elapsed_time(b) for a, b in zip(starts, ends)][3:]
    out["stage_b_wall_ms_inside_pipeline"] = round(sum(d) / len(d), 3)
    for s, o in zip(g.slots, orig):
        s["gb"] = o
    out["captured_launches_per_clip"] = g.captured_launches
    return out


def main():
    set_precision("bf16")
    res = {"library_temporal_stage": measure(False), "fused_temporal_stage": measure(True)}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "stage_overlap_probe.json"), "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
