"""Root-cause tool for VERDICT r1 "What's weak" #1: BENCH_r01 showed every re-captured runner differing from the FIRST
GraphedClipRunner's results by 1.5 % of the output scale on identical inputs.  This script separates the candidates:

  1. eager vs eager (run-to-run determinism of the kernels themselves)
  2. capture #1 vs eager, capture #2 vs eager, capture #1 vs capture #2 -- per stage (mask features, packed query block,
     final logits / masks), resident inputs
  3. replay-to-replay determinism of ONE capture, alone and with two clips in flight (overlap of stage A / stage B)
  4. the same in fp32 mode (are ulp-level GEMM differences amplified by thresholds / Hungarian decisions?)

    python tests/perf/diagnose_legs.py [--frames 16] > gpurun_out/diagnose_legs.json
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))


def frac_diff(a, b):
    return float((a != b).float().mean())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--queries", type=int, default=200)
    args = ap.parse_args()
    import bench
    from dvis_plus_b200.modules.precision import set_precision
    from dvis_plus_b200.pipeline import GraphedClipRunner
    dev = torch.device("cuda", 0)
    report = {}
    for prec in ("bf16", "fp32"):
        set_precision(prec)
        runner = bench.build_models(dev, queries=args.queries)
        host = bench.synthetic_features(args.frames, pin=False)
        host = {k: v.contiguous(memory_format=torch.channels_last).pin_memory() for k, v in host.items()}
        resident = {k: v.to(dev) for k, v in host.items()}
        C = 512
        r = {}

        def eager():
            blk, mf = runner.segment_stage(resident)
            out = runner.temporal_from_block(runner.gather_queries(blk), mf, C)
            return dict(block=blk.clone(), mf=mf.clone(), logits=out["pred_logits"].clone(), masks=out["pred_masks"].clone(),
                        embds=out["pred_embds"].clone())

        def cmp(a, b):
            return {k: rel(a[k], b[k]) for k in a}

        e1 = eager()
        e2 = eager()
        r["eager_vs_eager"] = cmp(e1, e2)
        # the tracker's eager path replays its own per-frame CUDA graph; compare with the plain loop
        runner.tracker.use_cuda_graph = False
        e3 = eager()
        runner.tracker.use_cuda_graph = True
        r["eager_trackergraph_vs_plain"] = cmp(e1, e3)

        def snap(slot):
            torch.cuda.synchronize()
            o = slot["out"]
            return dict(block=slot["block"].clone(), mf=slot["mf"].clone(), logits=o["pred_logits"].clone(),
                        masks=o["pred_masks"].clone(), embds=o["pred_embds"].clone())

        caps = []
        for i in range(2):
            g = GraphedClipRunner(runner, resident, depth=2)
            s0 = g.submit(None, None)
            g.wait_all()
            a = snap(s0)
            s1 = g.submit(None, None)
            g.wait_all()
            b = snap(s1)
            r[f"capture{i + 1}_slot0_vs_eager"] = cmp(a, e1)
            r[f"capture{i + 1}_slot1_vs_slot0"] = cmp(b, a)
            # replay determinism, one clip at a time
            s0 = g.submit(None, None)
            g.wait_all()
            r[f"capture{i + 1}_slot0_replay_vs_first"] = cmp(snap(s0), a)
            # two clips in flight (stage B of clip i overlaps stage A of clip i+1), 6 clips
            for _ in range(6):
                last = g.submit(None, None)
            g.wait_all()
            r[f"capture{i + 1}_pipelined_vs_first"] = cmp(snap(last), a)
            # end to end: host inputs, results to pinned host buffers
            d2h = {k: torch.empty(last["out"][k].shape, dtype=last["out"][k].dtype).pin_memory() for k in ("pred_masks", "pred_logits")}
            for _ in range(4):
                last = g.submit(host, d2h)
            g.wait_all()
            torch.cuda.synchronize()
            r[f"capture{i + 1}_e2e_vs_first"] = {"masks": rel(d2h["pred_masks"].to(dev), a["masks"]),
                                                 "logits": rel(d2h["pred_logits"].to(dev), a["logits"])}
            caps.append(a)
            del g
        r["capture2_vs_capture1"] = cmp(caps[1], caps[0])
        g3 = GraphedClipRunner(runner, resident, depth=3, d2h_stream=True)
        d2h = {k: torch.empty(caps[0][n].shape, dtype=caps[0][n].dtype).pin_memory() for k, n in (("pred_masks", "masks"), ("pred_logits", "logits"))}
        for _ in range(6):
            g3.submit(host, d2h)
        g3.wait_all()
        torch.cuda.synchronize()
        r["depth3_d2hstream_e2e_vs_capture1"] = {"masks": rel(d2h["pred_masks"].to(dev), caps[0]["masks"]),
                                                  "logits": rel(d2h["pred_logits"].to(dev), caps[0]["logits"])}
        report[prec] = r
        del g3, runner
        torch.cuda.empty_cache()
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
