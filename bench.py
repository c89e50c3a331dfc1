#!/usr/bin/env python
"""Headline benchmark: frames/sec of a T=16, 720p, Q=200 clip through pixel decoder + mask head + tracker + refiner
(BASELINE.json metric / configs[3]: "DVIS++ Swin-L offline"), on N B200s of one box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one clip.  Frames are sharded contiguously across ranks (strong scaling: the clip is fixed), one NCCL
all-gather of the per-frame query block, tracker + refiner replicated, final masks per rank for its own frames.
One JSON line on rank 0.  `--impl reference` times the oracle port of the reference's own CPU path (pure PyTorch
F.grid_sample / einsum / attention, fp32, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BACKBONE_CHANNELS = {  # only input_proj / lateral conv widths depend on the backbone
    "swinl": dict(res2=192, res3=384, res4=768, res5=1536),
    "r50": dict(res2=256, res3=512, res4=1024, res5=2048),
    "vitl": dict(res2=1024, res3=1024, res4=1024, res5=1024),   # ViT-Adapter-L (D/dvis_daq/.../adapter.py:619-624), config 5
}
STRIDES = dict(res2=4, res3=8, res4=16, res5=32)
IMG_H, IMG_W = 736, 1280          # 720p padded to a multiple of 32 (P/dvis_Plus/meta_architecture.py:187,639)
NUM_CLASSES = 25


def synthetic_features(T, backbone="swinl", seed=0, dtype=torch.bfloat16, pin=False, hw=(IMG_H, IMG_W)):
    """Random backbone feature maps for T frames (host tensors)."""
    g = torch.Generator().manual_seed(seed)
    feats = {}
    for k, c in BACKBONE_CHANNELS[backbone].items():
        h, w = hw[0] // STRIDES[k], hw[1] // STRIDES[k]
        # channels_last memory (logical NCHW): what Swin / ViT backbones natively produce (token-major) before the
        # reference permutes them (P/mask2former/modeling/backbone/swin.py); the pixel decoder accepts either layout
        t = torch.empty(T, c, h, w, dtype=dtype).contiguous(memory_format=torch.channels_last)
        for i in range(T):
            t[i] = torch.randn(c, h, w, generator=g).to(dtype)
        feats[k] = t.pin_memory() if pin else t
    return feats


def build_models(device, queries=200, backbone="swinl", seed=0, enc_layers=6, dec_layers=9, trk_layers=6, ref_layers=6):
    """Random-init modules of the DVIS++ offline architecture (P/configs/dvis_Plus/*/*Offline*.yaml shapes).  The
    zero-initialised sampling-offset / attention-weight projections of MSDeformAttn are perturbed so that sampling
    locations vary per query like in a trained model (default init makes every query use the same ring pattern)."""
    from dvis_plus_b200 import modules as M
    from dvis_plus_b200.modules.pixel_decoder import ShapeSpec
    from dvis_plus_b200.pipeline import OfflineClipRunner
    torch.manual_seed(seed)
    ch = BACKBONE_CHANNELS[backbone]
    pd = M.MSDeformAttnPixelDecoder({k: ShapeSpec(channels=ch[k], stride=STRIDES[k]) for k in ch}, transformer_dropout=0.0,
                                    transformer_nheads=8, transformer_dim_feedforward=1024, transformer_enc_layers=enc_layers,
                                    conv_dim=256, mask_dim=256, norm="GN", transformer_in_features=["res3", "res4", "res5"],
                                    common_stride=4)
    for layer in pd.transformer.encoder.layers:
        torch.nn.init.normal_(layer.self_attn.sampling_offsets.weight, std=0.01)
        torch.nn.init.normal_(layer.self_attn.attention_weights.weight, std=0.05)
    dec = M.VideoMultiScaleMaskedTransformerDecoder_dvisPlus(
        256, True, num_classes=NUM_CLASSES, hidden_dim=256, num_queries=queries, nheads=8, dim_feedforward=2048,
        dec_layers=dec_layers, pre_norm=False, mask_dim=256, enforce_input_project=False, num_frames=1,
        num_reid_head_layers=3, reid_hidden_dim=256)
    trk = M.ReferringTracker_noiser(hidden_channel=512, feedforward_channel=2048, num_head=8, decoder_layer_num=trk_layers,
                                    mask_dim=256, class_num=NUM_CLASSES, noise_mode="none")
    rfn = M.TemporalRefiner(hidden_channel=512, feedforward_channel=2048, num_head=8, decoder_layer_num=ref_layers,
                            mask_dim=256, class_num=NUM_CLASSES, windows=16)
    rfn.mask_dtype = torch.bfloat16      # 16-bit mask logits, like the reference's fp16 einsum under eval autocast
    for m in (pd, dec, trk, rfn):
        m.eval().to(device)
    return OfflineClipRunner(pd, dec, trk, rfn)


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's CPU path
# ---------------------------------------------------------------------------------------------------------
def cpu_reference_step(runner_cpu_sd, feats, queries):
    from oracle import torch_port as tp
    pd_sd, dec_sd, trk_sd, rfn_sd = runner_cpu_sd
    mf, _, ms = tp.pixel_decoder_forward_features(pd_sd, feats, num_layers=6)
    seg = tp.predictor_forward(dec_sd, ms, mf, num_layers=9)
    trk = tp.tracker_forward(trk_sd, seg["pred_embds"], None, seg["pred_embds_without_norm"], num_layers=6, with_masks=False)
    out = tp.refiner_forward(rfn_sd, trk["pred_embds"], seg["pred_embds_without_norm"], mf[None], num_layers=6)
    return out["pred_masks"]


def run_cpu_reference(sample_frames, steps, warmup, queries):
    torch.set_num_threads(os.cpu_count())
    runner = build_models("cpu", queries=queries)
    sds = tuple({k: v.detach().float() for k, v in m.state_dict().items()}
                for m in (runner.pixel_decoder, runner.predictor, runner.tracker, runner.refiner))
    feats = {k: v.float() for k, v in synthetic_features(sample_frames, dtype=torch.float32).items()}
    with torch.no_grad():
        for _ in range(warmup):
            cpu_reference_step(sds, feats, queries)
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_reference_step(sds, feats, queries)
        dt = (time.perf_counter() - t0) / steps
    return sample_frames / dt, dt


# ---------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def merge_e2e_legs(e2e, frames, ms_first, ms_overlapped, rel_diff, tol=1e-2):
    """Fold the second end-to-end leg (3 clips in flight, result copies on their own stream) into the `e2e` object: both
    legs' times are always reported; the headline value switches to the second leg only if it is faster AND its results
    agree with the first leg's within `tol` (north_star's bf16 tolerance, relative to the largest magnitude)."""
    e2e["legs_ms_per_step"] = {"2 in flight, copies on the temporal stream": round(ms_first, 3),
                               "3 in flight, copies on their own stream": round(ms_overlapped, 3)}
    e2e["overlapped_leg_rel_max_diff_vs_first_leg"] = round(rel_diff, 6) if rel_diff < float("inf") else None   # no NaN / Infinity in JSON
    if rel_diff <= tol and ms_overlapped < ms_first:
        e2e.update(value=round(frames / ms_overlapped * 1e3, 2), ms_per_step=round(ms_overlapped, 3),
                   pipeline="3 clips in flight, result copies on their own stream")
    return e2e


def merge_round_robin_leg(line, frames, ms_resident, ms_e2e, rel_diff, tol=1e-2):
    """Fold the leg with the temporal stage owned round-robin (N > 1) into the line: its times are always reported; `value`
    and `e2e` switch to it only where it is faster AND its results agree with the replicated leg's within `tol`."""
    line["temporal_stage_legs_ms_per_step"] = {"replicated on every rank": line["ms_per_step"],
                                               "owned round-robin per clip + 1 broadcast": round(ms_resident, 3)}
    line["round_robin_leg_rel_max_diff_vs_replicated"] = round(rel_diff, 6) if rel_diff < float("inf") else None
    line["e2e"].setdefault("legs_ms_per_step", {})["round-robin temporal stage, copies on the masks stream"] = round(ms_e2e, 3)
    if not rel_diff <= tol:
        return line
    if ms_resident < line["ms_per_step"]:
        line.update(value=round(frames / ms_resident * 1e3, 2), ms_per_step=round(ms_resident, 3))
        line["config"]["parallelism"] += "; temporal stage owned round-robin per clip + 1 broadcast"
        line["config"]["execution"] = "3 CUDA graphs per clip (per-frame, temporal on the owner rank, masks), N + 2 clips in flight"
    if ms_e2e < line["e2e"]["ms_per_step"]:
        line["e2e"].update(value=round(frames / ms_e2e * 1e3, 2), ms_per_step=round(ms_e2e, 3),
                           pipeline="temporal stage owned round-robin per clip, N + 2 clips in flight")
    return line


def merge_sm_carveout_leg(line, frames, sms, ms_resident, rel_diff, tol=1e-2):
    """Fold the leg captured with cuBLASLt leaving `sms` SMs to the temporal stage (1 GPU) into the line: always reported,
    adopted as `value` only if faster and in agreement with the first leg."""
    line.setdefault("sm_carveout_legs", []).append(
        {"sms_left_free_by_cublaslt": sms, "ms_per_step": round(ms_resident, 3),
         "rel_max_diff_vs_first_leg": round(rel_diff, 6) if rel_diff < float("inf") else None})
    if rel_diff <= tol and ms_resident < line["ms_per_step"]:
        line.update(value=round(frames / ms_resident * 1e3, 2), ms_per_step=round(ms_resident, 3))
        line["config"]["execution"] = line["config"]["execution"].split("; cuBLASLt kernels leave")[0] + \
            "; cuBLASLt kernels leave %d SMs free for the temporal stage's stream" % sms
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--queries", type=int, default=200)
    ap.add_argument("--cpu-sample-frames", type=int, default=8)   # ~10 s of host work on the GPU box (16 cores)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="no CUDA graphs / no pipelining across clips (debug)")
    ap.add_argument("--d2h-stream", action="store_true",
                    help="end-to-end mode: device->host copies on their own stream (GraphedClipRunner d2h_stream; not yet timed)")
    ap.add_argument("--postprocess", default="none", choices=["none", "vis"],
                    help="vis: end the clip with the fused video-instance post-processing (top-10 instances selected before "
                         "the final mask GEMM, 720p bit-packed masks) instead of all Q stride-4 mask logits -- a different, "
                         "smaller result; not the default metric's workload and not yet timed on a B200")
    ap.add_argument("--temporal", default="replicated", choices=["replicated", "round_robin"],
                    help="N > 1: tracker + refiner replicated on every rank (default, measured) or owned round-robin per clip "
                         "with one broadcast (pipeline.RoundRobinClipRunner; not yet timed on a multi-GPU box)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    T, Q = args.frames, args.queries
    config = {"workload": "DVIS++ Swin-L offline: T=16 clip 720p (736x1280 padded), Q=200, pixel decoder (6 MSDeformAttn "
                          "encoder layers + FPN) + mask head (9-layer masked-attention predictor, 10 mask GEMMs/frame) + "
                          "ReferringTracker (6 layers) + TemporalRefiner (6 layers) + final mask GEMM; backbone features synthetic",
              "frames": T, "queries": Q, "backbone_channels": "swinl",
              "parallelism": f"frames sharded {world}x{T // max(world, 1)} + 1 NCCL all-gather of frame queries" if world > 1 else "1 GPU",
              "l2": "inputs larger than L2 (0.68 GB of bf16 backbone features per step >> 126 MB)",
              "execution": "2 CUDA graphs per clip (per-frame stage, temporal stage), software-pipelined across consecutive "
                           "clips on 2 streams (depth 2); latency_ms_per_clip is one clip run alone, eagerly"}

    if args.impl == "reference":
        if rank != 0:
            return
        sample = max(4, args.cpu_sample_frames)     # 4 frames per step: ~5-20 s of host work per step
        fps, dt = run_cpu_reference(sample, max(1, min(args.steps, 3)), min(args.warmup, 1), Q)
        line = {"impl": "reference", "metric": "frames/sec (720p, T=16, Q=200) pixel-decoder+mask+refiner", "value": round(fps, 4),
                "unit": "frames/s", "n_gpus": args.gpus, "steps": max(1, min(args.steps, 3)), "warmup": min(args.warmup, 1),
                "ms_per_step": round(dt * 1e3, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": round(fps, 4), "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                 "sample": f"{sample} frame(s) of the 720p Q={Q} workload per step through the oracle port "
                                           "(oracle/torch_port.py: reference's pure-PyTorch CPU path), fp32, all host threads"},
                "e2e": {"value": round(fps, 4), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    import torch.distributed as dist
    from dvis_plus_b200 import _lib
    from dvis_plus_b200.modules.precision import set_precision
    assert torch.cuda.is_available(), "bench.py --impl ours needs a CUDA device (there is no CPU fallback)"
    assert T % world == 0, "frames must divide evenly across ranks"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    set_precision("bf16")
    _lib.lib()
    runner = build_models(dev, queries=Q)
    t_local = T // world
    host = synthetic_features(T, pin=False)
    host = {k: v[rank * t_local:(rank + 1) * t_local].contiguous(memory_format=torch.channels_last).pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from dvis_plus_b200.pipeline import GraphedClipRunner

    vis = None
    if args.postprocess == "vis":
        from dvis_plus_b200.modules.postprocess import VideoPostProcessor
        vis = dict(post=VideoPostProcessor(NUM_CLASSES, num_queries=Q, max_num=10), img_size=(720, IMG_W), output_size=(720, IMG_W),
                   packed=True)
        config["workload"] += " + fused VIS post-processing (10 instances selected before the final mask GEMM, 720p bit-packed masks)"

    def step_eager():
        if vis is None:
            return runner(resident)
        blk, mf = runner.segment_stage(resident)
        return runner.vis_from_block(runner.gather_queries(blk), mf, (blk.shape[-1] - (NUM_CLASSES + 1)) // 2, **vis)

    out0 = step_eager()
    d2h_keys = ("pred_masks", "pred_logits") if vis is None else ("pred_masks", "pred_scores", "pred_labels", "pred_ids")
    d2h = {k: torch.empty(out0[k].shape, dtype=out0[k].dtype).pin_memory() for k in d2h_keys}
    d2h_bytes = sum(v.numel() * v.element_size() for v in d2h.values())

    if args.eager:
        graphed = None
    elif args.temporal == "round_robin":
        from dvis_plus_b200.pipeline import RoundRobinClipRunner
        graphed = RoundRobinClipRunner(runner, resident, vis=vis)
        config["parallelism"] += "; temporal stage owned round-robin per clip + 1 broadcast"
        config["execution"] = "3 CUDA graphs per clip (per-frame, temporal on the owner rank, masks), %d clips in flight" % graphed.depth
    else:
        graphed = GraphedClipRunner(runner, resident, depth=2, vis=vis, d2h_stream=args.d2h_stream)

    def run_steps(n, mode):
        """n clips back to back; every clip's results are complete when this returns (after the closing barrier)."""
        if graphed is None:
            for _ in range(n):
                if mode == "e2e":
                    feats = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
                    if vis is None:
                        out = runner(feats)
                    else:
                        blk, mf = runner.segment_stage(feats)
                        out = runner.vis_from_block(runner.gather_queries(blk), mf, (blk.shape[-1] - (NUM_CLASSES + 1)) // 2, **vis)
                    for k, v in d2h.items():
                        v.copy_(out[k], non_blocking=True)
                else:
                    step_eager()
        else:
            for _ in range(n):
                graphed.submit(host if mode == "e2e" else None, d2h if mode == "e2e" else None)
            graphed.wait_all()

    def timed(mode, steps, warmup):
        run_steps(warmup, mode)
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        run_steps(steps, mode)
        e.record()
        barrier()
        t = torch.tensor([s.elapsed_time(e) / steps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev = timed("resident", args.steps, max(3, args.warmup))
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed("e2e", args.steps, 2)

    # single-clip latency and per-kernel CUDA-event timing: an instrumented EAGER pass of the same clip
    # (kernels inside CUDA graphs cannot be bracketed from the host)
    for _ in range(2):
        step_eager()
    barrier()
    _lib.start_timing()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    n_lat = min(args.steps, 5)
    for _ in range(n_lat):
        step_eager()
    e.record()
    barrier()
    kern = _lib.stop_timing()
    ms_latency = s.elapsed_time(e) / n_lat
    launches = (graphed.captured_launches if graphed is not None else sum(v[0] for v in kern.values()) // n_lat) * args.steps

    def finish():
        """Leave without tearing NCCL down rank by rank (ProcessGroupNCCL teardown can wait on peers that already left):
        one last barrier so rank 0 has printed, then a hard exit on every rank."""
        sys.stdout.flush()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
            os._exit(0)

    # Two more configurations of the SAME workload, timed after everything above on every rank (their collectives need all
    # of them).  Both are new on a device this round, so each runs under a watchdog, its results are compared with the first
    # leg's, and the line's headline numbers switch to it only if it is faster and agrees (merge_e2e_legs /
    # merge_round_robin_leg); whatever happens, the numbers measured above are printed.
    extra_legs = graphed is not None and not args.d2h_stream and args.temporal == "replicated"
    d2h_ref = {k: v.clone() for k, v in d2h.items()} if extra_legs else None

    def rel_diff_to_first_leg():
        """same inputs -> same results, up to the GEMM algorithms a second capture may pick: relative max difference"""
        torch.cuda.synchronize()
        diff = max(float((d2h[k].float() - d2h_ref[k].float()).abs().max() / d2h_ref[k].float().abs().max().clamp_min(1e-6))
                   for k in d2h)
        t = torch.tensor([diff if diff == diff else float("inf")], device=dev)     # NaN anywhere counts as a mismatch
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)                               # every rank's frames must agree
        return float(t.item())

    def e2e_overlapped_leg():
        """-> (ms per clip, relative max difference of its results to the first end-to-end leg's) or None.  In the first
        leg a clip's 377 MB device->host copy sits on the temporal stage's stream and its slot is one of two, so nothing else
        runs while it drains (27.1 vs 20.2 ms per clip in round 1's run).  Same public API (GraphedClipRunner.submit with
        host buffers), same per-step copies inside the timed region."""
        nonlocal graphed
        first = graphed
        try:
            graphed = GraphedClipRunner(runner, resident, depth=3, vis=vis, d2h_stream=True)
            ms = timed("e2e", args.steps, 3)
            return ms, rel_diff_to_first_leg()
        except Exception as exc:                                  # keep the measured first leg; say why
            sys.stderr.write("bench: overlapped end-to-end leg failed: %r\n" % (exc,))
            return None
        finally:
            graphed = first

    def round_robin_leg():
        """N > 1 -> (resident ms per clip, end-to-end ms per clip, relative max difference to the first leg) or None: the
        temporal stage owned round-robin per clip + one broadcast (pipeline.RoundRobinClipRunner) instead of replicated on
        every rank, which is the Amdahl term of the strong scaling."""
        nonlocal graphed
        first = graphed
        try:
            from dvis_plus_b200.pipeline import RoundRobinClipRunner
            graphed = RoundRobinClipRunner(runner, resident, vis=vis)
            ms_res = timed("resident", args.steps, max(3, args.warmup, graphed.depth))   # every slot used once before timing
            ms_e2e_rr = timed("e2e", args.steps, 3)
            return ms_res, ms_e2e_rr, rel_diff_to_first_leg()
        except Exception as exc:
            sys.stderr.write("bench: round-robin leg failed: %r\n" % (exc,))
            return None
        finally:
            graphed = first

    def sm_carveout_leg(sms=8):
        """1 GPU -> (resident ms per clip, relative max difference to the first leg) or None.  The temporal stage is a chain
        of ~900 tiny dependent kernels on a high-priority stream; while a persistent cuBLASLt kernel of the next clip's
        per-frame stage owns every SM, the chain's next kernel has nowhere to run (pipelined 20.2 ms per clip against
        16.2 ms of per-frame work in round 1's run).  Here the graphs are captured with cuBLASLt told to leave `sms` SMs
        free (torch's SM carve-out -> CUBLASLT_MATMUL_DESC_SM_COUNT_TARGET)."""
        nonlocal graphed
        first = graphed
        prev = torch._C._get_sm_carveout_experimental()
        try:
            torch._C._set_sm_carveout_experimental(sms)
            graphed = GraphedClipRunner(runner, resident, depth=2, vis=vis)
            ms = timed("resident", args.steps, 3)
            graphed.submit(None, d2h)                             # one clip's results for the comparison
            graphed.wait_all()
            return ms, rel_diff_to_first_leg()
        except Exception as exc:
            sys.stderr.write("bench: SM carve-out leg failed: %r\n" % (exc,))
            return None
        finally:
            torch._C._set_sm_carveout_experimental(prev)
            graphed = first

    if rank != 0:
        if extra_legs:
            watchdog = threading.Timer(150.0, lambda: os._exit(0))   # stays armed: rank 0 may leave without the barrier
            watchdog.daemon = True
            watchdog.start()
            if e2e_overlapped_leg() is None or (world > 1 and round_robin_leg() is None):
                os._exit(0)                                       # rank 0 prints what was measured before; no teardown
        finish()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
    # dominant kernel of this repo on the path: the fused MSDA forward (6 launches / step / rank)
    S, M_, D_, L_, P_ = 19320, 8, 32, 3, 4
    msda_bytes = t_local * (S * M_ * D_ * 2 + S * M_ * L_ * P_ * 3 * 2 + S * L_ * 2 * 4 + S * M_ * D_ * 2)
    roof = None
    if kern and "dvis_msda_fused_forward" in kern:
        n, tot = kern["dvis_msda_fused_forward"]
        us = tot / n * 1e3
        ach = msda_bytes / us / 1e3
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from profiles/r1_ncu_msda_bf16_fused_and_pair_N8.txt
        # (ncu --set full, 8 frames per launch: 173.1 + 62.3 MB) scaled to the frames of one launch here
        traffic = int((173.107712e6 + 62.283008e6) / 8 * t_local)
        roof = {"kernel": "msda_fwd_staged_kernel (dvis_msda_fused_forward)", "bound": "hbm", "achieved": round(ach, 1),
                "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": traffic, "us_per_launch": round(us, 1),
                "algorithmic_bytes_per_launch": msda_bytes, "peak_source": peak_src,
                "share_of_clip_latency": round(tot / (ms_latency * n_lat), 4),
                "measured": "CUDA events around each launch in an eager instrumented pass of the same clip",
                "our_kernels_ms_per_clip": {k: round(v[1] / n_lat, 3) for k, v in kern.items()}}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        fps, dt = run_cpu_reference(args.cpu_sample_frames, 1, 0, Q)
        cpu = {"value": round(fps, 4), "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{args.cpu_sample_frames} frame(s) of the same 720p Q={Q} workload, one pass of the oracle port "
                         f"(reference's pure-PyTorch CPU path), fp32, all host threads, {dt:.1f} s"}
    line = {"metric": "frames/sec (720p, T=16, Q=200) pixel-decoder+mask+refiner", "value": round(T / ms_dev * 1e3, 2),
            "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": round(ms_dev, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic (random backbone features, random-init weights, perturbed MSDeformAttn offsets)",
            "config": config, "clocks": clocks, "latency_ms_per_clip": round(ms_latency, 3),
            "e2e": {"value": round(T / ms_e2e * 1e3, 2), "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes * world,
                    "d2h_bytes_per_step": d2h_bytes * world, "ms_per_step": round(ms_e2e, 3)},
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu}
    line["e2e"]["pipeline"] = ("eager, one clip at a time" if graphed is None else
                               "result copies on their own stream" if args.d2h_stream else
                               "2 clips in flight, result copies on the temporal stage's stream")
    if extra_legs:
        # The line above is complete and measured: if an extra leg hangs or fails, print it as it stands and leave.
        def leave(why):
            line["extra_legs"] = why
            print(json.dumps(line), flush=True)
            sys.stdout.flush()
            os._exit(0)                                           # the CUDA context / NCCL may be unusable: no teardown
        watchdog = threading.Timer(150.0, leave, args=("timed out",))
        watchdog.daemon = True
        watchdog.start()
        res = e2e_overlapped_leg()
        if res is None:
            leave("overlapped end-to-end leg failed (see stderr)")
        merge_e2e_legs(line["e2e"], T, ms_e2e, *res)
        if world > 1:
            res = round_robin_leg()
            if res is None:
                leave("round-robin leg failed (see stderr)")
            merge_round_robin_leg(line, T, *res)
        else:
            for sms in (8, 16):
                res = sm_carveout_leg(sms)
                if res is None:
                    leave("SM carve-out leg failed (see stderr)")
                merge_sm_carveout_leg(line, T, sms, *res)
        watchdog.cancel()
        line["extra_legs"] = "completed"
    print(json.dumps(line), flush=True)
    finish()


if __name__ == "__main__":
    main()
