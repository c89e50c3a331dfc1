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
METRIC = "frames/sec (720p, T=16, Q=200) pixel-decoder+mask+refiner"


def rel_max_diff(a, b):
    """max |a - b| / max |b| over a dict of tensors (NaN anywhere -> inf)."""
    worst = 0.0
    for k in b:
        x, y = a[k].float(), b[k].float().to(a[k].device)
        d = float((x.detach() - y.detach()).abs().max() / y.detach().abs().max().clamp_min(1e-6))
        worst = max(worst, d if d == d else float("inf"))
    return worst


class GraphedStep:
    """One whole step (configs 2 / 3: a single frame / a 5-frame online clip) captured into ONE CUDA graph with static input
    buffers; same submit / wait_all surface as the clip runners."""
    depth = 1

    def __init__(self, fn, example):
        self.inp = {k: v.clone() for k, v in example.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):
                fn(self.inp)
        torch.cuda.current_stream().wait_stream(side)
        from dvis_plus_b200 import _lib
        n0 = _lib.launch_count
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.out = fn(self.inp)
        self.captured_launches = _lib.launch_count - n0

    def submit(self, features=None, d2h=None):
        if features is not None:
            for k, v in features.items():
                self.inp[k].copy_(v, non_blocking=True)
        self.graph.replay()
        if d2h is not None:
            for k, v in d2h.items():
                v.copy_(self.out[k], non_blocking=True)

    def wait_all(self):
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=[2, 3, 4],
                    help="BASELINE.json config: 4 = the headline metric's workload (default); 2 = R50 pixel decoder + predictor, "
                         "single 720p frame, Q=100; 3 = R50 online clip T=5, Q=200 with the tracker (1 GPU each)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="fp32: every GEMM / conv in fp32 like the reference's forced-fp32 pixel decoder (msdeformattn.py:314)")
    ap.add_argument("--frames", type=int, default=None)
    ap.add_argument("--queries", type=int, default=None)
    ap.add_argument("--cpu-sample-frames", type=int, default=8)   # ~10 s of host work on the GPU box (16 cores)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="no CUDA graphs / no pipelining across clips (debug)")
    ap.add_argument("--postprocess", default="none", choices=["none", "vis"],
                    help="vis: end the clip with the fused video-instance post-processing (top-10 instances selected before "
                         "the final mask GEMM, 720p bit-packed masks) instead of all Q stride-4 mask logits -- a different, "
                         "smaller result, not the default metric's workload")
    ap.add_argument("--temporal", default="auto", choices=["auto", "replicated", "round_robin"],
                    help="N > 1: tracker + refiner replicated on every rank, or owned round-robin per clip with one broadcast "
                         "(pipeline.RoundRobinClipRunner); auto = round_robin when N > 1")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = {4: dict(T=16, Q=200, backbone="swinl"), 3: dict(T=5, Q=200, backbone="r50"), 2: dict(T=1, Q=100, backbone="r50")}[args.config]
    T = args.frames or cfg["T"]
    Q = args.queries or cfg["Q"]
    backbone = cfg["backbone"]
    workloads = {
        4: "DVIS++ Swin-L offline: T=16 clip 720p (736x1280 padded), Q=200, pixel decoder (6 MSDeformAttn encoder layers + FPN) + "
           "mask head (9-layer masked-attention predictor, 10 mask GEMMs/frame) + ReferringTracker (6 layers) + TemporalRefiner "
           "(6 layers) + final mask GEMM; backbone features synthetic",
        3: "DVIS++ R50 online: T=5 clip 720p, Q=200: pixel decoder + mask head + ReferringTracker with its mask GEMM "
           "(BASELINE config 3); backbone features synthetic",
        2: "R50 MSDeformAttnPixelDecoder, 4 scales, 720p single frame + Q=100 predictor (10 mask-head calls) (BASELINE config 2); "
           "backbone features synthetic"}
    config = {"workload": workloads[args.config], "baseline_config": args.config, "frames": T, "queries": Q,
              "backbone_channels": backbone, "precision": args.precision,
              "parallelism": f"frames sharded {world}x{T // max(world, 1)} + 1 NCCL all-gather of frame queries" if world > 1 else "1 GPU",
              "l2": "inputs larger than L2 (%.2f GB of backbone features per step >> 126 MB)" % (0.0424 * T) if T >= 4 else
                    "L2 flushed between steps (256 MB write)"}

    if args.impl == "reference":
        if rank != 0:
            return
        sample = max(4, args.cpu_sample_frames)     # 4 frames per step: ~5-20 s of host work per step
        fps, dt = run_cpu_reference(sample, max(1, min(args.steps, 3)), min(args.warmup, 1), Q)
        line = {"impl": "reference", "metric": METRIC, "value": round(fps, 4),
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
    assert args.config == 4 or world == 1, "configs 2 and 3 are single-GPU workloads"
    assert T % world == 0, "frames must divide evenly across ranks"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    set_precision(args.precision)
    _lib.lib()
    runner = build_models(dev, queries=Q, backbone=backbone)
    if os.environ.get("DVIS_BENCH_TRACKER_LIBRARY_ATTENTION"):       # experiment switch (profiles/r2_stage_overlap_probe.json)
        runner.tracker.use_custom_attention = False
    t_local = T // world
    host = synthetic_features(T, backbone, pin=False)
    host = {k: v[rank * t_local:(rank + 1) * t_local].contiguous(memory_format=torch.channels_last).pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from dvis_plus_b200.pipeline import GraphedClipRunner, OnlineClipRunner, RoundRobinClipRunner

    vis = None
    if args.postprocess == "vis":
        assert args.config == 4
        from dvis_plus_b200.modules.postprocess import VideoPostProcessor
        vis = dict(post=VideoPostProcessor(NUM_CLASSES, num_queries=Q, max_num=10), img_size=(720, IMG_W), output_size=(720, IMG_W),
                   packed=True)
        config["workload"] += " + fused VIS post-processing (10 instances selected before the final mask GEMM, 720p bit-packed masks)"

    online = OnlineClipRunner(runner.pixel_decoder, runner.predictor, runner.tracker, window_size=T) if args.config == 3 else None

    @torch.no_grad()
    def step_eager(feats=None):
        """one clip through the public runner API, eagerly; -> dict of device tensors"""
        feats = resident if feats is None else feats
        if args.config == 2:
            mf, _, ms = runner.pixel_decoder.forward_features(feats)
            seg = runner.predictor(ms, mf)
            return {"pred_masks": seg["pred_masks"], "pred_logits": seg["pred_logits"]}
        if args.config == 3:
            out = online(feats)
            return {"pred_masks": out["pred_masks"], "pred_logits": out["pred_logits"]}
        if vis is None:
            return runner(feats)
        blk, mf = runner.segment_stage(feats)
        return runner.vis_from_block(runner.gather_queries(blk), mf, (blk.shape[-1] - (NUM_CLASSES + 1)) // 2, **vis)

    out0 = step_eager()
    d2h_keys = ("pred_masks", "pred_logits") if vis is None else ("pred_masks", "pred_scores", "pred_labels", "pred_ids")
    d2h = {k: torch.empty(out0[k].shape, dtype=out0[k].dtype).pin_memory() for k in d2h_keys}
    d2h_bytes = sum(v.numel() * v.element_size() for v in d2h.values())
    eager_ref = {k: out0[k].clone() for k in d2h_keys}
    del out0

    temporal = args.temporal if args.temporal != "auto" else ("round_robin" if world > 1 else "replicated")
    graphed = None
    if args.config == 4 and not args.eager:
        if temporal == "round_robin" and world > 1:
            graphed = RoundRobinClipRunner(runner, resident, vis=vis)
            config["parallelism"] += "; temporal stage owned round-robin per clip + 1 broadcast"
            config["execution"] = ("3 CUDA graphs per clip (per-frame stage, temporal stage on the owner rank, masks), %d clips "
                                   "in flight" % graphed.depth)
        else:
            graphed = GraphedClipRunner(runner, resident, depth=3, vis=vis, d2h_stream=True)
            config["execution"] = ("2 CUDA graphs per clip (per-frame stage, temporal stage), software-pipelined across 3 clips "
                                   "in flight on 2 streams; host<->device copies on streams of their own")
    elif not args.eager:
        graphed = GraphedStep(step_eager, resident)
        config["execution"] = "one CUDA graph per step, one step at a time; host<->device copies on the same stream"
    else:
        config["execution"] = "eager, one clip at a time (the tracker replays a CUDA graph per frame)"
    if T < 4:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def run_steps(n, mode):
        """n clips back to back; every clip's results are complete when this returns (after the closing barrier)."""
        if graphed is None:
            for _ in range(n):
                if T < 4:
                    flush.zero_()                                   # inputs smaller than L2: flush it between steps
                if mode == "e2e":
                    out = step_eager({k: v.to(dev, non_blocking=True) for k, v in host.items()})
                    for k, v in d2h.items():
                        v.copy_(out[k], non_blocking=True)
                else:
                    step_eager()
        else:
            for _ in range(n):
                if T < 4:
                    flush.zero_()                                   # inputs smaller than L2: flush it between steps
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

    warm = max(3, args.warmup, graphed.depth if graphed is not None else 0)     # every slot used once before timing
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev = timed("resident", args.steps, warm)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed("e2e", args.steps, warm)

    # parity of the TIMED path: the result the end-to-end leg just delivered to the host buffers against the eager
    # runner's result for the same clip (every rank's frames; bit-identical since round 2, profiles/r2_determinism.md)
    torch.cuda.synchronize()
    diff = torch.tensor([rel_max_diff({k: d2h[k].to(dev) for k in d2h}, eager_ref)], device=dev)
    if world > 1:
        dist.all_reduce(diff, op=dist.ReduceOp.MAX)
    parity = {"timed_e2e_output_vs_eager_runner_rel_max_diff": float(diff.item()), "bit_identical": float(diff.item()) == 0.0,
              "what": "pred_masks + pred_logits delivered to the host by the last timed end-to-end step vs OfflineClipRunner run "
                      "eagerly on the same clip, max over ranks; the eager runner vs the oracle port is tests/test_configs_gpu.py"}

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
    eager_launches = sum(v[0] for v in kern.values()) // n_lat
    launches = (graphed.captured_launches if graphed is not None else eager_launches) * args.steps

    if rank != 0:
        if world > 1:                                               # leave together: rank 0 prints first
            dist.barrier()
            torch.cuda.synchronize()
            os._exit(0)
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
    # dominant kernel of this repo on the path: the fused MSDA forward (6 launches / step / rank)
    S, M_, D_, L_, P_ = 19320, 8, 32, 3, 4
    esz = 2 if args.precision == "bf16" else 4
    msda_bytes = t_local * (S * M_ * D_ * esz + S * M_ * L_ * P_ * 3 * esz + S * L_ * 2 * 4 + S * M_ * D_ * esz)
    roof = None
    msda_entry = next((k for k in ("dvis_msda_fused_forward_hm", "dvis_msda_fused_forward") if kern and k in kern), None)
    if msda_entry:
        n, tot = kern[msda_entry]
        us = tot / n * 1e3
        ach = msda_bytes / us / 1e3
        roof = {"kernel": "msda_fwd_staged_kernel (%s)" % msda_entry, "bound": "hbm", "achieved": round(ach, 1),
                "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                # not measurable inside this run (needs ncu): imported from the committed ncu --set full capture, per launch,
                # scaled by frames -- null for shapes that capture does not cover
                "traffic": int(MSDA_NCU_DRAM_BYTES_PER_FRAME * t_local) if args.precision == "bf16" else None,
                "traffic_source": MSDA_NCU_SOURCE if args.precision == "bf16" else None,
                "us_per_launch": round(us, 1), "algorithmic_bytes_per_launch": msda_bytes, "peak_source": peak_src,
                "share_of_clip_latency": round(tot / (ms_latency * n_lat), 4),
                "measured": "CUDA events around each launch in an eager instrumented pass of the same clip",
                "our_kernels_ms_per_clip": {k: round(v[1] / n_lat, 3) for k, v in kern.items()},
                "our_kernel_launches_per_clip": {k: v[0] // n_lat for k, v in kern.items()}}
    cpu = None
    if world == 1 and not args.no_cpu_baseline and args.config == 4:
        fps, dt = run_cpu_reference(args.cpu_sample_frames, 1, 0, Q)
        cpu = {"value": round(fps, 4), "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{args.cpu_sample_frames} frame(s) of the same 720p Q={Q} workload, one pass of the oracle port "
                         f"(reference's pure-PyTorch CPU path), fp32, all host threads, {dt:.1f} s"}
    metric = METRIC if args.config == 4 else "frames/sec, BASELINE config %d (%s)" % (args.config, workloads[args.config].split(":")[0])
    line = {"metric": metric, "value": round(T / ms_dev * 1e3, 2),
            "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": round(ms_dev, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32",
            "data": "synthetic (random backbone features, random-init weights, perturbed MSDeformAttn offsets)",
            "config": config, "clocks": clocks, "latency_ms_per_clip": round(ms_latency, 3),
            "e2e": {"value": round(T / ms_e2e * 1e3, 2), "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes * world,
                    "d2h_bytes_per_step": d2h_bytes * world, "ms_per_step": round(ms_e2e, 3)},
            "gpu_launches": launches, "parity_check": parity, "roofline": roof, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    sys.stdout.flush()
    if world > 1:
        # leave without tearing NCCL down rank by rank (ProcessGroupNCCL teardown can wait on peers that already left)
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)


# dram__bytes_read.sum + dram__bytes_write.sum of msda_fwd_staged_kernel per 720p frame, bf16 fused variant, from the committed
# `ncu --set full` capture of the head-major gather (8 frames per launch: 172.9 + 61.9 MB); refreshed whenever the kernel changes
MSDA_NCU_DRAM_BYTES_PER_FRAME = (172.921344e6 + 61.916416e6) / 8
MSDA_NCU_SOURCE = "imported: profiles/r2_ncu_msda.txt (ncu --set full of dvis_msda_fused_forward_hm, 8 frames per launch), scaled to this launch"


if __name__ == "__main__":
    main()
