"""BASELINE.json configs at FULL size on the GPU against the oracle port (CPU, fp32) on identical seeded inputs and
weights.  The oracle needs a few seconds per case on the host.

  config 2: R50 MSDeformAttnPixelDecoder, 720p single frame, + the Q=100 predictor (10 mask-head calls)
  config 3: T=5 clip, Q=200, ReferringTracker (hidden 512) -- the online model's final masks come from the tracker
  config 4: T=16, Q=200 TemporalRefiner (hidden 512), via the frame-sharded pipeline's temporal stage
Tolerances: bf16 GEMM operands vs the fp32 oracle -> 2e-2 of the output scale everywhere (north star: 1e-2 per bf16 op; these
are 6..9 stacked layers).  Measured on the B200 (profiles/r2_test_errors.json, tests/perf/report_test_errors.py): 7e-3 .. 1.2e-2;
round 1 allowed 3e-2, and 8e-2 downstream of the thresholded attention masks (then a dense bf16 bias, now bits from fp32 logits).
"""
import pytest
import torch

from dvis_plus_b200 import _lib
from dvis_plus_b200.modules.precision import precision
from oracle import torch_port as tp

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    assert a.shape == b.shape, (a.shape, b.shape)
    return (a.double().cpu() - b.double()).abs().max().item() / max(1e-6, b.abs().max().item())


def sd_of(m):
    return {k: v.detach().float().cpu() for k, v in m.state_dict().items()}


@torch.no_grad()
def test_config2_r50_pixel_decoder_and_predictor_720p():
    import bench
    runner = bench.build_models("cpu", queries=100, backbone="r50")
    feats = {k: v.float().contiguous() for k, v in bench.synthetic_features(1, "r50", seed=3, dtype=torch.float32).items()}
    pd_sd, dec_sd = sd_of(runner.pixel_decoder), sd_of(runner.predictor)
    ref_mf, ref_o0, ref_ms = tp.pixel_decoder_forward_features(pd_sd, feats, num_layers=6)
    ref_seg = tp.predictor_forward(dec_sd, ref_ms, ref_mf, num_layers=9)
    pd, dec = runner.pixel_decoder.cuda(), runner.predictor.cuda()
    n0 = _lib.launch_count
    with precision("bf16"):
        mf, o0, ms = pd.forward_features({k: v.cuda() for k, v in feats.items()})
        assert mf.shape == (1, 256, 184, 320) and mf.dtype == torch.bfloat16
        assert rel_err(mf.float(), ref_mf) < 2e-2
        assert rel_err(o0.float(), ref_o0) < 2e-2
        for a, b in zip(ms, ref_ms):
            assert rel_err(a.float(), b) < 2e-2
        # predictor on the ORACLE's pixel-decoder outputs so that only the predictor's own error is measured
        seg = dec([m.cuda() for m in ref_ms], ref_mf.cuda())
    assert _lib.launch_count - n0 > 20, "libdvis_b200 kernels did not run"
    assert seg["pred_masks"].shape == (1, 100, 1, 184, 320)
    for k in ("pred_logits", "pred_masks", "pred_embds"):
        assert rel_err(seg[k].float(), ref_seg[k]) < 2e-2, (k, rel_err(seg[k].float(), ref_seg[k]))


@torch.no_grad()
def test_config3_online_tracker_T5_Q200():
    import bench
    torch.manual_seed(0)
    runner = bench.build_models("cpu", queries=200)
    trk = runner.tracker
    T, Q = 5, 200
    base = torch.randn(1, 512, 1, Q)
    fe = base + 0.3 * torch.randn(1, 512, T, Q)               # frames of one video: related query embeddings
    fn = fe + 0.1 * torch.randn(1, 512, T, Q)
    mfeat = torch.randn(1, T, 256, 46, 80)                    # 1/16-size maps keep the CPU oracle fast; masks are checked
    ref = tp.tracker_forward(sd_of(trk), fe, mfeat, fn, num_layers=6)
    trk = trk.cuda()
    with precision("bf16"):
        out, idx = trk(fe.cuda(), mfeat.cuda(), resume=False, return_indices=True, frame_embeds_no_norm=fn.cuda())
    for a, b in zip(idx, ref["indices"]):
        assert (torch.as_tensor(a) == torch.as_tensor(b)).all(), "GPU Hungarian differs from SciPy"
    assert rel_err(out["pred_embds"].float(), ref["pred_embds"]) < 2e-2
    assert rel_err(out["pred_logits"].float(), ref["pred_logits"]) < 2e-2
    assert rel_err(out["pred_masks"].float(), ref["pred_masks"]) < 2e-2


@torch.no_grad()
def test_config4_offline_temporal_stage_T16_Q200():
    import bench
    from dvis_plus_b200.pipeline import OfflineClipRunner
    torch.manual_seed(1)
    runner = bench.build_models("cpu", queries=200)
    T, Q, K1 = 16, 200, bench.NUM_CLASSES + 1
    base = torch.randn(1, 512, 1, Q)
    seg = dict(pred_embds=base + 0.3 * torch.randn(1, 512, T, Q), pred_logits=torch.randn(1, T, Q, K1))
    seg["pred_embds_without_norm"] = seg["pred_embds"] + 0.1 * torch.randn(1, 512, T, Q)
    mfeat = torch.randn(T, 256, 46, 80)
    trk_ref = tp.tracker_forward(sd_of(runner.tracker), seg["pred_embds"], None, seg["pred_embds_without_norm"], num_layers=6, with_masks=False)
    ref = tp.refiner_forward(sd_of(runner.refiner), trk_ref["pred_embds"], seg["pred_embds_without_norm"], mfeat[None], num_layers=6)
    r = OfflineClipRunner(None, None, runner.tracker.cuda(), runner.refiner.cuda())
    with precision("bf16"):
        out = r.temporal_stage({k: v.cuda() for k, v in seg.items()}, mfeat.cuda().to(torch.bfloat16, memory_format=torch.channels_last))
    assert out["pred_masks"].shape == (1, Q, T, 46, 80)
    assert rel_err(out["pred_embds"].float(), ref["pred_embds"]) < 2e-2
    assert rel_err(out["pred_logits"].float(), ref["pred_logits"]) < 2e-2
    assert rel_err(out["pred_masks"].float(), ref["pred_masks"]) < 2e-2
