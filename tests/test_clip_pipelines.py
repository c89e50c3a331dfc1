"""Clip-level runners on the emulated device.  pipeline.OnlineClipRunner = DVIS_Plus_online.run_window_inference between backbone and post-processing
(P/dvis_Plus/meta_architecture.py:774-816): windows through the segmenter head and the referring tracker with state carried
across windows.  Checked against the oracle port chained the same way (oracle/torch_port.py, `state=`), on the modules'
autograd / CPU branch in fp32 and on the emulated device (B200 fast path, bf16 GEMMs)."""
import os
import sys

import pytest
import torch

from dvis_plus_b200 import modules as M
from dvis_plus_b200.modules.pixel_decoder import ShapeSpec
from dvis_plus_b200.modules.postprocess import VideoPostProcessor
from dvis_plus_b200.modules.precision import precision
from dvis_plus_b200.pipeline import OnlineClipRunner
from oracle import torch_port as tp

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt"))

CH = dict(res2=16, res3=24, res4=32, res5=48)
STRIDES = dict(res2=4, res3=8, res4=16, res5=32)
T, WINDOW, Q, K, HID = 5, 2, 10, 5, 128           # 3 windows: 2 + 2 + 1 frames


def build():
    torch.manual_seed(0)
    pd = M.MSDeformAttnPixelDecoder({k: ShapeSpec(channels=CH[k], stride=STRIDES[k]) for k in CH}, transformer_dropout=0.0,
                                    transformer_nheads=8, transformer_dim_feedforward=256, transformer_enc_layers=2, conv_dim=HID,
                                    mask_dim=HID, norm="GN", transformer_in_features=["res3", "res4", "res5"], common_stride=4).eval()
    for layer in pd.transformer.encoder.layers:
        torch.nn.init.normal_(layer.self_attn.sampling_offsets.weight, std=0.02)
        torch.nn.init.normal_(layer.self_attn.attention_weights.weight, std=0.1)
    dec = M.VideoMultiScaleMaskedTransformerDecoder_dvisPlus(
        HID, True, num_classes=K, hidden_dim=HID, num_queries=Q, nheads=8, dim_feedforward=256, dec_layers=3, pre_norm=False,
        mask_dim=HID, enforce_input_project=False, num_frames=WINDOW, num_reid_head_layers=3, reid_hidden_dim=HID).eval()
    # the segmenter's query embedding is [decoder output | ReID embedding] (meta_architecture.py:550-553): 2 * HID channels
    trk = M.ReferringTracker_noiser(hidden_channel=2 * HID, feedforward_channel=256, num_head=8, decoder_layer_num=2, mask_dim=HID,
                                    class_num=K, noise_mode="none").eval()
    trk.use_cuda_graph = False
    feats = {k: torch.randn(T, CH[k], 32 // STRIDES[k], 64 // STRIDES[k]) for k in CH}
    return pd, dec, trk, feats


def oracle_chain(pd, dec, trk, feats):
    sd = lambda m: {k: v.detach().float() for k, v in m.state_dict().items()}
    state, logits, masks, embds = None, [], [], []
    for s in range(0, T, WINDOW):
        win = {k: v[s:s + WINDOW] for k, v in feats.items()}
        mf, _, ms = tp.pixel_decoder_forward_features(sd(pd), win, num_layers=2)
        seg = tp.predictor_forward(sd(dec), ms, mf, num_layers=3)
        out = tp.tracker_forward(sd(trk), seg["pred_embds"], mf[None], seg["pred_embds_without_norm"], num_layers=2, state=state)
        state = out["state"]
        logits.append(out["pred_logits"]); masks.append(out["pred_masks"]); embds.append(out["pred_embds"])
    return torch.cat(logits, 1), torch.cat(masks, 2), torch.cat(embds, 2)


def rel_err(a, b):
    assert a.shape == b.shape, (a.shape, b.shape)
    return (a.double() - b.double()).abs().max().item() / max(1e-6, b.abs().max().item())


@pytest.mark.timeout(1200)
@torch.no_grad()
def test_online_windows_carry_the_tracker_state():
    """The window loop and the resume / keep flags: identical to chaining the modules by hand,
    and the whole thing agrees with the oracle chain.  MSDeformAttn has no CPU implementation (like the reference's
    extension), so this runs on the emulated device."""
    from emulated_device import emulated_b200
    pd, dec, trk, feats = build()
    ref_logits, ref_masks, ref_embds = oracle_chain(pd, dec, trk, feats)
    with emulated_b200(), precision("fp32"):
        out = OnlineClipRunner(pd, dec, trk, window_size=WINDOW)(feats)
        logits, masks = [], []
        for i, s in enumerate(range(0, T, WINDOW)):
            mf, _, ms = pd.forward_features({k: v[s:s + WINDOW] for k, v in feats.items()})
            seg = dec(ms, mf)
            o = trk(seg["pred_embds"], mf.unsqueeze(0), resume=i != 0, frame_embeds_no_norm=seg["pred_embds_without_norm"])
            logits.append(o["pred_logits"].float()); masks.append(o["pred_masks"])
        # (not bit-equal: the GroupNorm statistics are accumulated with atomics, whose order varies from run to run)
        assert rel_err(out["pred_logits"], torch.cat(logits, 1)) < 1e-4 and rel_err(out["pred_masks"], torch.cat(masks, 2)) < 5e-3   # bf16 mask features: a last-bit flip is 4e-3 relative
        # a second call with keep=True continues the video: its first window resumes from the previous call's last frame
        cont = OnlineClipRunner(pd, dec, trk, window_size=WINDOW)({k: v[:WINDOW] for k, v in feats.items()}, keep=True)
        fresh = OnlineClipRunner(pd, dec, trk, window_size=WINDOW)({k: v[:WINDOW] for k, v in feats.items()}, keep=False)
        assert rel_err(cont["pred_logits"], fresh["pred_logits"]) > 1e-2
        assert rel_err(fresh["pred_logits"], out["pred_logits"][:, :WINDOW]) < 1e-4
    assert out["pred_logits"].shape == (1, T, Q, K + 1) and out["pred_masks"].shape[:3] == (1, Q, T)
    assert rel_err(out["pred_logits"], ref_logits) < 8e-2        # bf16 mask logits are thresholded inside the predictor
    assert rel_err(out["pred_embds"], ref_embds) < 8e-2
    assert rel_err(out["pred_masks"].float(), ref_masks) < 8e-2


@pytest.mark.timeout(1200)
@torch.no_grad()
def test_online_windows_on_the_emulated_device_and_vis_postprocessing():
    from emulated_device import emulated_b200
    from dvis_plus_b200 import _lib
    pd, dec, trk, feats = build()
    ref_logits, ref_masks, ref_embds = oracle_chain(pd, dec, trk, feats)
    with emulated_b200(), precision("bf16"):
        calls = _lib.launch_count
        out = OnlineClipRunner(pd, dec, trk, window_size=WINDOW)(feats)
        assert _lib.launch_count - calls > 30, "libdvis kernels did not run"
        # bf16 GEMMs + thresholded attention masks in the predictor (see tests/test_modules_gpu.py); on these tiny maps
        # (8 x 16 mask pixels) a single flipped attention-mask bit moves the output by ~1e-1 of its scale
        assert rel_err(out["pred_logits"], ref_logits) < 0.15
        assert rel_err(out["pred_embds"], ref_embds) < 0.15
        assert rel_err(out["pred_masks"].float(), ref_masks) < 0.15
        # the rest of DVIS_Plus_online.forward's eval branch (py:686-706): post_processing + inference_video_vis
        post = VideoPostProcessor(K, num_queries=Q, max_num=4)
        o = post.post_processing(dict(out))
        res = post.inference_video_task(o["pred_logits"][0], o["pred_masks"][0], (30, 60), 45, 91, (32, 64), o["ids"][0])
        assert len(res["pred_masks"]) == 4 and res["pred_masks"][0].shape == (T, 45, 91) and res["task"] == "vis"
        mv = post.inference_video(o["pred_logits"][0], o["pred_masks"][0], (30, 60), 45, 91, (32, 64))
        assert set(mv) == {"image_size", "pred_scores", "pred_labels", "pred_masks"} and len(mv["pred_masks"]) == 10


@pytest.mark.timeout(1200)
@torch.no_grad()
def test_offline_clip_runner_on_the_emulated_device_vs_oracle_chain():
    """The benchmark's pipeline (OfflineClipRunner.__call__: pixel decoder -> predictor -> tracker -> refiner -> final mask
    GEMM) end to end on the emulated device against the same chain of oracle functions bench.py times as the CPU baseline."""
    from emulated_device import emulated_b200
    from dvis_plus_b200.pipeline import OfflineClipRunner
    pd, dec, trk, feats = build()
    torch.manual_seed(5)
    rfn = M.TemporalRefiner(hidden_channel=2 * HID, feedforward_channel=256, num_head=8, decoder_layer_num=2, mask_dim=HID,
                            class_num=K, windows=T).eval()
    dec.num_frames = T
    sd = lambda m: {k: v.detach().float() for k, v in m.state_dict().items()}
    mf, _, ms = tp.pixel_decoder_forward_features(sd(pd), feats, num_layers=2)
    seg = tp.predictor_forward(sd(dec), ms, mf, num_layers=3)
    trk_ref = tp.tracker_forward(sd(trk), seg["pred_embds"], None, seg["pred_embds_without_norm"], num_layers=2, with_masks=False)
    ref = tp.refiner_forward(sd(rfn), trk_ref["pred_embds"], seg["pred_embds_without_norm"], mf[None], num_layers=2)
    with emulated_b200(), precision("fp32"):
        out = OfflineClipRunner(pd, dec, trk, rfn)(feats)
    assert out["pred_masks"].shape == ref["pred_masks"].shape == (1, Q, T, 8, 16)
    for k in ("pred_logits", "pred_embds", "pred_masks"):
        assert rel_err(out[k].float(), ref[k]) < 0.15, (k, rel_err(out[k].float(), ref[k]))   # thresholded bf16 attention masks
