"""The PRODUCT's B200 fast paths -- dvis_plus_b200.ops front ends and the module code that runs when tensors are on the
device with autograd off -- executed on CPU through the emulated library (tests/simt/emulated_device.py): every
libdvis_b200 kernel runs from its original source on the SIMT emulator (the tcgen05 mask GEMM is a plain-loop test double),
torch library calls run as CPU ops.  Same fixtures and tolerances as the -m gpu module tests (tests/test_modules_gpu.py):
golden outputs of the unmodified reference modules and the oracle port.  What this adds to the kernel-level emulator tests
is the host side of the fast path: operand layouts, strides, slices of fused projections, cached constants, kernel
selection."""
import os
import sys

import numpy as np
import pytest
import torch

from dvis_plus_b200 import _lib
from dvis_plus_b200 import modules as M
from dvis_plus_b200.modules.pixel_decoder import ShapeSpec
from dvis_plus_b200.modules.postprocess import VideoPostProcessor
from dvis_plus_b200.modules.precision import precision
from oracle import postprocess_port as pp
from postproc_util import assert_labels_match, assert_masks_match, sort_instances
from test_modules_cpu import build_predictor, build_refiner, build_tracker

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt"))
from emulated_device import emulated_b200  # noqa: E402

pytestmark = pytest.mark.timeout(1200)


def rel_err(a, b):
    assert a.shape == b.shape, (a.shape, b.shape)
    return (a.double() - b.double()).abs().max().item() / max(1e-6, b.abs().max().item())


@pytest.fixture
def device():
    with emulated_b200(), torch.no_grad():
        yield


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_msdeformattn_module(golden, device, mode, tol):
    g = golden("msdeformattn_module.pt")
    m = M.MSDeformAttn(d_model=64, n_levels=3, n_heads=8, n_points=4).eval()
    m.load_state_dict(g["state_dict"])
    sh = g["shapes"]
    lsi = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    calls = _lib.launch_count
    with precision(mode):
        out = m(g["query"], g["ref"], g["src"], sh, lsi, None)
        out_pad = m(g["query"], g["ref"], g["src"], sh, lsi, g["padding_mask"])
        out_box = m(g["query"], g["ref4"], g["src"], sh, lsi, None)
    assert _lib.launch_count >= calls + 3, "the kernel did not run"
    assert rel_err(out, g["out"]) < tol and rel_err(out_pad, g["out_pad"]) < tol and rel_err(out_box, g["out_box"]) < tol


def test_pixel_decoder_fused_path_vs_oracle(device):
    """conv_dim = 128 (the fused LayerNorm / GroupNorm kernels need C % 128 == 0): the channels-last fused encoder + FPN
    path of forward_features against the oracle port."""
    from oracle import torch_port as tp
    torch.manual_seed(0)
    chans = dict(res2=16, res3=24, res4=32, res5=48)
    strides = dict(res2=4, res3=8, res4=16, res5=32)
    pd = M.MSDeformAttnPixelDecoder({k: ShapeSpec(channels=chans[k], stride=strides[k]) for k in chans},
                                    transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=256,
                                    transformer_enc_layers=2, conv_dim=128, mask_dim=128, norm="GN",
                                    transformer_in_features=["res3", "res4", "res5"], common_stride=4).eval()
    for layer in pd.transformer.encoder.layers:
        torch.nn.init.normal_(layer.self_attn.sampling_offsets.weight, std=0.02)
        torch.nn.init.normal_(layer.self_attn.attention_weights.weight, std=0.1)
    feats = {k: torch.randn(2, chans[k], 64 // strides[k], 96 // strides[k]) for k in chans}
    sd = {k: v.detach() for k, v in pd.state_dict().items()}
    ref_mf, ref_o0, ref_ms = tp.pixel_decoder_forward_features(sd, feats, num_layers=2)
    assert pd._fused_ok()
    for mode, tol in (("fp32", 1e-3), ("bf16", 3e-2)):
        calls = _lib.launch_count
        with precision(mode):
            mf, o0, ms = pd.forward_features(feats)
        assert _lib.launch_count - calls == 2 * 3 + 3 + 2, "fused MSDA / LayerNorm / GroupNorm kernels did not run"
        assert mf.shape == ref_mf.shape and mf.is_contiguous(memory_format=torch.channels_last)
        assert rel_err(mf.float(), ref_mf) < tol, (mode, rel_err(mf.float(), ref_mf))
        assert rel_err(o0.float(), ref_o0) < tol
        for a, b in zip(ms, ref_ms):
            assert rel_err(a.float(), b) < tol


def test_pixel_decoder_tensor_core_projections_vs_oracle(device):
    """conv_dim = 256, 8 heads (32 channels per head, the production geometry): the encoder's value projection goes through
    dvis_linear_tc_heads (test double here) into the HEAD-MAJOR gather dvis_msda_fused_forward_hm (emulated kernel), and
    output_proj + residual + norm1 through dvis_linear_tc_add_ln; switching `use_tc_linear` off gives the library path."""
    from oracle import torch_port as tp
    torch.manual_seed(1)
    chans = dict(res2=8, res3=16, res4=24, res5=32)
    strides = dict(res2=4, res3=8, res4=16, res5=32)
    pd = M.MSDeformAttnPixelDecoder({k: ShapeSpec(channels=chans[k], stride=strides[k]) for k in chans},
                                    transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=128,
                                    transformer_enc_layers=2, conv_dim=256, mask_dim=128, norm="GN",
                                    transformer_in_features=["res3", "res4", "res5"], common_stride=4).eval()
    for layer in pd.transformer.encoder.layers:
        torch.nn.init.normal_(layer.self_attn.sampling_offsets.weight, std=0.02)
        torch.nn.init.normal_(layer.self_attn.attention_weights.weight, std=0.1)
        torch.nn.init.normal_(layer.self_attn.value_proj.bias, std=0.1)
        torch.nn.init.normal_(layer.self_attn.output_proj.bias, std=0.1)
    feats = {k: torch.randn(2, chans[k], 64 // strides[k], 96 // strides[k]) for k in chans}
    sd = {k: v.detach() for k, v in pd.state_dict().items()}
    ref_mf, ref_o0, ref_ms = tp.pixel_decoder_forward_features(sd, feats, num_layers=2)
    outs = {}
    for tc in (True, False):
        for layer in pd.transformer.encoder.layers:
            layer.self_attn.use_tc_linear = tc
            layer.self_attn.fuse_output_norm = tc
        calls = _lib.launch_count
        with precision("bf16"):
            mf, o0, ms = pd.forward_features(feats)
        assert _lib.launch_count - calls == 2 * (4 if tc else 3) + 3 + 2, "the expected kernels did not run"
        assert rel_err(mf.float(), ref_mf) < 3e-2 and rel_err(o0.float(), ref_o0) < 3e-2
        for a, b in zip(ms, ref_ms):
            assert rel_err(a.float(), b) < 3e-2
        outs[tc] = mf.float()
    assert rel_err(outs[True], outs[False]) < 2e-2


@pytest.mark.parametrize("materialize", [True, False])
def test_predictor_golden(golden, device, materialize):
    g = golden("predictor_small.pt")
    d = build_predictor(g)
    d.materialize_aux_masks = materialize
    calls = _lib.launch_count
    with precision("fp32"):
        out = d(list(g["multi_scale"]), g["mask_features"])
    assert _lib.launch_count > calls
    # downstream of thresholded bf16 mask logits (see tests/test_modules_gpu.py::test_predictor_golden); CPU bf16 GEMMs
    # round differently from cuBLAS, hence a little more slack than the 5e-2 used on the device
    for k in ("pred_logits", "pred_masks", "pred_embds", "pred_embds_without_norm"):
        assert rel_err(out[k].float(), g[k]) < 8e-2, (k, rel_err(out[k].float(), g[k]))
    if materialize:
        assert len(out["aux_outputs"]) == 3


def test_tracker_golden_with_emulated_hungarian(golden, device):
    g = golden("tracker_small.pt")
    t = build_tracker(g)
    t.use_cuda_graph = False                      # graph capture is a CUDA-runtime feature; the frame body is the same code
    fe, fn, mf = g["frame_embeds"], g["frame_embeds_no_norm"], g["mask_features"]
    with precision("bf16"):
        o1, i1 = t(fe[:, :, :2], mf[:, :2], resume=False, return_indices=True, frame_embeds_no_norm=fn[:, :, :2])
        o2, i2 = t(fe[:, :, 2:], mf[:, 2:], resume=True, return_indices=True, frame_embeds_no_norm=fn[:, :, 2:])
    for a, b in zip(i1 + i2, g["indices"]):
        assert np.array_equal(np.asarray(a), b.numpy()), "emulated GPU Hungarian differs from the reference's SciPy result"
    assert rel_err(torch.cat([o1["pred_embds"], o2["pred_embds"]], 2).float(), g["pred_embds"]) < 3e-2
    assert rel_err(torch.cat([o1["pred_logits"], o2["pred_logits"]], 1).float(), g["pred_logits"]) < 3e-2
    assert rel_err(torch.cat([o1["pred_masks"], o2["pred_masks"]], 2).float(), g["pred_masks"]) < 3e-2


def test_refiner_golden(golden, device):
    g = golden("refiner_small.pt")
    r = build_refiner(g)
    with precision("bf16"):
        o = r(g["instance_embeds"], g["frame_embeds"], g["mask_features"])
    for k in ("pred_embds", "pred_logits", "pred_masks"):
        assert rel_err(o[k].float(), g[k]) < 3e-2, k


def test_refiner_row_layout_equals_module_path(golden, device):
    """bf16 inference, batch 1: TemporalRefiner._refine_rows (tokens as (t, q, c) rows, no permutation copies, convolutions over
    time as gather + GEMM) against the module-by-module path of the same class; the former is what the default path runs."""
    g = golden("refiner_small.pt")
    r = build_refiner(g)
    outs = {}
    for rows in (True, False):
        r.use_row_layout = rows
        calls = _lib.launch_count
        with precision("bf16"):
            outs[rows] = r(g["instance_embeds"], g["frame_embeds"], g["mask_features"])
        outs[rows]["launches"] = _lib.launch_count - calls
    for k in ("pred_embds", "pred_logits", "pred_masks"):
        assert rel_err(outs[True][k].float(), outs[False][k].float()) < 2e-2, k
        assert rel_err(outs[True][k].float(), g[k]) < 3e-2, k


def test_postprocessor_with_emulated_kernels_vs_reference_golden(golden, device):
    g = golden("postprocess_vis.pt")
    for name, c in g["cases"].items():
        Ho, Wo = c["output_size"]
        for packed in (False, True):
            post = VideoPostProcessor(g["num_classes"], num_queries=12, max_num=c["max_num"])
            post.packed_transfer = packed
            aux = g["aux_cls"] if c["use_aux"] else None
            out = post.inference_video_task(g["pred_cls"], g["pred_masks"], g["img_size"], Ho, Wo, g["first_resize_size"], g["pred_id"],
                                            aux_pred_cls=aux)
            s, l, i, m = sort_instances(out["pred_scores"], out["pred_labels"], out["pred_ids"], torch.stack(out["pred_masks"]))
            rs, rl, ri, rm = sort_instances(c["pred_scores"], c["pred_labels"], c["pred_ids"], c["pred_masks"])
            torch.testing.assert_close(s, rs, rtol=1e-5, atol=1e-7)
            assert torch.equal(l, rl) and torch.equal(i, ri), name
            o = pp.inference_video_vis(g["pred_cls"], g["pred_masks"], g["img_size"], Ho, Wo, g["first_resize_size"], g["pred_id"],
                                       g["num_classes"], c["max_num"], aux_pred_cls=aux, return_logits=True)
            _, _, _, lg = sort_instances(o["pred_scores"], o["pred_labels"], o["pred_ids"], o["resized_logits"])
            assert_masks_match(m, rm, lg, tol=2e-5)
    g = golden("postprocess_vps.pt")
    for name, c in g["cases"].items():
        Ho, Wo = c["output_size"]
        post = VideoPostProcessor(g["num_classes"], object_mask_threshold=c["object_mask_threshold"],
                                  overlap_threshold=c["overlap_threshold"], num_thing_classes=g["num_thing_classes"], task="vps")
        out = post.inference_video_task(g["pred_cls"], g["pred_masks"], g["img_size"], Ho, Wo, g["first_resize_size"], g["pred_id"],
                                        aux_pred_cls=g["aux_cls"] if c["use_aux"] else None)
        assert out["segments_infos"] == c["segments_infos"] and [int(i) for i in out["pred_ids"]] == c["pred_ids"], name
        assert (out["pred_masks"] != c["pred_masks"]).float().mean().item() < 1e-3, name
    g = golden("postprocess_vss.pt")
    for name, c in g["cases"].items():
        Ho, Wo = c["output_size"]
        aux = g["aux_cls"] if c["use_aux"] else None
        out = VideoPostProcessor(g["num_classes"], task="vss").inference_video_task(
            g["pred_cls"], g["pred_masks"], g["img_size"], Ho, Wo, g["first_resize_size"], None, aux_pred_cls=aux)
        ref = pp.inference_video_vss(g["pred_cls"], g["pred_masks"], g["img_size"], Ho, Wo, g["first_resize_size"], aux_pred_cls=aux,
                                     return_scores=True)
        assert_labels_match(out["pred_masks"], c["pred_masks"], ref["semseg"], tol=1e-5)
    g = golden("postprocess_logits.pt")
    mv = VideoPostProcessor(5).post_processing_minvis(dict(pred_logits=g["pred_logits"].clone(), pred_masks=g["pred_masks"].clone(),
                                                           pred_embds=g["pred_embds"].clone()))
    torch.testing.assert_close(mv["pred_logits"], g["minvis_logits"], rtol=1e-5, atol=1e-6)
    assert torch.equal(mv["pred_masks"], g["minvis_masks"])


def test_pipeline_vis_from_block_on_the_emulated_device(device):
    from dvis_plus_b200.pipeline import OfflineClipRunner
    T, Q, C, K, H, W = 3, 12, 64, 5, 8, 12
    torch.manual_seed(0)
    trk = M.ReferringTracker_noiser(hidden_channel=C, feedforward_channel=128, num_head=8, decoder_layer_num=2, mask_dim=64,
                                    class_num=K, noise_mode="none").eval()
    trk.use_cuda_graph = False
    rfn = M.TemporalRefiner(hidden_channel=C, feedforward_channel=128, num_head=8, decoder_layer_num=2, mask_dim=64, class_num=K,
                            windows=2).eval()
    g = torch.Generator().manual_seed(1)
    seg = dict(pred_embds=torch.randn(1, C, T, Q, generator=g), pred_embds_without_norm=torch.randn(1, C, T, Q, generator=g),
               pred_logits=torch.randn(1, T, Q, K + 1, generator=g))
    mf = torch.randn(T, 64, H, W, generator=g).to(torch.bfloat16, memory_format=torch.channels_last)
    runner = OfflineClipRunner(None, None, trk, rfn)
    post = VideoPostProcessor(K, num_queries=Q, max_num=5)
    img, out_size = (30, 45), (41, 60)
    with precision("bf16"):
        block = runner.pack_queries(seg)
        fused = runner.vis_from_block(block, mf, C, post, img, out_size)
        full = runner.temporal_from_block(block, mf, C)
    outs, aux = post.post_processing(dict(pred_logits=full["pred_logits"], pred_masks=full["pred_masks"]),
                                     aux_logits=full["online_pred_logits"])
    ref = post.inference_video_vis(outs["pred_logits"][0], outs["pred_masks"][0], img, *out_size, (4 * H, 4 * W), outs["ids"][0],
                                   aux_pred_cls=aux)
    torch.testing.assert_close(fused["pred_scores"], torch.tensor(ref["pred_scores"]), rtol=1e-5, atol=1e-7)
    assert fused["pred_labels"].tolist() == ref["pred_labels"] and fused["pred_ids"].tolist() == ref["pred_ids"]
    assert fused["pred_masks"].shape == (5, T, *out_size) and fused["pred_masks"].dtype == torch.bool
    assert (fused["pred_masks"] != torch.stack(ref["pred_masks"])).float().mean().item() < 1e-3


def test_daq_track_query_matching_device_path_equals_host(device):
    g = torch.Generator().manual_seed(21)
    cutter = M.VideoInstanceCutter(hidden_dim=64, feedforward_dim=128, num_head=8, decoder_layer_num=1, mask_dim=64, num_classes=5).eval()
    for n_trk, n_seg in ((7, 20), (20, 20), (25, 12)):
        trc, seg = torch.randn(n_trk, 1, 64, generator=g), torch.randn(n_seg, 1, 64, generator=g)
        cutter.match_on_host = True
        host = cutter.match_with_embeds(trc, seg)
        cutter.match_on_host = False
        calls = _lib.launch_count
        assert torch.equal(cutter.match_with_embeds(trc, seg), host)
        assert _lib.launch_count == calls + 1


def test_predictor_fast_path_production_width(device):
    """hidden_dim 128 (multiple of 128 -> the batch-first fused inference path: fused LayerNorm kernels, resized mask
    features, attention-bias GEMM epilogue) against the oracle port."""
    from oracle import torch_port as tp
    torch.manual_seed(0)
    d = M.VideoMultiScaleMaskedTransformerDecoder_dvisPlus(
        128, True, num_classes=7, hidden_dim=128, num_queries=12, nheads=8, dim_feedforward=256, dec_layers=3,
        pre_norm=False, mask_dim=128, enforce_input_project=False, num_frames=1, num_reid_head_layers=3,
        reid_hidden_dim=128).eval()
    ms = [torch.randn(2, 128, 2, 3), torch.randn(2, 128, 4, 6), torch.randn(2, 128, 8, 12)]
    mf = torch.randn(2, 128, 16, 24)
    ref = tp.predictor_forward({k: v.detach() for k, v in d.state_dict().items()}, ms, mf, num_layers=3)
    calls = _lib.launch_count
    with precision("fp32"):
        out = d(ms, mf)
    assert _lib.launch_count - calls > 10, "fused LayerNorm / mask kernels did not run"
    for k in ("pred_logits", "pred_masks", "pred_embds", "pred_embds_without_norm"):
        assert rel_err(out[k].float(), ref[k]) < 8e-2, (k, rel_err(out[k].float(), ref[k]))
    with precision("bf16"):
        out = d(ms, mf)
    for k in ("pred_logits", "pred_masks", "pred_embds"):
        assert rel_err(out[k].float(), ref[k]) < 0.12, (k, rel_err(out[k].float(), ref[k]))


def test_predictor_prenorm_variant_golden(golden, device):
    """pre_norm=True + enforce_input_project + no ReID head: the generic loop (the fused path is post-norm only)."""
    g, base = golden("predictor_prenorm_small.pt"), golden("predictor_small.pt")
    d = M.VideoMultiScaleMaskedTransformerDecoder_dvisPlus(
        64, True, num_classes=5, hidden_dim=64, num_queries=12, nheads=8, dim_feedforward=128, dec_layers=2, pre_norm=True,
        mask_dim=64, enforce_input_project=True, num_frames=2, num_reid_head_layers=0, reid_hidden_dim=64).eval()
    d.load_state_dict(g["state_dict"])
    with precision("fp32"):
        out = d(list(base["multi_scale"]), base["mask_features"])
    for k in ("pred_logits", "pred_masks", "pred_embds", "pred_embds_without_norm"):
        assert rel_err(out[k].float(), g[k]) < 8e-2, (k, rel_err(out[k].float(), g[k]))


def test_mask_logits_query_slices_beyond_256(device):
    """Q = 300 (the DAQ stress size): ops.mask_logits splits the queries into in-place slices of <= 256 and
    ops.mask_attn_bias takes its two-kernel route; both against plain torch."""
    from dvis_plus_b200 import ops
    g = torch.Generator().manual_seed(3)
    emb = torch.randn(2, 300, 64, generator=g)
    feat = torch.randn(2, 64, 5, 7, generator=g)
    eb, fb = emb.bfloat16().float(), feat.bfloat16().float()
    ref = torch.einsum("bqc,bchw->bqhw", eb, fb)
    calls = _lib.launch_count
    out = ops.mask_logits(emb, feat.to(torch.bfloat16, memory_format=torch.channels_last), torch.float32)
    assert _lib.launch_count == calls + 2                                       # two strided launches
    assert rel_err(out, ref) < 1e-5
    bias = ops.mask_attn_bias(emb, feat.to(torch.bfloat16, memory_format=torch.channels_last), torch.float32)
    masked = ref.flatten(2) < 0
    masked[masked.all(-1)] = False
    assert torch.equal(torch.isinf(bias) & (bias < 0), masked)


# ---- round 2: the fused temporal-stage kernels (csrc/small_linear.cu, csrc/flash_attn.cu) under the module fast paths ----

def build_tracker_w128(g):
    t = M.ReferringTracker_noiser(hidden_channel=128, feedforward_channel=256, num_head=4, decoder_layer_num=2,
                                  mask_dim=64, class_num=5, noise_mode="none").eval()
    assert not any(t.load_state_dict(g["state_dict"]))
    return t


def build_refiner_w128(g):
    r = M.TemporalRefiner(hidden_channel=128, feedforward_channel=256, num_head=4, decoder_layer_num=2, mask_dim=64,
                          class_num=5, windows=3).eval()
    assert not any(r.load_state_dict(g["state_dict"]))
    return r


def test_tracker_fused_kernels_vs_reference_golden(golden, device):
    """hidden 128 (4 heads x 32): first frame, later frames and a resumed window all run on dvis_linear_small /
    dvis_flash_attn only -- against the unmodified reference's outputs (tests/golden/make_golden_r2.py)."""
    g = golden("tracker_w128.pt")
    t = build_tracker_w128(g)
    t.use_cuda_graph = False
    t.use_fused_kernels = True                    # opt-in: every linear step on csrc/small_linear.cu
    fe, fn, mf = g["frame_embeds"], g["frame_embeds_no_norm"], g["mask_features"]
    calls = _lib.launch_count
    with precision("bf16"):
        assert t._fused_ok(t._stacked(torch.bfloat16))
        o1, i1 = t(fe[:, :, :3], mf[:, :3], resume=False, return_indices=True, frame_embeds_no_norm=fn[:, :, :3])
        o2, i2 = t(fe[:, :, 3:], mf[:, 3:], resume=True, return_indices=True, frame_embeds_no_norm=fn[:, :, 3:])
    # window 1: kv + first frame (2 layers x 11) + 2 later frames (6 + 2 x 5) + final LN; window 2: kv + 1 frame + final LN (+ LAP, masks)
    assert _lib.launch_count - calls >= (1 + 22 + 2 * 16 + 1) + (1 + 16 + 1)
    for a, b in zip(i1 + i2, g["indices"]):
        assert np.array_equal(np.asarray(a), b.numpy())
    assert rel_err(torch.cat([o1["pred_embds"], o2["pred_embds"]], 2).float(), g["pred_embds"]) < 3e-2
    assert rel_err(torch.cat([o1["pred_references"], o2["pred_references"]], 2).float(), g["pred_references"]) < 3e-2
    assert rel_err(torch.cat([o1["pred_logits"], o2["pred_logits"]], 1).float(), g["pred_logits"]) < 3e-2
    assert rel_err(torch.cat([o1["pred_masks"], o2["pred_masks"]], 2).float(), g["pred_masks"]) < 3e-2
    # the recurrent state keeps the reference's layout: (1 + layers, q, b, c)
    assert t.last_outputs.shape == (3, 10, 1, 128) and t.last_reference.shape == (10, 1, 128)


def test_tracker_default_path_with_flash_attention_vs_reference_golden(golden, device):
    """the DEFAULT bf16 path at a width the attention kernel covers (4 heads x 32): library GEMMs + dvis_flash_attn"""
    g = golden("tracker_w128.pt")
    t = build_tracker_w128(g)
    t.use_cuda_graph = False
    assert t.use_custom_attention and not t.use_fused_kernels
    fe, fn, mf = g["frame_embeds"], g["frame_embeds_no_norm"], g["mask_features"]
    with precision("bf16"):
        o1 = t(fe[:, :, :3], mf[:, :3], resume=False, frame_embeds_no_norm=fn[:, :, :3])
        o2 = t(fe[:, :, 3:], mf[:, 3:], resume=True, frame_embeds_no_norm=fn[:, :, 3:])
    assert rel_err(torch.cat([o1["pred_embds"], o2["pred_embds"]], 2).float(), g["pred_embds"]) < 3e-2
    assert rel_err(torch.cat([o1["pred_masks"], o2["pred_masks"]], 2).float(), g["pred_masks"]) < 3e-2


def test_tracker_fused_equals_library_path(golden, device):
    g = golden("tracker_w128.pt")
    fe, fn = g["frame_embeds"], g["frame_embeds_no_norm"]
    outs = []
    for fused in (True, False):
        t = build_tracker_w128(g)
        t.use_cuda_graph = False
        t.use_fused_kernels = fused
        with precision("bf16"):
            outs.append(t(fe, None, resume=False, frame_embeds_no_norm=fn, with_masks=False))
    assert rel_err(outs[0]["pred_embds"].float(), outs[1]["pred_embds"].float()) < 2e-2
    assert rel_err(outs[0]["pred_logits"].float(), outs[1]["pred_logits"].float()) < 2e-2


def test_refiner_fused_kernels_vs_reference_golden(golden, device):
    """hidden 128, T=7: time attention through strided views, Conv1d k5 / k3 as tap GEMMs with replicate padding, object
    and cross attention -- all on dvis_linear_small / dvis_flash_attn -- against the unmodified reference."""
    g = golden("refiner_w128.pt")
    r = build_refiner_w128(g)
    r.use_fused_kernels = True
    calls = _lib.launch_count
    with precision("bf16"):
        o = r(g["instance_embeds"], g["frame_embeds"], g["mask_features"])
    assert _lib.launch_count - calls >= 1 + 2 * 14 + 1
    for k in ("pred_embds", "pred_logits", "pred_masks"):
        assert rel_err(o[k].float(), g[k]) < 3e-2, (k, rel_err(o[k].float(), g[k]))
    r.use_fused_kernels = False
    with precision("bf16"):
        o_lib = r(g["instance_embeds"], g["frame_embeds"], g["mask_features"])
    assert rel_err(o["pred_embds"].float(), o_lib["pred_embds"].float()) < 2e-2


def test_predictor_level_tokens_kernel_path_equals_torch_path(device):
    """channels-last level maps (what the pixel decoder hands over) take dvis_level_tokens (x + level_embed [+ pos] -> bf16 in
    one pass); NCHW-contiguous maps take the torch ops: same values up to the order of the two fp32 additions"""
    torch.manual_seed(1)
    d = M.VideoMultiScaleMaskedTransformerDecoder_dvisPlus(
        128, True, num_classes=7, hidden_dim=128, num_queries=12, nheads=4, dim_feedforward=256, dec_layers=3,
        pre_norm=False, mask_dim=128, enforce_input_project=False, num_frames=1, num_reid_head_layers=3,
        reid_hidden_dim=128).eval()
    ms = [torch.randn(2, 128, 2, 3), torch.randn(2, 128, 4, 6), torch.randn(2, 128, 8, 12)]
    mf = torch.randn(2, 128, 16, 24)
    with precision("bf16"):
        ref = d(ms, mf)
        calls = _lib.launch_count
        out = d([m.contiguous(memory_format=torch.channels_last) for m in ms], mf)
        tok, key = __import__("dvis_plus_b200").ops.level_tokens(ms[1].contiguous(memory_format=torch.channels_last),
                                                                 d.level_embed.weight[1].detach().float().contiguous(),
                                                                 d._pos(4, 6, ms[1].device)[:, 0].contiguous())
    assert _lib.launch_count - calls > 3
    want = ms[1].permute(0, 2, 3, 1).reshape(2, 24, 128) + d.level_embed.weight[1]
    assert torch.equal(tok, want.to(torch.bfloat16))
    assert (key.float() - (want + d._pos(4, 6, ms[1].device)[:, 0])).abs().max() < 2e-2
    for k in ("pred_logits", "pred_embds"):
        assert rel_err(out[k].float(), ref[k].float()) < 2e-2, k
