"""Host-side logic of the module drop-ins on CPU (torch ops only, no kernels) against the reference's golden outputs:
state_dict compatibility, recurrent tracker state, window handling, output dictionary layout."""
import numpy as np
import pytest
import torch

from dvis_plus_b200 import modules as M


def close(a, b, tol=2e-4):
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a.double() - b.double()).abs().max().item()
    assert err <= tol * max(1.0, b.abs().max().item()), err


def build_tracker(g):
    t = M.ReferringTracker_noiser(hidden_channel=64, feedforward_channel=128, num_head=8, decoder_layer_num=2,
                                  mask_dim=64, class_num=5, noise_mode="none").eval()
    assert not any(t.load_state_dict(g["state_dict"]))
    return t


def build_refiner(g):
    r = M.TemporalRefiner(hidden_channel=64, feedforward_channel=128, num_head=8, decoder_layer_num=2, mask_dim=64,
                          class_num=5, windows=3).eval()
    assert not any(r.load_state_dict(g["state_dict"]))
    return r


def build_predictor(g):
    d = M.VideoMultiScaleMaskedTransformerDecoder_dvisPlus(
        64, True, num_classes=5, hidden_dim=64, num_queries=12, nheads=8, dim_feedforward=128, dec_layers=3,
        pre_norm=False, mask_dim=64, enforce_input_project=False, num_frames=2, num_reid_head_layers=3,
        reid_hidden_dim=64).eval()
    assert not any(d.load_state_dict(g["state_dict"]))
    return d


@torch.no_grad()
def test_tracker_cpu_matches_reference(golden):
    g = golden("tracker_small.pt")
    t = build_tracker(g)
    fe, fn, mf = g["frame_embeds"], g["frame_embeds_no_norm"], g["mask_features"]
    o1, i1 = t(fe[:, :, :2], mf[:, :2], resume=False, return_indices=True, frame_embeds_no_norm=fn[:, :, :2])
    o2, i2 = t(fe[:, :, 2:], mf[:, 2:], resume=True, return_indices=True, frame_embeds_no_norm=fn[:, :, 2:])
    for a, b in zip(i1 + i2, g["indices"]):
        assert np.array_equal(np.asarray(a), b.numpy())
    close(torch.cat([o1["pred_logits"], o2["pred_logits"]], 1), g["pred_logits"])
    close(torch.cat([o1["pred_masks"], o2["pred_masks"]], 2), g["pred_masks"])
    close(torch.cat([o1["pred_embds"], o2["pred_embds"]], 2), g["pred_embds"])
    close(torch.cat([o1["pred_references"], o2["pred_references"]], 2), g["pred_references"])
    o3 = t(fe[:, :, :2], None, resume=False, frame_embeds_no_norm=fn[:, :, :2], with_masks=False)
    assert o3["pred_masks"] is None
    close(o3["pred_embds"], g["pred_embds"][:, :, :2])


@torch.no_grad()
def test_refiner_cpu_matches_reference(golden):
    g = golden("refiner_small.pt")
    r = build_refiner(g)
    o = r(g["instance_embeds"], g["frame_embeds"], g["mask_features"])
    close(o["pred_logits"], g["pred_logits"])
    close(o["pred_masks"], g["pred_masks"])
    close(o["pred_embds"], g["pred_embds"])


@torch.no_grad()
def test_predictor_cpu_matches_reference(golden):
    g = golden("predictor_small.pt")
    d = build_predictor(g)
    out = d(g["multi_scale"], g["mask_features"])
    for k in ("pred_logits", "pred_masks", "pred_embds", "pred_embds_without_norm"):
        close(out[k], g[k])
    assert len(out["aux_outputs"]) == 3
    for a, b in zip(out["aux_outputs"], g["aux_masks"]):
        close(a["pred_masks"], b)


@torch.no_grad()
def test_mask_head_cpu_matches_reference(golden):
    g = golden("mask_head_small.pt")
    d = build_predictor(golden("predictor_small.pt"))
    cls, masks, am = d.forward_prediction_heads(g["output"], g["mask_features"], g["target_size"])
    close(cls, g["cls"])
    close(masks, g["masks"])
    assert (am != g["attn_mask"]).float().mean().item() < 1e-3


def test_msdeformattn_module_api(golden):
    g = golden("msdeformattn_module.pt")
    m = M.MSDeformAttn(d_model=64, n_levels=3, n_heads=8, n_points=4, ratio=0.5)
    assert not any(m.load_state_dict(g["state_dict"]))
    assert m.im2col_step == 128
    with pytest.raises(ValueError):
        M.MSDeformAttn(d_model=30, n_heads=8)
    with pytest.warns(UserWarning):
        M.MSDeformAttn(d_model=24, n_heads=8)
    with pytest.raises(RuntimeError, match="CPU"):          # no silent PyTorch fallback
        m(g["query"], g["ref"], g["src"], g["shapes"], torch.zeros(3, dtype=torch.long), None)
    # _reset_parameters reproduces the reference initialisation (OPS/modules/ms_deform_attn.py:66-80)
    torch.manual_seed(0)
    fresh = M.MSDeformAttn(d_model=64, n_levels=3, n_heads=8, n_points=4)
    assert fresh.sampling_offsets.weight.abs().max() == 0 and fresh.attention_weights.weight.abs().max() == 0
    b = fresh.sampling_offsets.bias.view(8, 3, 4, 2)
    assert torch.allclose(b[0, 0, :, 0], torch.tensor([1., 2., 3., 4.])) and torch.allclose(b[2, 1, :, 1], torch.tensor([1., 2., 3., 4.]))


@torch.no_grad()
def test_prenorm_predictor_variant_cpu(golden):
    """pre_norm=True blocks, enforce_input_project=True (1x1 convs), no ReID head: constructor variants of
    P/dvis_Plus/video_mask2former_transformer_decoder.py:177-230 that the DVIS configs do not use but the API offers."""
    g = golden("predictor_prenorm_small.pt")
    base = golden("predictor_small.pt")
    d = M.VideoMultiScaleMaskedTransformerDecoder_dvisPlus(
        64, True, num_classes=5, hidden_dim=64, num_queries=12, nheads=8, dim_feedforward=128, dec_layers=2, pre_norm=True,
        mask_dim=64, enforce_input_project=True, num_frames=2, num_reid_head_layers=0, reid_hidden_dim=64).eval()
    assert not any(d.load_state_dict(g["state_dict"]))
    out = d(base["multi_scale"], base["mask_features"])
    for k in ("pred_logits", "pred_masks", "pred_embds", "pred_embds_without_norm"):
        close(out[k], g[k])


@torch.no_grad()
def test_daq_refiner_with_local_conv_branch_cpu(golden):
    g = golden("daq_refiner_localattn_small.pt")
    rf = M.DAQTemporalRefiner(hidden_channel=64, feedforward_channel=128, num_head=8, decoder_layer_num=2, mask_dim=64,
                              class_num=5, windows=4, use_local_attn=True).eval()
    assert not any(rf.load_state_dict(g["state_dict"]))
    o = rf(g["instance_embeds"], None, g["frame_embeds"], g["mask_features"], None)
    close(o["pred_logits"], g["pred_logits"])
    close(o["pred_masks"], g["pred_masks"])
    close(o["pred_embds"], g["pred_embds"])
