"""Pin the oracle (oracle/) against golden outputs of the unmodified reference (tests/golden/*.pt).

CPU only.  Tolerances: fp64 cases use default allclose like OPS/test.py:43; fp32 module-level cases use
1e-4 abs/rel (same algorithm, different op order).
"""
import numpy as np
import torch

from oracle import c_oracle, torch_port as tp


def lsi_of(shapes):
    return torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))


def close(a, b, tol=1e-4):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a.double() - b.double()).abs().max().item()
    scale = max(1.0, b.double().abs().max().item())
    assert err <= tol * scale, f"max abs err {err:.3e} (scale {scale:.3e})"


def test_c_oracle_optest_shapes(golden):
    g = golden("msda_optest.pt")
    out = c_oracle.msda_forward(g["value64"].numpy(), g["shapes"].numpy(), g["lsi"].numpy(), g["loc64"].numpy(), g["attn64"].numpy())
    assert torch.allclose(torch.from_numpy(out), g["out64"])           # OPS/test.py:43
    out = c_oracle.msda_forward(g["value32"].numpy(), g["shapes"].numpy(), g["lsi"].numpy(), g["loc32"].numpy(), g["attn32"].numpy())
    assert torch.allclose(torch.from_numpy(out), g["out32"], rtol=1e-2, atol=1e-3)   # OPS/test.py:59
    assert np.abs(out - g["out32"].numpy()).max() < 1e-8


def test_c_oracle_forward_backward_small(golden):
    g = golden("msda_small.pt")
    args = [g[k].numpy() for k in ("value", "shapes", "lsi", "loc", "attn")]
    out = c_oracle.msda_forward(*args)
    assert np.allclose(out, g["out"].numpy(), rtol=1e-10, atol=1e-12)
    gv, gl, ga = c_oracle.msda_backward(*args, g["grad_out"].numpy())
    assert np.allclose(gv, g["grad_value"].numpy(), rtol=1e-9, atol=1e-11)
    assert np.allclose(gl, g["grad_loc"].numpy(), rtol=1e-9, atol=1e-11)
    assert np.allclose(ga, g["grad_attn"].numpy(), rtol=1e-9, atol=1e-11)


def test_torch_port_msda_core(golden):
    g = golden("msda_small.pt")
    out = tp.msda_core(g["value"], g["shapes"].tolist(), g["loc"], g["attn"])
    assert torch.allclose(out, g["out"])


def test_oracles_config1(golden):
    """BASELINE config 1: 256x256, 1 level, 8 heads, 4 points, Q=100; inputs regenerated from the seed."""
    g = golden("msda_cfg1_out.pt")
    torch.manual_seed(g["seed"])
    value = torch.rand(1, 65536, 8, 32) * 0.01
    loc = torch.rand(1, 100, 8, 1, 4, 2)
    attn = torch.rand(1, 100, 8, 1, 4) + 1e-5
    attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    sh = torch.as_tensor([(256, 256)])
    out = c_oracle.msda_forward(value.numpy(), sh.numpy(), np.zeros(1, np.int64), loc.numpy(), attn.numpy())
    assert np.abs(out - g["out"].numpy()).max() < 1e-7
    close(tp.msda_core(value, [(256, 256)], loc, attn), g["out"], 1e-6)


def test_torch_port_msdeformattn_module(golden):
    g = golden("msdeformattn_module.pt")
    sd = {"m." + k: v for k, v in g["state_dict"].items()}
    sh = g["shapes"].tolist()
    close(tp.ms_deform_attn(sd, "m", g["query"], g["ref"], g["src"], sh, 8, 3, 4), g["out"])
    close(tp.ms_deform_attn(sd, "m", g["query"], g["ref"], g["src"], sh, 8, 3, 4, input_padding_mask=g["padding_mask"]), g["out_pad"])
    close(tp.ms_deform_attn(sd, "m", g["query"], g["ref4"], g["src"], sh, 8, 3, 4), g["out_box"])


def test_torch_port_pixel_decoder(golden):
    g = golden("pixel_decoder_small.pt")
    mf, o0, ms = tp.pixel_decoder_forward_features(g["state_dict"], g["features"], num_layers=2)
    close(mf, g["mask_features"])
    close(o0, g["out0"])
    for a, b in zip(ms, g["multi_scale"]):
        close(a, b)


def test_torch_port_mask_head(golden):
    g = golden("mask_head_small.pt")
    sd = golden("predictor_small.pt")["state_dict"]
    cls, masks, am = tp.prediction_heads(sd, "", g["output"], g["mask_features"], g["target_size"], 8)
    close(cls, g["cls"])
    close(masks, g["masks"])
    assert (am != g["attn_mask"]).float().mean().item() < 1e-3
    close(c_oracle.mask_logits(tp.mlp(sd, "mask_embed", tp.layer_norm(sd, "decoder_norm", g["output"]).transpose(0, 1)).numpy(),
                               g["mask_features"].numpy()), g["masks"])


def test_torch_port_predictor(golden):
    g = golden("predictor_small.pt")
    out = tp.predictor_forward(g["state_dict"], g["multi_scale"], g["mask_features"], num_layers=3)
    for k in ("pred_logits", "pred_masks", "pred_embds", "pred_embds_without_norm"):
        close(out[k], g[k], 2e-4)
    for a, b in zip(out["all_masks"][:-1], g["aux_masks"]):
        close(a.permute(1, 0, 2, 3)[None], b, 2e-4)


def test_torch_port_tracker(golden):
    g = golden("tracker_small.pt")
    sd = g["state_dict"]
    fe, fn, mf = g["frame_embeds"], g["frame_embeds_no_norm"], g["mask_features"]
    o1 = tp.tracker_forward(sd, fe[:, :, :2], mf[:, :2], fn[:, :, :2], num_layers=2)
    o2 = tp.tracker_forward(sd, fe[:, :, 2:], mf[:, 2:], fn[:, :, 2:], num_layers=2, state=o1["state"])
    for a, b in zip(o1["indices"] + o2["indices"], g["indices"]):
        assert np.array_equal(np.asarray(a), b.numpy())
    close(torch.cat([o1["pred_logits"], o2["pred_logits"]], 1), g["pred_logits"])
    close(torch.cat([o1["pred_masks"], o2["pred_masks"]], 2), g["pred_masks"])
    close(torch.cat([o1["pred_embds"], o2["pred_embds"]], 2), g["pred_embds"])
    close(torch.cat([o1["pred_references"], o2["pred_references"]], 2), g["pred_references"])


def test_torch_port_refiner(golden):
    g = golden("refiner_small.pt")
    o = tp.refiner_forward(g["state_dict"], g["instance_embeds"], g["frame_embeds"], g["mask_features"], num_layers=2)
    close(o["pred_logits"], g["pred_logits"])
    close(o["pred_masks"], g["pred_masks"])
    close(o["pred_embds"], g["pred_embds"])
