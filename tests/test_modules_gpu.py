"""GPU parity of the module-level drop-ins against golden outputs of the unmodified reference modules.

Tolerances (relative to the output scale): "fp32" mode 1e-3 (north star fp32 bound; mask logits run on TF32 tensor-core
operands there, 2e-3 downstream of a mask GEMM); "bf16" mode 1e-2 for one op, 1.5e-2 .. 2e-2 after 6..9 stacked layers (3e-2
where thresholded attention masks feed later layers).  The errors the B200 actually measures are in profiles/r2_test_errors.json
(tests/perf/report_test_errors.py): every asserted bound is ~2x the measured value; round 1 asserted 3e-2 .. 8e-2.
"""
import numpy as np
import pytest
import torch

from dvis_plus_b200 import _lib
from dvis_plus_b200 import modules as M
from dvis_plus_b200.modules.pixel_decoder import ShapeSpec
from dvis_plus_b200.modules.precision import precision
from test_modules_cpu import build_predictor, build_refiner, build_tracker

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    assert a.shape == b.shape, (a.shape, b.shape)
    return (a.double().cpu() - b.double()).abs().max().item() / max(1e-6, b.abs().max().item())


def cuda(x):
    if isinstance(x, dict):
        return {k: cuda(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [cuda(v) for v in x]
    return x.cuda()


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
@torch.no_grad()
def test_msdeformattn_module(golden, mode, tol):
    g = golden("msdeformattn_module.pt")
    m = M.MSDeformAttn(d_model=64, n_levels=3, n_heads=8, n_points=4).cuda().eval()
    m.load_state_dict(g["state_dict"])
    sh = g["shapes"].cuda()
    lsi = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    calls = _lib.launch_count
    with precision(mode):
        out = m(g["query"].cuda(), g["ref"].cuda(), g["src"].cuda(), sh, lsi, None)
        out_pad = m(g["query"].cuda(), g["ref"].cuda(), g["src"].cuda(), sh, lsi, g["padding_mask"].cuda())
        out_box = m(g["query"].cuda(), g["ref4"].cuda(), g["src"].cuda(), sh, lsi, None)
    assert _lib.launch_count >= calls + 3, "the CUDA kernel did not run"
    assert rel_err(out, g["out"]) < tol and rel_err(out_pad, g["out_pad"]) < tol and rel_err(out_box, g["out_box"]) < tol


def test_msdeformattn_autograd_path_matches_fused(golden):
    g = golden("msdeformattn_module.pt")
    m = M.MSDeformAttn(d_model=64, n_levels=3, n_heads=8, n_points=4).cuda()
    m.load_state_dict(g["state_dict"])
    sh = g["shapes"].cuda()
    lsi = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    q = g["query"].cuda().requires_grad_()
    out = m(q, g["ref"].cuda(), g["src"].cuda(), sh, lsi, None)     # reference formulation around the plain op
    assert out.requires_grad
    assert rel_err(out.detach(), g["out"]) < 1e-4


def _pixel_decoder(g):
    chans = dict(res2=8, res3=16, res4=24, res5=32)
    strides = dict(res2=4, res3=8, res4=16, res5=32)
    pd = M.MSDeformAttnPixelDecoder({k: ShapeSpec(channels=chans[k], stride=strides[k]) for k in chans},
                                    transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=128,
                                    transformer_enc_layers=2, conv_dim=64, mask_dim=64, norm="GN",
                                    transformer_in_features=["res3", "res4", "res5"], common_stride=4).cuda().eval()
    pd.load_state_dict(g["state_dict"])
    return pd


# conv_dim = 64 in the fixture is not a multiple of 128, which the fused LayerNorm kernel requires: the 256-wide
# production geometry is covered by test_pixel_decoder_production_width below (vs the oracle port)
@torch.no_grad()
def test_pixel_decoder_production_width():
    from oracle import torch_port as tp
    torch.manual_seed(0)
    chans = dict(res2=16, res3=24, res4=32, res5=48)
    strides = dict(res2=4, res3=8, res4=16, res5=32)
    pd = M.MSDeformAttnPixelDecoder({k: ShapeSpec(channels=chans[k], stride=strides[k]) for k in chans},
                                    transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=1024,
                                    transformer_enc_layers=3, conv_dim=256, mask_dim=256, norm="GN",
                                    transformer_in_features=["res3", "res4", "res5"], common_stride=4).eval()
    for layer in pd.transformer.encoder.layers:
        torch.nn.init.normal_(layer.self_attn.sampling_offsets.weight, std=0.02)
        torch.nn.init.normal_(layer.self_attn.attention_weights.weight, std=0.1)
    feats = {k: torch.randn(2, chans[k], 96 // strides[k], 160 // strides[k]) for k in chans}
    sd = {k: v.detach() for k, v in pd.state_dict().items()}
    ref_mf, ref_o0, ref_ms = tp.pixel_decoder_forward_features(sd, feats, num_layers=3)
    pd = pd.cuda()
    for mode, tol in (("fp32", 1e-3), ("bf16", 2e-2)):
        calls = _lib.launch_count
        with precision(mode):
            mf, o0, ms = pd.forward_features(cuda(feats))
        assert _lib.launch_count > calls
        assert mf.shape == ref_mf.shape and mf.is_contiguous(memory_format=torch.channels_last)
        assert rel_err(mf.float(), ref_mf) < tol, (mode, rel_err(mf.float(), ref_mf))
        assert rel_err(o0.float(), ref_o0) < tol
        for a, b in zip(ms, ref_ms):
            assert rel_err(a.float(), b) < tol


@torch.no_grad()
def test_pixel_decoder_golden_small_width(golden):
    """The reference's own fixture (conv_dim 64): runs the autograd-style module path on the plain MSDA op."""
    g = golden("pixel_decoder_small.pt")
    pd = _pixel_decoder(g)
    with torch.enable_grad():
        mf, o0, ms = pd.forward_features(cuda(g["features"]))
    assert rel_err(mf.detach(), g["mask_features"]) < 1e-3
    assert rel_err(o0.detach(), g["out0"]) < 1e-3
    for a, b in zip(ms, g["multi_scale"]):
        assert rel_err(a.detach(), b) < 1e-3


@pytest.mark.parametrize("materialize", [True, False])
@torch.no_grad()
def test_predictor_golden(golden, materialize):
    g = golden("predictor_small.pt")
    d = build_predictor(g).cuda()
    d.materialize_aux_masks = materialize
    with precision("fp32"):
        out = d(cuda(g["multi_scale"]), g["mask_features"].cuda())
    # fp32 tier: every GEMM in fp32, the mask head on TF32 tensor-core operands (csrc/mask_gemm.cu kTf32).  The production
    # formulation (materialize=False: attention masks from E @ resize(F) on each level's grid) reproduces the reference to
    # fp32 round-off everywhere except the TF32 mask logits themselves (measured 4e-7 / 1.05e-3, tests/perf/fp32_tier_errors.py;
    # round 1 with bf16 operands: 3e-2 .. 5e-2).  With materialize=True the attention masks are thresholded RESIZED full-size
    # logits like the reference's (decoder.py:367-372): with random-init weights many of those sit near 0, TF32's 1e-3 flips a
    # few signs and the discrete change propagates through 3 layers -- 5e-2 of the output scale bounds that.
    tols = dict(pred_logits=1e-4, pred_masks=2e-3, pred_embds=1e-4, pred_embds_without_norm=1e-4) if not materialize else \
        dict(pred_logits=5e-2, pred_masks=5e-2, pred_embds=5e-2, pred_embds_without_norm=5e-2)
    for k, tol in tols.items():
        assert rel_err(out[k], g[k]) < tol, (k, rel_err(out[k], g[k]))
    if materialize:
        assert len(out["aux_outputs"]) == 3
        for a, b in zip(out["aux_outputs"], g["aux_masks"]):
            assert rel_err(a["pred_masks"], b) < 5e-2


@torch.no_grad()
def test_mask_head_golden_and_lowres_equivalence(golden):
    g = golden("mask_head_small.pt")
    d = build_predictor(golden("predictor_small.pt")).cuda()
    cls, masks, am = d.forward_prediction_heads(g["output"].cuda(), g["mask_features"].cuda(), g["target_size"])
    assert rel_err(cls, g["cls"]) < 1e-2 and rel_err(masks, g["masks"]) < 1e-2
    flips = (am.cpu() != g["attn_mask"])
    assert flips.float().mean().item() < 2e-2
    # the low-resolution formulation agrees with the reference mask except where the resized logit is ~0
    import torch.nn.functional as F
    lvl = F.interpolate(g["mask_features"].cuda(), size=g["target_size"], mode="bilinear", align_corners=False)
    low = d._heads_lowres(g["output"].cuda(), lvl.to(torch.bfloat16).contiguous(memory_format=torch.channels_last))
    ref_low = g["attn_mask"].reshape(2, 8, 12, -1)[:, 0]
    ref_logit = F.interpolate(g["masks"], size=g["target_size"], mode="bilinear", align_corners=False).flatten(2)
    bad = (low.cpu() != ref_low) & (ref_logit.abs() > 2e-2 * ref_logit.abs().max())
    assert not bad.any()


@pytest.mark.parametrize("mode,tol", [("fp32", 2e-3), ("bf16", 1.5e-2)])
@torch.no_grad()
def test_tracker_golden(golden, mode, tol):
    g = golden("tracker_small.pt")
    t = build_tracker(g).cuda()
    fe, fn, mf = g["frame_embeds"].cuda(), g["frame_embeds_no_norm"].cuda(), g["mask_features"].cuda()
    with precision(mode):
        o1, i1 = t(fe[:, :, :2], mf[:, :2], resume=False, return_indices=True, frame_embeds_no_norm=fn[:, :, :2])
        o2, i2 = t(fe[:, :, 2:], mf[:, 2:], resume=True, return_indices=True, frame_embeds_no_norm=fn[:, :, 2:])
    for a, b in zip(i1 + i2, g["indices"]):
        assert np.array_equal(np.asarray(a), b.numpy())
    emb_tol = 1e-3 if mode == "fp32" else tol
    assert rel_err(torch.cat([o1["pred_embds"], o2["pred_embds"]], 2), g["pred_embds"]) < emb_tol
    assert rel_err(torch.cat([o1["pred_logits"], o2["pred_logits"]], 1), g["pred_logits"]) < emb_tol
    assert rel_err(torch.cat([o1["pred_masks"], o2["pred_masks"]], 2), g["pred_masks"]) < tol   # folded conv + bf16 GEMM


@pytest.mark.parametrize("mode,tol", [("fp32", 2e-3), ("bf16", 1.5e-2)])
@torch.no_grad()
def test_refiner_golden(golden, mode, tol):
    g = golden("refiner_small.pt")
    r = build_refiner(g).cuda()
    with precision(mode):
        o = r(g["instance_embeds"].cuda(), g["frame_embeds"].cuda(), g["mask_features"].cuda())
    emb_tol = 1e-3 if mode == "fp32" else tol      # fp32 mode: cuDNN convs run TF32 by default, like the reference
    assert rel_err(o["pred_embds"], g["pred_embds"]) < emb_tol
    assert rel_err(o["pred_logits"], g["pred_logits"]) < emb_tol
    assert rel_err(o["pred_masks"], g["pred_masks"]) < tol


@torch.no_grad()
@torch.no_grad()
def test_refiner_row_layout_equals_module_path(golden):
    """TemporalRefiner._refine_rows (default on the bf16 inference path) against the module-by-module path on the device."""
    g = golden("refiner_small.pt")
    r = build_refiner(g).cuda()
    outs = {}
    for rows in (True, False):
        r.use_row_layout = rows
        with precision("bf16"):
            outs[rows] = r(g["instance_embeds"].cuda(), g["frame_embeds"].cuda(), g["mask_features"].cuda())
    for k in ("pred_embds", "pred_logits", "pred_masks"):
        assert rel_err(outs[True][k], outs[False][k].float().cpu()) < 2e-2, k
        assert rel_err(outs[True][k], g[k]) < 1.5e-2, k


def test_add_layernorm_kernel():
    from dvis_plus_b200 import ops
    torch.manual_seed(0)
    for C in (128, 256, 512, 1024):
        x = torch.randn(37, 5, C, device="cuda")
        r = torch.randn(37, 5, C, device="cuda")
        w, b = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
        pos = torch.randn(5, C, device="cuda")
        ref = torch.nn.functional.layer_norm(x + r, (C,), w, b, 1e-5)
        y32, ylp, yq = ops.add_layernorm(x, r, w, b, lp_dtype=torch.bfloat16, pos=pos)
        assert (y32 - ref).abs().max() < 1e-5 * max(1.0, ref.abs().max().item())
        assert (ylp.float() - ref).abs().max() < 1e-2 * ref.abs().max()
        assert (yq.float() - (ref + pos)).abs().max() < 1e-2 * (ref + pos).abs().max()
        y32b, _, _ = ops.add_layernorm(x.bfloat16(), r, w, b)
        refb = torch.nn.functional.layer_norm(x.bfloat16().float() + r, (C,), w, b, 1e-5)
        assert (y32b - refb).abs().max() < 1e-5 * max(1.0, refb.abs().max().item())
        y32n, _, _ = ops.add_layernorm(x, None, w, b)
        assert (y32n - torch.nn.functional.layer_norm(x, (C,), w, b, 1e-5)).abs().max() < 1e-5 * 10


@torch.no_grad()
def test_groupnorm_nhwc_kernel():
    from dvis_plus_b200 import ops
    import torch.nn.functional as F
    torch.manual_seed(0)
    N, H, W, C, G = 3, 12, 20, 256, 32
    x = torch.randn(N, H * W, C, device="cuda") * 2 + 0.5
    w, b = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
    ref = F.group_norm(x.transpose(1, 2).reshape(N, C, H, W), G, w, b, 1e-5)               # NCHW reference
    ref_tok = ref.flatten(2).transpose(1, 2)
    # plain, into a slice of a bigger token buffer, with pos
    buf32 = torch.zeros(N, H * W + 7, C, device="cuda")
    buf_lp = torch.zeros(N, H * W + 7, C, device="cuda", dtype=torch.bfloat16)
    buf_q = torch.zeros_like(buf_lp)
    pos = torch.randn(H * W, C, device="cuda")
    ops.groupnorm_nhwc(x, G, w, b, pos=pos, out_f32=buf32[:, 7:], out_lp=buf_lp[:, 7:], out_lp_pos=buf_q[:, 7:])
    assert (buf32[:, 7:] - ref_tok).abs().max() < 1e-4 * ref_tok.abs().max()
    assert buf32[:, :7].abs().max() == 0
    assert (buf_lp[:, 7:].float() - ref_tok).abs().max() < 1e-2 * ref_tok.abs().max()
    assert (buf_q[:, 7:].float() - (ref_tok + pos)).abs().max() < 1e-2 * (ref_tok + pos).abs().max()
    # upsample-add + relu, bf16 input
    up = torch.randn(N, (H // 2) * (W // 2), C, device="cuda")
    up_ref = F.interpolate(up.transpose(1, 2).reshape(N, C, H // 2, W // 2), size=(H, W), mode="bilinear", align_corners=False)
    xb = x.bfloat16()
    ref2 = F.relu(F.group_norm(xb.float().transpose(1, 2).reshape(N, C, H, W), G, w, b, 1e-5) + up_ref).flatten(2).transpose(1, 2)
    out = torch.empty(N, H * W, C, device="cuda")
    ops.groupnorm_nhwc(xb, G, w, b, relu=True, up=up, up_hw=(H // 2, W // 2), hw=(H, W), out_f32=out)
    assert (out - ref2).abs().max() < 1e-4 * ref2.abs().max()


@torch.no_grad()
def test_predictor_prenorm_variant_golden(golden):
    """pre_norm=True + enforce_input_project + no ReID head: runs the generic loop (the fused path is post-norm only)."""
    g, base = golden("predictor_prenorm_small.pt"), golden("predictor_small.pt")
    d = M.VideoMultiScaleMaskedTransformerDecoder_dvisPlus(
        64, True, num_classes=5, hidden_dim=64, num_queries=12, nheads=8, dim_feedforward=128, dec_layers=2, pre_norm=True,
        mask_dim=64, enforce_input_project=True, num_frames=2, num_reid_head_layers=0, reid_hidden_dim=64).eval()
    d.load_state_dict(g["state_dict"])
    d = d.cuda()
    calls = _lib.launch_count
    with precision("fp32"):
        out = d(cuda(base["multi_scale"]), base["mask_features"].cuda())
    assert _lib.launch_count > calls
    for k in ("pred_logits", "pred_masks", "pred_embds", "pred_embds_without_norm"):
        assert rel_err(out[k], g[k]) < 3e-2, (k, rel_err(out[k], g[k]))     # thresholded bf16 logits, see above


@torch.no_grad()
def test_predictor_fast_path_production_width():
    """hidden_dim 256: the batch-first inference path against the oracle port (CPU, fp32)."""
    from oracle import torch_port as tp
    torch.manual_seed(0)
    d = M.VideoMultiScaleMaskedTransformerDecoder_dvisPlus(
        256, True, num_classes=7, hidden_dim=256, num_queries=20, nheads=8, dim_feedforward=512, dec_layers=4,
        pre_norm=False, mask_dim=256, enforce_input_project=False, num_frames=1, num_reid_head_layers=3,
        reid_hidden_dim=256).eval()
    ms = [torch.randn(2, 256, 4, 6), torch.randn(2, 256, 8, 12), torch.randn(2, 256, 16, 24)]
    mf = torch.randn(2, 256, 32, 48)
    sd = {k: v.detach() for k, v in d.state_dict().items()}
    ref = tp.predictor_forward(sd, ms, mf, num_layers=4)
    d = d.cuda()
    calls = _lib.launch_count
    with precision("fp32"):
        out = d(cuda(ms), mf.cuda())
    assert _lib.launch_count > calls
    for k in ("pred_logits", "pred_masks", "pred_embds", "pred_embds_without_norm"):
        assert rel_err(out[k], ref[k]) < 3e-2, (k, rel_err(out[k], ref[k]))     # thresholded bf16 logits, see above
    with precision("bf16"):
        out = d(cuda(ms), mf.cuda())
    for k in ("pred_logits", "pred_masks", "pred_embds"):
        assert rel_err(out[k], ref[k]) < 3e-2, (k, rel_err(out[k], ref[k]))


def test_lap_chain_matches_scipy():
    """GPU Hungarian (one CTA per frame) vs scipy.optimize.linear_sum_assignment, incl. the index chain of the tracker."""
    from scipy.optimize import linear_sum_assignment
    from dvis_plus_b200 import ops
    torch.manual_seed(0)
    for T, n in ((1, 1), (3, 7), (5, 33), (16, 200), (2, 300)):
        emb = torch.randn(T + 1, n, 64)
        emb = emb / emb.norm(dim=-1, keepdim=True)
        emb[1:] = 0.7 * emb[:-1][:, torch.randperm(n)] + 0.3 * emb[1:]            # consecutive frames are related
        cost = 1 - torch.bmm(emb[:-1], emb[1:].transpose(1, 2))                   # rows = previous frame, cols = current
        cost[0, 0, 0] = float("nan")                                              # NaN -> 0 like noiser.py:52
        sigma, idx = ops.lap_chain(cost.cuda())
        c = torch.where(torch.isnan(cost), torch.zeros_like(cost), cost).numpy()
        prev = None
        for t in range(T):
            ref = linear_sum_assignment(c[t])[1]
            assert np.array_equal(sigma[t].cpu().numpy(), ref), (T, n, t)
            prev = ref if prev is None else ref[prev]
            assert np.array_equal(idx[t].cpu().numpy(), prev)
    # adversarial: random uniform costs (long augmenting paths)
    cost = torch.rand(4, 150, 150)
    sigma, _ = ops.lap_chain(cost.cuda())
    for t in range(4):
        ref = linear_sum_assignment(cost[t].numpy())[1]
        got = sigma[t].cpu().numpy()
        assert sorted(got.tolist()) == list(range(150))
        assert abs(cost[t].numpy()[np.arange(150), got].sum() - cost[t].numpy()[np.arange(150), ref].sum()) < 1e-4


@torch.no_grad()
def test_resize_and_attn_bias_kernels():
    from dvis_plus_b200 import ops
    import torch.nn.functional as F
    torch.manual_seed(0)
    x = torch.randn(3, 64, 23, 40, device="cuda").to(torch.bfloat16, memory_format=torch.channels_last)
    for size in ((12, 20), (6, 10), (46, 80), (5, 7)):
        ref = F.interpolate(x.float(), size=size, mode="bilinear", align_corners=False)
        out = ops.resize_bilinear_nhwc(x, size)
        assert out.shape == ref.shape and out.is_contiguous(memory_format=torch.channels_last)
        assert (out.float() - ref).abs().max() < 1e-2 * ref.abs().max()
    logits = torch.randn(4, 9, 333, device="cuda")
    logits[1, 3] = -logits[1, 3].abs() - 0.1          # a fully masked row -> must become all zeros (decoder.py:297)
    bias = ops.attn_bias_from_logits(logits, torch.float32)
    m = logits.sigmoid() < 0.5
    m[m.all(-1)] = False
    ref = torch.zeros_like(logits).masked_fill_(m, float("-inf"))
    assert torch.equal(bias, ref) and bias[1, 3].abs().max() == 0


@pytest.mark.parametrize("B,Lq,Lk,H,dh", [(1, 200, 200, 8, 64), (6, 200, 200, 8, 64), (3, 16, 16, 8, 64), (2, 37, 53, 4, 32),
                                         (16, 200, 200, 8, 32), (1, 1, 300, 8, 64)])
@torch.no_grad()
def test_flash_attn_kernel_tracker_layouts(B, Lq, Lk, H, dh):
    """dvis_flash_attn vs an fp32 softmax(QK^T)V on the same bf16 inputs, with q / k / v as strided slices of packed
    projections (the layouts the tracker uses)."""
    from dvis_plus_b200 import ops
    import torch.nn.functional as F
    torch.manual_seed(B * 1000 + Lq)
    C = H * dh
    if Lq == Lk:
        qkv = torch.randn(B, Lq, 3, H, dh, device="cuda").bfloat16()                 # packed self-attention projection
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    else:
        q = torch.randn(Lq, B, H, dh, device="cuda").bfloat16().permute(1, 0, 2, 3)  # batch-strided query
        kv = torch.randn(B, Lk, 2, H, dh, device="cuda").bfloat16()
        k, v = kv[:, :, 0], kv[:, :, 1]
    scale = dh ** -0.5
    out = ops.flash_attn(q, k, v, scale)
    ref = F.scaled_dot_product_attention(q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2), scale=scale)
    ref = ref.transpose(1, 2).reshape(B, Lq, C)
    assert out.shape == (B, Lq, C)
    assert (out.float() - ref).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())
