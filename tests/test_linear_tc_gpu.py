"""csrc/linear_tc.cu (tcgen05 linear layers with the encoder's epilogues) and the head-major MSDA gather, on the B200, against
plain PyTorch fp32 references of the same ops (OPS/modules/ms_deform_attn.py:98-101,118; msdeformattn.py:118-119)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()


def _case(rows_shape, N, K, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(*rows_shape, K, generator=g, device="cuda").bfloat16()
    w = (torch.randn(N, K, generator=g, device="cuda") / K ** 0.5).bfloat16()
    b = torch.randn(N, generator=g, device="cuda")
    return x, w, b


@pytest.mark.parametrize("rows,N,K,relu", [(1000, 256, 256, False), (128, 32, 64, True), (129, 96, 512, False), (7, 256, 128, True),
                                           (148 * 128 * 2 + 5, 256, 256, False)])
def test_linear_tc_plain(rows, N, K, relu):
    from dvis_plus_b200 import ops
    x, w, b = _case((rows,), N, K, rows)
    y = ops.linear_tc(x, w, b, relu=relu)
    ref = F.linear(x.float(), w.float(), b)
    if relu:
        ref = ref.relu()
    assert y.shape == ref.shape and y.dtype == torch.bfloat16
    assert _rel(y, ref) < 6e-3                                            # one bf16 rounding of the result
    y0 = ops.linear_tc(x, w, None)
    assert _rel(y0, F.linear(x.float(), w.float())) < 6e-3


@pytest.mark.parametrize("B,S,masked", [(3, 777, True), (1, 128, False), (2, 19320, False), (5, 33, True)])
def test_linear_tc_heads_layout(B, S, masked):
    from dvis_plus_b200 import ops
    x, w, b = _case((B, S), 256, 256, S)
    mask = (torch.rand(B, S, device="cuda") < 0.2) if masked else None
    hm = ops.linear_tc_heads(x, w, b, row_mask=mask)
    ref = F.linear(x.float(), w.float(), b)
    if masked:
        ref = ref.masked_fill(mask[..., None], 0.0)
    ref = ref.view(B, S, 8, 32).permute(0, 2, 1, 3)
    assert hm.shape == (B, 8, S, 32) and hm.is_contiguous()
    assert _rel(hm, ref) < 6e-3
    if masked:
        assert (hm.permute(0, 2, 1, 3)[mask] == 0).all()


@pytest.mark.parametrize("rows,pos_rows,N,K", [(5000, 250, 256, 256), (128, 0, 256, 256), (777, 777, 128, 64), (2 * 19320, 19320, 256, 256),
                                               (300, 0, 32, 512)])
def test_linear_tc_add_layernorm(rows, pos_rows, N, K):
    from dvis_plus_b200 import ops
    x, w, b = _case((rows,), N, K, rows + 1)
    g = torch.Generator(device="cuda").manual_seed(5)
    res = torch.randn(rows, N, generator=g, device="cuda") * 2 + 0.5
    gamma, beta = torch.randn(N, generator=g, device="cuda"), torch.randn(N, generator=g, device="cuda")
    pos = torch.randn(pos_rows, N, generator=g, device="cuda") if pos_rows else None
    y32, ylp, ypos = ops.linear_tc_add_layernorm(x, w, b, res, gamma, beta, 1e-5, pos=pos)
    ref = F.layer_norm(res + F.linear(x.float(), w.float(), b), (N,), gamma, beta, 1e-5)
    assert _rel(y32, ref) < 1e-4                                           # tolerance: fp32 accumulation order only
    assert _rel(ylp, ref) < 6e-3
    if pos is not None:
        refp = (ref.view(-1, pos_rows, N) + pos[None]).view(rows, N)
        assert _rel(ypos, refp) < 6e-3
    else:
        assert ypos is None
    only_lp = ops.linear_tc_add_layernorm(x, w, b, res, gamma, beta, 1e-5, want_f32=False)
    assert only_lp[0] is None and torch.equal(only_lp[1], ylp)
    # the two-kernel form it replaces: library GEMM (bf16 output) + add_layernorm
    two = ops.add_layernorm(F.linear(x, w, b.bfloat16()), res, gamma, beta, 1e-5, lp_dtype=torch.bfloat16)[0]
    assert _rel(y32, two) < 2e-2


def _msda_inputs(N, shapes, seed, pdt=torch.bfloat16, scale=2.0):
    M, D, P = 8, 32, 4
    L = len(shapes)
    sh = torch.as_tensor(shapes, dtype=torch.long, device="cuda")
    lsi = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    S = int(sh.prod(1).sum())
    g = torch.Generator(device="cuda").manual_seed(seed)
    value = torch.randn(N, S, M, D, generator=g, device="cuda").bfloat16()
    fused = (torch.randn(N, S, M * L * P * 3, generator=g, device="cuda") *
             torch.cat([torch.full((M * L * P * 2,), scale), torch.ones(M * L * P)]).cuda()).to(pdt)
    ref_pts = torch.rand(N, S, L, 2, generator=g, device="cuda")
    return value, sh, lsi, fused[..., :M * L * P * 2], fused[..., M * L * P * 2:], ref_pts, M, L, P


@pytest.mark.parametrize("shapes,N,pdt", [(((12, 20), (6, 10), (3, 5)), 2, torch.bfloat16), (((5, 7), (3, 2)), 3, torch.float32),
                                          (((92, 160), (46, 80), (23, 40)), 2, torch.bfloat16)])
def test_msda_head_major_vs_oracle_and_token_major(shapes, N, pdt):
    from dvis_plus_b200 import ops
    from dvis_plus_b200.locality import tiled_item_order
    from oracle import c_oracle
    value, sh, lsi, offsets, logits, ref_pts, M, L, P = _msda_inputs(N, shapes, 3, pdt, scale=4.0)
    order = tiled_item_order(tuple(shapes), M, value.device) if len(shapes) == 3 else None
    hm = value.permute(0, 2, 1, 3).contiguous()
    out = ops.msda_fused_forward_hm(hm, sh, lsi, offsets, logits, ref_pts, L, P, item_order=order)
    tok = ops.msda_fused_forward(value, sh, lsi, offsets, logits, ref_pts, M, L, P, item_order=order)
    S = value.shape[1]
    off = offsets.float().view(N, S, M, L, P, 2).cpu()
    aw = logits.float().view(N, S, M, L * P).softmax(-1).view(N, S, M, L, P).cpu()
    norm = torch.stack([sh[:, 1], sh[:, 0]], -1).float().cpu()
    loc = ref_pts.cpu()[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    ref = c_oracle.msda_forward(value.float().cpu().numpy(), sh.cpu().numpy(), lsi.cpu().numpy(), loc.numpy(), aw.numpy())
    scale = max(1.0, np.abs(ref).max())
    assert np.abs(out.float().cpu().numpy() - ref).max() <= 1e-2 * scale           # the bf16 tier's tolerance
    assert (out.float() - tok.float()).abs().max().item() <= 1e-2 * scale


def test_pixel_decoder_tensor_core_path_matches_library_path():
    """Production width (conv_dim 256, 8 heads, 6 layers): forward_features with the tcgen05 projections + head-major gather
    against the same module on the library GEMMs + token-major gather."""
    from dvis_plus_b200 import modules as M
    from dvis_plus_b200.modules.pixel_decoder import ShapeSpec
    from dvis_plus_b200.modules.precision import precision
    torch.manual_seed(0)
    chans = dict(res2=192, res3=384, res4=768, res5=1536)
    strides = dict(res2=4, res3=8, res4=16, res5=32)
    pd = M.MSDeformAttnPixelDecoder({k: ShapeSpec(channels=chans[k], stride=strides[k]) for k in chans},
                                    transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=1024,
                                    transformer_enc_layers=6, conv_dim=256, mask_dim=256, norm="GN",
                                    transformer_in_features=["res3", "res4", "res5"], common_stride=4).eval().cuda()
    for layer in pd.transformer.encoder.layers:
        torch.nn.init.normal_(layer.self_attn.sampling_offsets.weight, std=0.02)
        torch.nn.init.normal_(layer.self_attn.attention_weights.weight, std=0.1)
        torch.nn.init.normal_(layer.self_attn.value_proj.bias, std=0.1)
        torch.nn.init.normal_(layer.self_attn.output_proj.bias, std=0.1)
    feats = {k: torch.randn(2, chans[k], 384 // strides[k], 640 // strides[k], device="cuda") for k in chans}
    outs = {}
    with torch.no_grad(), precision("bf16"):
        for tc in (True, False):
            for layer in pd.transformer.encoder.layers:
                layer.self_attn.use_tc_linear = tc
                layer.self_attn.fuse_output_norm = tc
            layer.self_attn.fuse_output_norm = tc
            mf, o0, ms = pd.forward_features(feats)
            outs[tc] = [mf.float(), o0.float()] + [m.float() for m in ms]
    for a, b in zip(outs[True], outs[False]):
        assert torch.isfinite(a).all()
        assert _rel(a, b) < 2e-2


def test_sm_partition_streams_run_kernels_on_disjoint_sm_sets():
    """dvis_plus_b200.partition: two green-context streams (driver API through cuda-python); kernels launched on either give the
    same results as on the default stream, and the partition sizes add up to the device."""
    from dvis_plus_b200 import ops
    from dvis_plus_b200.partition import sm_partition_streams
    big, small, info = sm_partition_streams(16)
    assert info["sms_small"] >= 16 and info["sms_big"] + info["sms_small"] <= torch.cuda.get_device_properties(0).multi_processor_count
    x, w, b = _case((1000,), 256, 256, 7)
    ref = ops.linear_tc(x, w, b)
    torch.cuda.synchronize()
    outs = []
    for st in (big, small):
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            outs.append(ops.linear_tc(x, w, b))
    torch.cuda.synchronize()
    assert torch.equal(outs[0], ref) and torch.equal(outs[1], ref)
