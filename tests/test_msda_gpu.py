"""GPU parity of the MSDA forward kernels against the oracle (C restatement + golden vectors).

Tolerances: fp64 default allclose (OPS/test.py:43); fp32 max-abs <= 1e-3 of the north star, in practice we
assert the much tighter 2e-5 * scale; bf16-value variants 1e-2 * scale.
"""
import numpy as np
import pytest
import torch

from oracle import c_oracle

pytestmark = pytest.mark.gpu


def lsi_of(shapes):
    return torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))


def _run(value, shapes, loc, attn, order=None):
    from dvis_plus_b200 import ops
    dev = "cuda"
    return ops.ms_deform_attn_forward(value.to(dev), shapes.to(dev), lsi_of(shapes).to(dev), loc.to(dev), attn.to(dev),
                                      128, item_order=order).cpu()


def _oracle(value, shapes, loc, attn):
    return torch.from_numpy(c_oracle.msda_forward(value.numpy(), shapes.numpy(), lsi_of(shapes).numpy(), loc.numpy(), attn.numpy()))


def test_optest_shapes_golden(golden):
    g = golden("msda_optest.pt")
    out = _run(g["value64"], g["shapes"], g["loc64"], g["attn64"])
    assert torch.allclose(out, g["out64"])
    out = _run(g["value32"], g["shapes"], g["loc32"], g["attn32"])
    assert torch.allclose(out, g["out32"], rtol=1e-2, atol=1e-3)
    assert (out - g["out32"]).abs().max() < 1e-8


def test_small_golden_fp64_and_fp32(golden):
    g = golden("msda_small.pt")
    out = _run(g["value"], g["shapes"], g["loc"], g["attn"])
    assert torch.allclose(out, g["out"])
    out32 = _run(g["value"].float(), g["shapes"], g["loc"].float(), g["attn"].float())
    assert (out32.double() - g["out"]).abs().max() < 2e-5 * g["out"].abs().max()


def test_config1_golden(golden):
    g = golden("msda_cfg1_out.pt")
    torch.manual_seed(g["seed"])
    value = torch.rand(1, 65536, 8, 32) * 0.01
    loc = torch.rand(1, 100, 8, 1, 4, 2)
    attn = torch.rand(1, 100, 8, 1, 4) + 1e-5
    attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    out = _run(value, torch.as_tensor([(256, 256)]), loc, attn)
    assert (out - g["out"]).abs().max() < 1e-7  # north star: <= 1e-3


@pytest.mark.parametrize("D,P,L", [(32, 4, 3), (64, 4, 3), (64, 4, 1), (16, 4, 2), (8, 2, 2), (128, 4, 1),
                                   (32, 3, 3), (30, 4, 2), (71, 2, 2), (1, 1, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_random_vs_oracle(D, P, L, dtype):
    torch.manual_seed(D * 100 + P * 10 + L)
    shapes = torch.as_tensor([(13, 17), (7, 9), (4, 5)][:L])
    S = int(shapes.prod(1).sum())
    N, M, Lq = 3, 4, 37
    value = torch.randn(N, S, M, D, dtype=dtype)
    loc = (torch.rand(N, Lq, M, L, P, 2) * 1.5 - 0.25).to(dtype)   # ~1/3 of the points outside the map
    attn = torch.rand(N, Lq, M, L, P).flatten(-2).softmax(-1).view(N, Lq, M, L, P).to(dtype)
    out = _run(value, shapes, loc, attn)
    ref = _oracle(value, shapes, loc, attn)
    tol = 1e-12 if dtype == torch.float64 else 2e-5
    assert (out - ref).abs().max() <= tol * max(1.0, ref.abs().max())


def test_edge_locations_exact_borders():
    """Samples exactly on pixel centres, borders, -1 / H boundaries, and NaN locations (skipped like the reference)."""
    H, W = 5, 6
    shapes = torch.as_tensor([(H, W)])
    value = torch.randn(1, H * W, 1, 32)
    xs = torch.tensor([0.5 / W, 0.0, 1.0, -0.5 / W, (W + 0.5) / W, 1.5 / W, float("nan"), 0.999999, 1e-7, 2.0, -1.0, 0.5])
    ys = torch.tensor([0.5 / H, 0.0, 1.0, 0.5, 0.5, (H + 0.5) / H, 0.5, 0.999999, 1e-7, 0.5, 0.5, float("nan")])
    Lq = xs.numel()
    loc = torch.stack([xs, ys], -1).view(1, Lq, 1, 1, 1, 2)
    attn = torch.ones(1, Lq, 1, 1, 1)
    out = _run(value, shapes, loc, attn)
    ref = _oracle(value, shapes, loc, attn)
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max() < 1e-5


def test_item_order_is_only_a_schedule():
    from dvis_plus_b200.locality import tiled_item_order
    torch.manual_seed(1)
    shapes = torch.as_tensor([(23, 40), (12, 20), (6, 10)])
    S = int(shapes.prod(1).sum())
    M, D, L, P = 8, 32, 3, 4
    value = torch.randn(2, S, M, D)
    loc = torch.rand(2, S, M, L, P, 2)
    attn = torch.rand(2, S, M, L, P).flatten(-2).softmax(-1).view(2, S, M, L, P)
    order = tiled_item_order(shapes.tolist(), M, "cuda")
    assert sorted(order.cpu().tolist()) == list(range(S * M))
    a = _run(value, shapes, loc, attn)
    b = _run(value, shapes, loc, attn, order=order)
    assert torch.equal(a, b)
    assert (a - _oracle(value, shapes, loc, attn)).abs().max() < 2e-5 * 4


def test_argument_errors_like_reference():
    from dvis_plus_b200 import ops
    shapes = torch.as_tensor([(4, 4)])
    v = torch.randn(3, 16, 2, 8)
    loc = torch.rand(3, 5, 2, 1, 2, 2)
    attn = torch.rand(3, 5, 2, 1, 2)
    with pytest.raises(RuntimeError, match="CPU"):
        ops.ms_deform_attn_forward(v, shapes, lsi_of(shapes), loc, attn, 128)
    c = lambda t: t.cuda()
    with pytest.raises(RuntimeError, match="contiguous"):
        ops.ms_deform_attn_forward(c(v).transpose(2, 3), c(shapes), c(lsi_of(shapes)), c(loc), c(attn), 128)
    with pytest.raises(RuntimeError, match="im2col_step"):
        ops.ms_deform_attn_forward(c(v), c(shapes), c(lsi_of(shapes)), c(loc), c(attn), 2)
    with pytest.raises(RuntimeError, match="not implemented"):
        ops.ms_deform_attn_forward(c(v).half(), c(shapes), c(lsi_of(shapes)), c(loc).half(), c(attn).half(), 128)


def test_full_size_720p_properties():
    """BASELINE 720p size (S=Lq=19320, L=3, M=8, D=32): linearity in value and in the attention weights, and
    agreement with the oracle on a random subset of queries."""
    from dvis_plus_b200.locality import tiled_item_order
    torch.manual_seed(7)
    shapes = torch.as_tensor([(92, 160), (46, 80), (23, 40)])
    S = int(shapes.prod(1).sum())
    M, D, L, P = 8, 32, 3, 4
    v1, v2 = torch.randn(1, S, M, D), torch.randn(1, S, M, D)
    loc = torch.rand(1, S, M, L, P, 2)
    attn = torch.rand(1, S, M, L, P).flatten(-2).softmax(-1).view(1, S, M, L, P)
    order = tiled_item_order(shapes.tolist(), M, "cuda")
    o1, o2 = _run(v1, shapes, loc, attn, order), _run(v2, shapes, loc, attn, order)
    o12 = _run(2 * v1 - 3 * v2, shapes, loc, attn)
    assert (o12 - (2 * o1 - 3 * o2)).abs().max() < 1e-4
    assert (_run(v1, shapes, loc, 0.5 * attn) - 0.5 * o1).abs().max() < 1e-5
    idx = torch.randperm(S)[:256]
    ref = _oracle(v1, shapes, loc[:, idx].contiguous(), attn[:, idx].contiguous())
    assert (o1[:, idx] - ref).abs().max() < 2e-5 * max(1.0, ref.abs().max())


@pytest.mark.parametrize("vdtype,odtype,ref_dim,L,D", [
    (torch.float32, torch.float32, 2, 3, 32), (torch.float32, torch.float32, 4, 3, 32),
    (torch.bfloat16, torch.bfloat16, 2, 3, 32), (torch.bfloat16, torch.float32, 2, 1, 64),
    (torch.float32, torch.bfloat16, 2, 4, 64)])
def test_fused_variant_vs_oracle(vdtype, odtype, ref_dim, L, D):
    from dvis_plus_b200 import ops
    torch.manual_seed(11)
    shapes = torch.as_tensor([(13, 17), (7, 9), (4, 5), (2, 3)][:L])
    S = int(shapes.prod(1).sum())
    N, M, P, Lq = 2, 8, 4, 53
    value = torch.randn(N, S, M, D)
    fused = torch.randn(N, Lq, M * L * P * 3)                 # one linear output: [offsets | logits]
    offsets, logits = fused[..., :M * L * P * 2], fused[..., M * L * P * 2:]
    ref_pts = torch.rand(N, Lq, L, ref_dim)
    if ref_dim == 4:
        ref_pts[..., 2:] *= 0.4
    # oracle side: softmax + location arithmetic exactly as OPS/modules/ms_deform_attn.py:101-112
    aw = logits.reshape(N, Lq, M, L * P).softmax(-1).view(N, Lq, M, L, P)
    off = offsets.reshape(N, Lq, M, L, P, 2)
    if ref_dim == 2:
        norm = torch.stack([shapes[:, 1], shapes[:, 0]], -1).float()
        loc = ref_pts[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    else:
        loc = ref_pts[:, :, None, :, None, :2] + off / P * ref_pts[:, :, None, :, None, 2:] * 0.5
    vq = value.to(vdtype).float()
    ref = _oracle(vq, shapes, loc.contiguous(), aw.contiguous())
    dev = "cuda"
    f = fused.to(dev)
    out = ops.msda_fused_forward(value.to(dev).to(vdtype), shapes.to(dev), lsi_of(shapes).to(dev),
                                 f[..., :M * L * P * 2], f[..., M * L * P * 2:], ref_pts.to(dev), M, L, P,
                                 out_dtype=odtype).float().cpu()
    tol = 1e-2 if odtype == torch.bfloat16 else 1e-4
    assert (out - ref).abs().max() <= tol * max(1.0, ref.abs().max())


@pytest.mark.parametrize("ref_dim,L,pdt", [(2, 3, torch.bfloat16), (4, 3, torch.float32), (2, 1, torch.bfloat16), (2, 4, torch.float32)])
def test_pair_packed_variant_vs_oracle(ref_dim, L, pdt):
    """bf16 pair-packed kernel: same contract as the fused variant; checked against the oracle on bf16-rounded value."""
    from dvis_plus_b200 import ops
    torch.manual_seed(5)
    shapes = torch.as_tensor([(13, 17), (7, 9), (4, 5), (2, 3)][:L])
    S = int(shapes.prod(1).sum())
    N, M, D, P, Lq = 2, 8, 32, 4, 61
    value = torch.randn(N, S, M, D)
    fused = (torch.randn(N, Lq, M * L * P * 3) * 1.5).to(pdt)
    offsets, logits = fused[..., :M * L * P * 2], fused[..., M * L * P * 2:]
    ref_pts = torch.rand(N, Lq, L, ref_dim) * 1.2 - 0.1          # some points fall outside the map
    if ref_dim == 4:
        ref_pts[..., 2:] = ref_pts[..., 2:].abs() * 0.4
    aw = logits.float().reshape(N, Lq, M, L * P).softmax(-1).view(N, Lq, M, L, P)
    off = offsets.float().reshape(N, Lq, M, L, P, 2)
    if ref_dim == 2:
        norm = torch.stack([shapes[:, 1], shapes[:, 0]], -1).float()
        loc = ref_pts[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    else:
        loc = ref_pts[:, :, None, :, None, :2] + off / P * ref_pts[:, :, None, :, None, 2:] * 0.5
    ref = _oracle(value.bfloat16().float(), shapes, loc.contiguous(), aw.contiguous())
    dev = "cuda"
    f = fused.to(dev)
    out = ops.msda_pair_forward(value.to(dev).bfloat16(), shapes.to(dev), lsi_of(shapes).to(dev), f[..., :M * L * P * 2],
                                f[..., M * L * P * 2:], ref_pts.to(dev), M, L, P).float().cpu()
    assert (out - ref).abs().max() <= 1e-2 * max(1.0, ref.abs().max())
    # and against the staged fused kernel on identical inputs (differs only by fp16 weights / bf16 output rounding)
    out2 = ops.msda_fused_forward(value.to(dev).bfloat16(), shapes.to(dev), lsi_of(shapes).to(dev), f[..., :M * L * P * 2],
                                  f[..., M * L * P * 2:], ref_pts.to(dev), M, L, P).float().cpu()
    assert (out - out2).abs().max() <= 1e-2 * max(1.0, ref.abs().max())


@pytest.mark.parametrize("kind", ["injector", "extractor"])
def test_vit_adapter_call_site_shapes(kind):
    """Second consumer of the op: ViT-Adapter Injector / Extractor (P/mask2former/modeling/backbones_vitAdapter/
    adapter.py:101-165, deform_inputs :39-58): d_model 1024 / 16 heads -> D=64; Injector: queries = ViT tokens (1 level grid),
    values = 3-level pyramid; Extractor: queries = pyramid tokens, values = 1 level; Lq != S in both."""
    from dvis_plus_b200 import ops
    torch.manual_seed(2)
    M, D, P = 16, 64, 4
    if kind == "injector":
        shapes = torch.as_tensor([(16, 24), (8, 12), (4, 6)])     # 1/8, 1/16, 1/32 pyramid
        Lq = 8 * 12                                               # ViT tokens at 1/16
    else:
        shapes = torch.as_tensor([(8, 12)])
        Lq = 16 * 24 + 8 * 12 + 4 * 6
    L = shapes.shape[0]
    S = int(shapes.prod(1).sum())
    value = torch.randn(2, S, M, D)
    loc = torch.rand(2, Lq, M, L, P, 2)
    attn = torch.rand(2, Lq, M, L, P).flatten(-2).softmax(-1).view(2, Lq, M, L, P)
    out = _run(value, shapes, loc, attn)
    ref = _oracle(value, shapes, loc, attn)
    assert out.shape == (2, Lq, M * D)
    assert (out - ref).abs().max() < 2e-5 * max(1.0, ref.abs().max())


def test_plain_op_propagates_nonfinite_values_like_the_reference():
    """Corners outside the map are skipped, not multiplied by 0 (ms_deform_im2col_cuda.cuh:61-83): NaN / Inf in `value` --
    including pixel 0 of a head -- reach exactly the outputs they reach in the reference (oracle = its conditionals)."""
    import numpy as np
    from dvis_plus_b200 import ops
    from oracle import c_oracle
    g = torch.Generator().manual_seed(4)
    sh = torch.tensor([(23, 40), (12, 20), (6, 10)])
    lsi = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    S = int(sh.prod(1).sum())
    value = torch.randn(2, S, 8, 32, generator=g)
    value[:, 0] = float("nan")
    value[0, 500, 1, 3] = float("inf")
    value[1, 1100, 2] = float("nan")
    loc = torch.rand(2, 300, 8, 3, 4, 2, generator=g) * 1.6 - 0.3
    attn = torch.rand(2, 300, 8, 3, 4, generator=g)
    out = ops.ms_deform_attn_forward(value.cuda(), sh.cuda(), lsi.cuda(), loc.cuda(), attn.cuda(), 128).cpu().numpy()
    ref = c_oracle.msda_forward(value.numpy(), sh.numpy(), lsi.numpy(), loc.numpy(), attn.numpy())
    assert np.array_equal(np.isnan(out), np.isnan(ref)) and np.array_equal(np.isinf(out), np.isinf(ref))
    assert np.isnan(ref).any() and (~np.isnan(ref)).any()
    fin = np.isfinite(ref)
    assert np.abs(out[fin] - ref[fin]).max() <= 2e-5 * max(1.0, np.abs(ref[fin]).max())
