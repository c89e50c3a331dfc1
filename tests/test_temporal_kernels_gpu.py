"""Round-2 temporal-stage kernels on the B200 through the C ABI (ops front ends): dvis_flash_attn and dvis_linear_small
against fp32 torch references on the same bf16 inputs, at the sizes the tracker / refiner / predictor use them, and the
fused tracker / refiner module paths against the unmodified reference's golden outputs (hidden 128) and against the
library path at production width (hidden 512, Q = 200)."""
import math

import numpy as np
import pytest
import torch

from dvis_plus_b200 import _lib, ops
from dvis_plus_b200.modules.precision import precision
from test_simt_temporal_kernels import _bf, _ln, pack_bits, ref_attention

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return (a.double() - b.double()).abs().max().item() / max(1e-6, b.double().abs().max().item())


@pytest.mark.parametrize("B,Lq,Lk,H,Dh", [(1, 200, 200, 8, 64), (6, 200, 200, 8, 64), (16, 200, 200, 8, 64), (200, 16, 16, 8, 64),
                                          (2, 300, 300, 8, 32), (1, 100, 920, 8, 32), (1, 7, 1, 1, 64), (3, 33, 1000, 2, 32)])
def test_flash_attn_vs_fp32_reference(B, Lq, Lk, H, Dh):
    torch.manual_seed(Lq + Lk)
    scale = 1 / math.sqrt(Dh)
    qp = torch.randn(B, Lq, 3, H, Dh, device="cuda").to(torch.bfloat16)       # slices of packed projections
    kvp = torch.randn(B, Lk, 2, H, Dh, device="cuda").to(torch.bfloat16)
    q, k, v = qp[:, :, 0], kvp[:, :, 0], kvp[:, :, 1]
    n0 = _lib.launch_count
    out = ops.flash_attn(q, k, v, scale)
    assert _lib.launch_count == n0 + 1
    ref = ref_attention(q, k, v, scale)
    assert rel_err(out.float(), ref) < 1e-2                                    # north star: 1e-2 bf16


def test_flash_attn_time_attention_layout():
    """the refiner's attention over time: batch = queries, rows = frames, output written transposed (t, q, c)"""
    T, Q, H, Dh = 16, 200, 8, 64
    qkv = torch.randn(T, Q, 3, H, Dh, device="cuda").to(torch.bfloat16)
    v = qkv.permute(1, 0, 2, 3, 4)
    o = torch.empty(T, Q, H * Dh, device="cuda", dtype=torch.bfloat16)
    ops.flash_attn(v[:, :, 0], v[:, :, 1], v[:, :, 2], 0.125, out=o.permute(1, 0, 2))
    ref = ref_attention(v[:, :, 0], v[:, :, 1], v[:, :, 2], 0.125).permute(1, 0, 2)
    assert rel_err(o.float(), ref) < 1e-2


@pytest.mark.parametrize("Lk", [920, 3680, 14720])
def test_flash_attn_bit_mask_predictor_levels(Lk):
    """masked cross-attention at the predictor's three memory lengths (720p levels), 8 heads x 32"""
    torch.manual_seed(Lk)
    B, Lq, H, Dh = 2, 200, 8, 32
    q = torch.randn(B, Lq, H, Dh, device="cuda").to(torch.bfloat16)
    k = torch.randn(B, Lk, H, Dh, device="cuda").to(torch.bfloat16)
    v = torch.randn(B, Lk, H, Dh, device="cuda").to(torch.bfloat16)
    mask = torch.rand(B, Lq, Lk) < 0.8
    mask[:, :, 5] = False
    out = ops.flash_attn(q, k, v, 1 / math.sqrt(Dh), pack_bits(mask).cuda())
    ref = ref_attention(q, k, v, 1 / math.sqrt(Dh), mask.cuda())
    assert rel_err(out.float(), ref) < 1e-2


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4])   # auto, 128-row, 64-row tiles, key split, 32-row tiles (DVIS_FLASH_VARIANT)
@pytest.mark.parametrize("B,Lk", [(2, 3680), (4, 920), (16, 920), (1, 14720)])
def test_flash_attn_variants_agree_on_predictor_shapes(variant, B, Lk, monkeypatch):
    """Masked cross-attention of the predictor at the frames-per-rank of 8 / 4 / 1 GPUs: whichever tiling the dispatch picks (or is
    forced to), the result is the fp32 reference's."""
    monkeypatch.setenv("DVIS_FLASH_VARIANT", str(variant))
    g = torch.Generator().manual_seed(B + Lk)
    q = torch.randn(B, 200, 8, 32, generator=g).bfloat16()
    k = torch.randn(B, Lk, 8, 32, generator=g).bfloat16()
    v = torch.randn(B, Lk, 8, 32, generator=g).bfloat16()
    mask = torch.rand(B, 200, Lk, generator=g) < 0.7
    mask[:, :, 11] = False
    out = ops.flash_attn(q.cuda(), k.cuda(), v.cuda(), 32 ** -0.5, mask_bits=pack_bits(mask).cuda()).float().cpu()
    ref = ref_attention(q, k, v, 32 ** -0.5, mask)
    assert (out - ref).abs().max() < 2e-2 * ref.abs().max()


def test_flash_attn_rejects_cpu_tensors():
    q = torch.zeros(1, 4, 1, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.flash_attn(q, q, q, 1.0)


@pytest.mark.parametrize("M,N,K", [(200, 512, 512), (200, 1536, 512), (200, 2048, 512), (200, 512, 2048), (200, 3072, 512),
                                   (3200, 512, 512), (3200, 2048, 512), (3200, 512, 2048), (45, 72, 128)])
def test_linear_small_plain(M, N, K):
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    b, res = torch.randn(N, device="cuda"), torch.randn(M, N, device="cuda")
    y32, y16, _, _ = ops.linear_small(w, b, x=x, relu=True, residual=res, out_f32=True)
    ref = torch.relu(x.float() @ w.float().t() + b) + res
    assert rel_err(y32, ref) < 1e-3                                            # north star: 1e-3 fp32 (same bf16 operands)
    assert rel_err(y16.float(), ref) < 1e-2


def test_linear_small_batched_out_projections():
    L, Q, C = 6, 200, 512
    x = torch.randn(L, Q, C, device="cuda").to(torch.bfloat16)
    w = (torch.randn(L, C, C, device="cuda") / C ** 0.5).to(torch.bfloat16)
    b = torch.randn(L, C, device="cuda")
    y32, _, _, _ = ops.linear_small(w, b, x=x, out_f32=True, out_bf16=False)
    assert rel_err(y32, torch.einsum("bmk,bnk->bmn", x.float(), w.float()) + b[:, None]) < 1e-3


@pytest.mark.parametrize("M,N,K", [(200, 1536, 512), (3200, 512, 512), (200, 768, 256), (37, 192, 384)])
def test_linear_small_layernorm_prologue(M, N, K):
    torch.manual_seed(N)
    dev = "cuda"
    src0, src1 = torch.randn(M, K, device=dev) * 2 + 0.5, torch.randn(M, K, device=dev).to(torch.bfloat16)
    g0, b0, g1, b1 = (torch.rand(K, device=dev) + 0.5, torch.randn(K, device=dev) * 0.1, torch.rand(K, device=dev) + 0.5,
                      torch.randn(K, device=dev) * 0.1)
    w, b = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16), torch.randn(N, device=dev)
    y32, _, s0, s1 = ops.linear_small(w, b, src0=src0, ln0=(g0, b0), src1=src1, ln1=(g1, b1), want_side0=True, want_side1=True,
                                     out_f32=True, out_bf16=False)
    r0 = _ln(src0, g0, b0)
    r1 = _ln(r0 + src1.float(), g1, b1)
    assert (s0 - r0).abs().max() < 2e-5 and (s1 - r1).abs().max() < 2e-5       # every element of both side outputs written
    assert rel_err(y32, _bf(r1) @ w.float().t() + b) < 2e-3


@pytest.mark.parametrize("k", [5, 3])
def test_linear_small_conv1d_over_time(k):
    T, Q, C = 16, 200, 512
    x = torch.randn(T, Q, C, device="cuda").to(torch.bfloat16)
    conv = torch.nn.Conv1d(C, C, k, padding="same", padding_mode="replicate").cuda()
    conv.weight.data = conv.weight.data.to(torch.bfloat16).float()
    wk = conv.weight.detach().permute(0, 2, 1).reshape(C, k * C).to(torch.bfloat16).contiguous()
    y32, _, _, _ = ops.linear_small(wk, conv.bias.detach().float(), x=x.view(T * Q, C), taps=k, tap_pad=k // 2, tap_period=Q,
                                    tap_len=T, out_f32=True, out_bf16=False)
    with torch.no_grad():
        ref = conv(x.float().permute(1, 2, 0)).permute(2, 0, 1).reshape(T * Q, C)
    assert rel_err(y32, ref) < 1e-3


@torch.no_grad()
def test_tracker_and_refiner_fused_paths_vs_reference_golden(golden):
    from test_simt_modules import build_refiner_w128, build_tracker_w128
    g = golden("tracker_w128.pt")
    t = build_tracker_w128(g).cuda()
    t.use_fused_kernels = True
    fe, fn, mf = g["frame_embeds"].cuda(), g["frame_embeds_no_norm"].cuda(), g["mask_features"].cuda()
    with precision("bf16"):
        o1, i1 = t(fe[:, :, :3], mf[:, :3], resume=False, return_indices=True, frame_embeds_no_norm=fn[:, :, :3])
        o2, i2 = t(fe[:, :, 3:], mf[:, 3:], resume=True, return_indices=True, frame_embeds_no_norm=fn[:, :, 3:])
    for a, b in zip(i1 + i2, g["indices"]):
        assert np.array_equal(np.asarray(a), b.numpy())
    assert rel_err(torch.cat([o1["pred_embds"], o2["pred_embds"]], 2).float().cpu(), g["pred_embds"]) < 1e-2
    assert rel_err(torch.cat([o1["pred_logits"], o2["pred_logits"]], 1).float().cpu(), g["pred_logits"]) < 1e-2
    assert rel_err(torch.cat([o1["pred_masks"], o2["pred_masks"]], 2).float().cpu(), g["pred_masks"]) < 1e-2
    g = golden("refiner_w128.pt")
    r = build_refiner_w128(g).cuda()
    r.use_fused_kernels = True
    n0 = _lib.launch_count
    with precision("bf16"):
        o = r(g["instance_embeds"].cuda(), g["frame_embeds"].cuda(), g["mask_features"].cuda())
    assert _lib.launch_count - n0 >= 30
    for k in ("pred_embds", "pred_logits", "pred_masks"):
        assert rel_err(o[k].float().cpu(), g[k]) < 1e-2, k


@torch.no_grad()
def test_fused_temporal_stage_equals_library_path_at_production_width():
    """hidden 512, Q = 200, T = 16 (BASELINE config 4): fused kernels vs the cuBLAS / cuDNN path of round 1, and the
    per-frame CUDA graph vs the plain loop (bit-identical: same kernels, same order)."""
    import bench
    torch.manual_seed(1)
    runner = bench.build_models("cuda", queries=200)
    T, Q = 16, 200
    base = torch.randn(1, 512, 1, Q, device="cuda")
    fe = base + 0.3 * torch.randn(1, 512, T, Q, device="cuda")
    fn = fe + 0.1 * torch.randn(1, 512, T, Q, device="cuda")
    trk, rfn = runner.tracker, runner.refiner
    res = {}
    for name, fused, graph in (("fused", True, True), ("fused_nograph", True, False), ("library", False, True)):
        trk.use_fused_kernels = rfn.use_fused_kernels = fused
        trk.use_custom_attention = fused                  # "library": cuBLAS + cuDNN SDPA only (round 1's path)
        trk.use_cuda_graph = graph
        with precision("bf16"):
            o = trk(fe, None, resume=False, frame_embeds_no_norm=fn, with_masks=False)
            res[name] = (o["pred_embds"].float(), o["pred_logits"].float(), rfn.refine(o["pred_embds"], fn).float())
    for a, b in zip(res["fused"], res["fused_nograph"]):
        assert torch.equal(a, b)
    for a, b in zip(res["fused"], res["library"]):
        assert rel_err(a, b) < 1e-2, rel_err(a, b)


@pytest.mark.parametrize("B,Q,hw", [(2, 200, (23, 40)), (1, 100, (46, 80)), (2, 300, (34, 60)), (1, 7, (5, 13))])
def test_mask_attn_bits_equal_thresholded_logits(B, Q, hw):
    """the tcgen05 mask GEMM's bit epilogue == (E @ F < 0) packed, fully masked rows cleared (decoder.py:297,370-371)"""
    torch.manual_seed(Q)
    h, w = hw
    C = 256
    emb = torch.randn(B, Q, C, device="cuda").to(torch.bfloat16)
    emb[0, 1] = -emb[0, 1].abs()                                     # a row that ends up fully masked ...
    feat = torch.randn(B, C, h, w, device="cuda").abs().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    emb[0, 2] = emb[0, 2].abs()                                      # ... and one fully open
    bits = ops.mask_attn_bits(emb, feat)
    logits = ops.mask_logits(emb, feat, torch.float32).flatten(2)    # same GEMM, plain epilogue
    masked = logits < 0
    masked[masked.all(-1)] = False
    HW = h * w
    got = ((bits[..., None].int() >> torch.arange(8, device="cuda")) & 1).bool().flatten(2)[..., :HW]
    assert bits.shape[-1] % 8 == 0 and bits.shape[-1] * 8 >= HW
    assert torch.equal(got, masked)
    assert not got[0, 1].any() and not got[0, 2].any()


@torch.no_grad()
def test_predictor_flash_attention_equals_library_attention(golden):
    """production width (hidden 256, 8 heads x 32, Q = 200, 720p levels): bit-mask + dvis_flash_attn path vs the dense
    additive bias + cuDNN SDPA path of round 1"""
    import bench
    runner = bench.build_models("cuda", queries=200)
    dec = runner.predictor
    torch.manual_seed(0)
    ms = [torch.randn(2, 256, 23, 40, device="cuda"), torch.randn(2, 256, 46, 80, device="cuda"), torch.randn(2, 256, 92, 160, device="cuda")]
    mf = torch.randn(2, 256, 184, 320, device="cuda")
    outs = {}
    for fused in (True, False):
        dec.use_fused_attention = fused
        with precision("bf16"):
            outs[fused] = dec(ms, mf)
    for k in ("pred_logits", "pred_embds", "pred_masks"):
        assert rel_err(outs[True][k].float(), outs[False][k].float()) < 3e-2, (k, rel_err(outs[True][k].float(), outs[False][k].float()))


@pytest.mark.parametrize("M,N,K", [(200, 512, 512), (200, 512, 2048), (77, 256, 1024)])
def test_linear_small_ln_epilogue(M, N, K):
    """out-proj / FFN2 form at tracker sizes: LayerNorm(s) by the last CTA of each 32-row block (K >= 1024: 4-way split-K)"""
    torch.manual_seed(K + M)
    dev = "cuda"
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    b, res = torch.randn(N, device=dev), torch.randn(M, N, device=dev)
    g1, b1, g2, b2 = (torch.rand(N, device=dev) + 0.5, torch.randn(N, device=dev) * 0.1, torch.rand(N, device=dev) + 0.5,
                      torch.randn(N, device=dev) * 0.1)
    src1 = torch.randn(M, N, device=dev).to(torch.bfloat16)
    y = x.float() @ w.float().t() + b + res
    r1 = _ln(y, g1, b1)
    r2 = _ln(r1 + src1.float(), g2, b2)
    outs = [ops.linear_small_ln(w, b, x, res, (g1, b1), src1=src1, ln2=(g2, b2)) for _ in range(3)]
    e1_32, e1_16, e2_32, e2_16 = outs[0]
    assert rel_err(e1_32, r1) < 1e-3 and rel_err(e2_32, r2) < 1e-3
    assert torch.equal(e1_16, e1_32.to(torch.bfloat16)) and torch.equal(e2_16, e2_32.to(torch.bfloat16))
    for o in outs[1:]:                                  # counters self-clean, sums in a fixed order: bit-identical repeats
        assert torch.equal(o[0], e1_32) and torch.equal(o[2], e2_32)
