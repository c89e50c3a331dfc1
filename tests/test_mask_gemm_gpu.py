"""GPU parity of the tcgen05 mask-logit GEMM against the oracle (double-accumulated C restatement of the einsum at
P/dvis_Plus/video_mask2former_transformer_decoder.py:363).

Inputs are rounded to bf16 on both sides (the kernel computes bf16 x bf16 -> fp32), so the only difference is the
accumulation order: tolerance 1e-3 * scale for fp32 output, 1e-2 * scale for bf16 output (north star: 1e-2 bf16).
The TF32 tests feed unrounded fp32 operands and hold the fp32 tier's 1e-3.
"""
import pytest
import torch

from oracle import c_oracle

pytestmark = pytest.mark.gpu


def _case(B, Q, C, H, W, seed=0):
    g = torch.Generator().manual_seed(seed)
    emb = torch.randn(B, Q, C, generator=g).bfloat16().float()
    feat = torch.randn(B, C, H, W, generator=g).bfloat16().float()
    return emb, feat


@pytest.mark.parametrize("B,Q,C,H,W", [(1, 12, 64, 16, 24), (2, 100, 256, 23, 40), (1, 200, 256, 46, 80),
                                      (3, 16, 128, 8, 16), (2, 7, 64, 5, 8), (1, 256, 256, 16, 16), (2, 33, 512, 12, 20)])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_mask_logits_vs_oracle(B, Q, C, H, W, out_dtype):
    from dvis_plus_b200 import ops
    emb, feat = _case(B, Q, C, H, W, seed=Q)
    ref = torch.from_numpy(c_oracle.mask_logits(emb.numpy(), feat.numpy()))
    out = ops.mask_logits(emb.cuda(), feat.cuda().to(torch.bfloat16, memory_format=torch.channels_last), out_dtype)
    assert out.shape == (B, Q, H, W) and out.dtype == out_dtype
    tol = 1e-3 if out_dtype == torch.float32 else 1e-2
    err = (out.float().cpu() - ref).abs().max().item()
    assert err <= tol * ref.abs().max().item(), err


@pytest.mark.parametrize("B,Q,C,H,W", [(1, 12, 64, 16, 24), (2, 100, 256, 23, 40), (1, 200, 256, 46, 80), (2, 7, 64, 5, 8),
                                      (1, 300, 256, 9, 13), (2, 129, 128, 12, 20), (2, 200, 256, 184, 320)])
def test_mask_logits_tf32_operands_vs_oracle(B, Q, C, H, W):
    """fp32-tier mask head: UNROUNDED fp32 operands multiplied as TF32 on the tensor cores against the double-accumulated oracle:
    the north star's 1e-3 of the output scale (bf16 operands give ~4e-3).  Q > 128 exercises the query slices."""
    from dvis_plus_b200 import ops
    g = torch.Generator().manual_seed(Q + H)
    emb, feat = torch.randn(B, Q, C, generator=g), torch.randn(B, C, H, W, generator=g)
    out = ops.mask_logits(emb.cuda(), feat.cuda(), torch.float32, operand_dtype=torch.float32)
    assert out.shape == (B, Q, H, W) and out.dtype == torch.float32
    if H * W <= 4096:
        ref = torch.from_numpy(c_oracle.mask_logits(emb.numpy(), feat.numpy()))
        got = out.cpu()
    else:                                           # full size: a strided pixel subset keeps the oracle in seconds
        idx = torch.arange(0, H * W, 97)
        sub = feat.flatten(2)[:, :, idx].reshape(B, C, -1, 1).contiguous()
        ref = torch.from_numpy(c_oracle.mask_logits(emb.numpy(), sub.numpy()))
        got = out.flatten(2)[:, :, idx.cuda()].reshape(B, Q, -1, 1).cpu()
    assert (got - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()


def test_mask_attn_bias_tf32_operands():
    from dvis_plus_b200 import ops
    g = torch.Generator().manual_seed(3)
    B, Q, C, h, w = 2, 200, 256, 23, 40
    emb, feat = torch.randn(B, Q, C, generator=g), torch.randn(B, C, h, w, generator=g)
    feat[0] = feat[0].abs()                           # a row that is masked everywhere must come back all zeros:
    emb[0, 5] = -emb[0, 5].abs()                      # positive features x negative embedding
    logits = torch.einsum("bqc,bchw->bqhw", emb.double(), feat.double()).flatten(2)
    bias = ops.mask_attn_bias(emb.cuda(), feat.cuda().contiguous(memory_format=torch.channels_last), torch.float32).cpu()
    ref = torch.where(logits < 0, float("-inf"), 0.0)
    ref[(logits < 0).all(-1)] = 0.0
    safe = logits.abs() > 1e-2 * logits.abs().max()   # pixels whose sign TF32 rounding cannot flip
    assert (logits[0, 5] < 0).all() and (bias[0, 5] == 0).all()
    assert torch.equal(bias[safe], ref.float()[safe])


@pytest.mark.parametrize("T,Q,H,W,dt", [(16, 200, 46, 80, torch.bfloat16), (3, 100, 23, 40, torch.float32), (5, 7, 5, 7, torch.bfloat16),
                                        (2, 256, 16, 24, torch.float32)])
def test_mask_logits_clip_layout(T, Q, H, W, dt):
    """dvis_mask_logits_clip writes "q t h w" (refiner.py:185-189) directly: bit-identical to the (t, q, h, w) GEMM transposed."""
    from dvis_plus_b200 import ops
    emb, feat = _case(T, Q, 256, H, W, seed=T)
    feat_cl = feat.cuda().to(torch.bfloat16, memory_format=torch.channels_last)
    ref = ops.mask_logits(emb.cuda(), feat_cl, dt, operand_dtype=torch.bfloat16).permute(1, 0, 2, 3)
    out = ops.mask_logits_clip(emb.cuda(), feat_cl, dt)
    assert out.shape == (Q, T, H, W) and out.is_contiguous() and out.dtype == dt
    assert torch.equal(out, ref)


def test_mask_logits_golden_mask_head(golden):
    """The mask head's einsum on the reference's own fixture (mask_embed recomputed by the oracle port)."""
    from dvis_plus_b200 import ops
    from oracle import torch_port as tp
    g = golden("mask_head_small.pt")
    sd = golden("predictor_small.pt")["state_dict"]
    me = tp.mlp(sd, "mask_embed", tp.layer_norm(sd, "decoder_norm", g["output"]).transpose(0, 1))
    out = ops.mask_logits(me.cuda(), g["mask_features"].cuda()).cpu()
    scale = g["masks"].abs().max().item()
    assert (out - g["masks"]).abs().max().item() <= 1e-2 * scale   # bf16 inputs vs the fp32 reference


def test_mask_logits_accepts_nchw_fp32_and_splits_queries():
    from dvis_plus_b200 import ops
    for (B, H, W, dt) in ((2, 9, 16, torch.float32), (3, 5, 7, torch.float32), (2, 12, 20, torch.bfloat16)):
        emb, feat = _case(B, 300, 64, H, W, seed=5)   # Q=300 > 256 (DAQ stress size): two in-place query slices
        ref = torch.einsum("bqc,bchw->bqhw", emb, feat)
        out = ops.mask_logits(emb.cuda(), feat.cuda(), dt).float().cpu()
        tol = 1e-3 if dt == torch.float32 else 1e-2
        assert (out - ref).abs().max().item() <= tol * ref.abs().max().item()


def test_mask_logits_720p_full_size_properties():
    """BASELINE size (Q=200, C=256, 184x320, several frames): linearity in emb and a checksum against cuBLAS-free
    column sums computed on the host for a pixel subset."""
    from dvis_plus_b200 import ops
    B, Q, C, H, W = 3, 200, 256, 184, 320
    g = torch.Generator(device="cuda").manual_seed(0)
    emb1 = torch.randn(B, Q, C, device="cuda", generator=g).bfloat16()
    emb2 = torch.randn(B, Q, C, device="cuda", generator=g).bfloat16()
    feat = torch.randn(B, C, H, W, device="cuda", generator=g).to(torch.bfloat16, memory_format=torch.channels_last)
    o1, o2 = ops.mask_logits(emb1, feat), ops.mask_logits(emb2, feat)
    o12 = ops.mask_logits((emb1.float() + emb2.float()).bfloat16(), feat)
    # (emb1+emb2) is re-rounded to bf16: allow bf16 rounding of the sum times |feat| mass
    assert (o12 - (o1 + o2)).abs().max().item() < 0.05 * o1.abs().max().item()
    idx = torch.randint(0, H * W, (512,))
    fs = feat.permute(0, 2, 3, 1).reshape(B, H * W, C)[:, idx.cuda()].float().cpu()
    ref = torch.einsum("bqc,bpc->bqp", emb1.float().cpu().double(), fs.double()).float()
    got = o1.reshape(B, Q, H * W)[:, :, idx.cuda()].cpu()
    assert (got - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_mask_attn_bias_fused_epilogue(dtype):
    """Threshold fused into the GEMM epilogue == attn_bias_from_logits(mask_logits(...)); includes a fully masked row."""
    from dvis_plus_b200 import ops
    emb, feat = _case(3, 50, 64, 23, 40, seed=9)
    emb[1, 7] = 0
    emb[1, 7, 0] = -50.0
    feat[1, 0] = feat[1, 0].abs() + 0.1              # row (1,7): every logit < 0 -> must come out as all zeros
    e, f = emb.cuda(), feat.cuda().to(torch.bfloat16, memory_format=torch.channels_last)
    logits = ops.mask_logits(e, f, torch.float32).flatten(2)
    assert (logits[1, 7] < 0).all()
    ref = ops.attn_bias_from_logits(logits, dtype)
    got = ops.mask_attn_bias(e, f, dtype)
    assert torch.equal(got, ref)
    assert got[1, 7].abs().max() == 0 and torch.isinf(got).any()
