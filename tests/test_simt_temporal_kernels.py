"""Round-2 kernels of the temporal stage executed on the SIMT emulator (tests/simt): the warp-MMA attention core
(csrc/flash_attn.cu) with the emulated mma.sync / ldmatrix of csrc/mma.cuh -- fragment bookkeeping, key-block split
across warps, online-softmax merge, bit masks, ragged tails -- against an fp32 softmax(QK^T)V on the same bf16 inputs.
Small sizes: one OS thread per CUDA thread."""
import math
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt"))
import simt_binding as simt  # noqa: E402


def ref_attention(q, k, v, scale, mask=None):
    """q (B,Lq,H,D), k / v (B,Lk,H,D) -> (B,Lq,H*D) fp32; mask bool (B,Lq,Lk), True = masked."""
    s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) * scale
    if mask is not None:
        s = s.masked_fill(mask[:, None], float("-inf"))
    p = s.softmax(-1)
    return torch.einsum("bhqk,bkhd->bqhd", p, v.float()).flatten(2)


def pack_bits(mask):
    """bool (B, Lq, Lk) -> uint8 (B, Lq, ceil(Lk/64)*8), bit (j % 8) of byte (j / 8) = mask[..., j]."""
    B, Lq, Lk = mask.shape
    nbytes = (Lk + 63) // 64 * 8
    m = torch.zeros(B, Lq, nbytes * 8, dtype=torch.bool)
    m[..., :Lk] = mask
    w = (2 ** torch.arange(8, dtype=torch.int32))
    return (m.view(B, Lq, nbytes, 8).int() * w).sum(-1).to(torch.uint8).contiguous()


@pytest.mark.parametrize("B,Lq,Lk,H,Dh", [(1, 20, 40, 2, 64), (2, 16, 7, 1, 32), (1, 33, 200, 1, 64), (1, 5, 330, 2, 32),
                                          (1, 70, 530, 1, 32), (1, 9, 577, 1, 64),       # Lk > 512: the row-split variant
                                          (1, 150, 530, 1, 32)])                         # two 128-row CTAs per head
def test_flash_attn_matches_fp32_reference(B, Lq, Lk, H, Dh):
    torch.manual_seed(B * 1000 + Lq * 10 + Lk)
    scale = 1 / math.sqrt(Dh)
    # q / k / v as strided slices of one packed projection, like the tracker's QKV GEMM output
    qkv_q = torch.randn(B, Lq, 3, H, Dh).to(torch.bfloat16)
    kv = torch.randn(B, Lk, 2, H, Dh).to(torch.bfloat16)
    q, k, v = qkv_q[:, :, 0], kv[:, :, 0], kv[:, :, 1]
    out = simt.flash_attn(q, k, v, scale)
    ref = ref_attention(q, k, v, scale)
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max() < 2e-2 * ref.abs().max()          # bf16 P and bf16 output rounding


@pytest.mark.parametrize("variant", [1, 2, 3, 4])   # 128-row, 64-row tiles, key split inside the CTA, 32-row tiles (DVIS_FLASH_VARIANT)
@pytest.mark.parametrize("B,Lq,Lk,H,Dh,masked", [(1, 70, 530, 1, 32, True), (1, 150, 577, 1, 64, False), (2, 37, 130, 2, 32, True)])
def test_flash_attn_every_variant_on_the_same_problem(variant, B, Lq, Lk, H, Dh, masked, monkeypatch):
    """The dispatch picks the tiling from the problem size (long memory, few (batch, head) pairs -> smaller row tiles -> key split);
    every variant must give the same result on any problem, with and without the bit mask."""
    monkeypatch.setenv("DVIS_FLASH_VARIANT", str(variant))
    torch.manual_seed(variant * 7 + Lq)
    q = torch.randn(B, Lq, H, Dh).to(torch.bfloat16)
    k = torch.randn(B, Lk, H, Dh).to(torch.bfloat16)
    v = torch.randn(B, Lk, H, Dh).to(torch.bfloat16)
    mask = None
    if masked:
        mask = torch.rand(B, Lq, Lk) < 0.6
        mask[:, :, 5] = False
    out = simt.flash_attn(q, k, v, 0.25, pack_bits(mask) if masked else None)
    ref = ref_attention(q, k, v, 0.25, mask)
    assert (out.float() - ref).abs().max() < 2e-2 * ref.abs().max()


def test_flash_attn_two_stage_ring_and_bit_mask():
    """Lk > 4 blocks -> 2-stage cp.async ring; random mask with at least one open key per row."""
    torch.manual_seed(5)
    B, Lq, Lk, H, Dh = 1, 18, 600, 1, 32
    q = torch.randn(B, Lq, H, Dh).to(torch.bfloat16)
    k = torch.randn(B, Lk, H, Dh).to(torch.bfloat16)
    v = torch.randn(B, Lk, H, Dh).to(torch.bfloat16)
    mask = torch.rand(B, Lq, Lk) < 0.7
    mask[:, :, 17] = False
    mask[0, 3, :] = True                                                       # one row keeps a single open key, in block 9
    mask[0, 3, 590] = False
    out = simt.flash_attn(q, k, v, 0.3, pack_bits(mask))
    ref = ref_attention(q, k, v, 0.3, mask)
    assert (out.float() - ref).abs().max() < 2e-2 * ref.abs().max()
    assert (out.float()[0, 3] - v[0, 590, 0].float()).abs().max() < 1e-2


def test_flash_attn_rejects_bad_arguments():
    q = torch.zeros(1, 4, 1, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="head dim"):
        simt.flash_attn(torch.zeros(1, 4, 1, 48, dtype=torch.bfloat16), torch.zeros(1, 4, 1, 48, dtype=torch.bfloat16),
                        torch.zeros(1, 4, 1, 48, dtype=torch.bfloat16), 1.0)
    with pytest.raises(RuntimeError, match="mask rows"):
        simt.flash_attn(q, q, q, 1.0, torch.zeros(1, 4, 4, dtype=torch.uint8))


# ---- csrc/small_linear.cu ------------------------------------------------------------------------------------------------

def _ln(x, g, b, eps=1e-5):
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), g, b, eps)


def _bf(x):
    return x.to(torch.bfloat16).float()


def test_linear_small_plain_bias_relu_residual():
    torch.manual_seed(0)
    M, N, K = 45, 72, 128                                   # ragged rows, N not a multiple of the 64-column tile
    x = torch.randn(M, K).to(torch.bfloat16)
    w = (torch.randn(N, K) / K ** 0.5).to(torch.bfloat16)
    b, res = torch.randn(N), torch.randn(M, N)
    y32, y16, _, _ = simt.linear_small(w, b, x=x, relu=True, residual=res, out_f32=True)
    ref = torch.relu(x.float() @ w.float().t() + b) + res
    assert (y32 - ref).abs().max() < 1e-4
    assert (y16.float() - ref).abs().max() < 2e-2
    # strided rows (a slice of a packed projection) and no bias
    packed = torch.randn(M, 3, K).to(torch.bfloat16)
    y32, _, _, _ = simt.linear_small(w, None, x=packed[:, 1], out_f32=True, out_bf16=False)
    assert (y32 - packed[:, 1].float() @ w.float().t()).abs().max() < 1e-4


def test_linear_small_batched():
    torch.manual_seed(1)
    B, M, N, K = 3, 20, 64, 64
    x = torch.randn(B, M, K).to(torch.bfloat16)
    w = (torch.randn(B, N, K) / 8).to(torch.bfloat16)
    b = torch.randn(B, N)
    y32, y16, _, _ = simt.linear_small(w, b, x=x, out_f32=True)
    ref = torch.einsum("bmk,bnk->bmn", x.float(), w.float()) + b[:, None]
    assert (y32 - ref).abs().max() < 1e-4 and (y16.float() - ref).abs().max() < 3e-2


def test_linear_small_layernorm_prologue_and_side_outputs():
    """A = LN1(LN0(src0) + src1): the tracker's  LN_cross(o_all[j] + LN_ffn(pre_ffn))  feeding the QKV projection."""
    torch.manual_seed(2)
    M, N, K = 37, 192, 256
    src0, src1 = torch.randn(M, K) * 2 + 0.5, torch.randn(M, K).to(torch.bfloat16)
    g0, b0, g1, b1 = torch.rand(K) + 0.5, torch.randn(K) * 0.1, torch.rand(K) + 0.5, torch.randn(K) * 0.1
    w, b = (torch.randn(N, K) / K ** 0.5).to(torch.bfloat16), torch.randn(N)
    y32, _, s0, s1 = simt.linear_small(w, b, src0=src0, ln0=(g0, b0), src1=src1, ln1=(g1, b1), want_side0=True, want_side1=True,
                                      out_f32=True, out_bf16=False)
    r0 = _ln(src0, g0, b0)
    r1 = _ln(r0 + src1.float(), g1, b1)
    assert (s0 - r0).abs().max() < 1e-5 and (s1 - r1).abs().max() < 1e-5
    assert (y32 - (_bf(r1) @ w.float().t() + b)).abs().max() < 2e-3       # the only rounding: A to bf16 (a flip moves a term by 1 ulp)
    # single LayerNorm, fp32 second source, N = 64 (one CTA column writes every side segment)
    w1 = w[:64].contiguous()
    y32, _, s0, s1 = simt.linear_small(w1, None, src0=src0, src1=r0, ln1=(g1, b1), want_side1=True, out_f32=True, out_bf16=False)
    r = _ln(src0 + r0, g1, b1)
    assert s0 is None and (s1 - r).abs().max() < 1e-5
    assert (y32 - _bf(r) @ w1.float().t()).abs().max() < 2e-3
    # no LayerNorm at all: plain fp32 -> bf16 operand
    y32, _, _, _ = simt.linear_small(w1, None, src0=src0, out_f32=True, out_bf16=False)
    assert (y32 - _bf(src0) @ w1.float().t()).abs().max() < 2e-3


def test_linear_small_conv1d_taps_equal_replicate_padded_conv():
    """Conv1d(k=5 / k=3, padding='same', padding_mode='replicate') over time as a GEMM (P/dvis_Plus/refiner.py:44-52)."""
    torch.manual_seed(3)
    T, Q, C = 6, 5, 64
    x = torch.randn(T, Q, C).to(torch.bfloat16)              # rows = t * Q + q
    for k in (5, 3):
        conv = torch.nn.Conv1d(C, C, k, padding="same", padding_mode="replicate")
        wk = conv.weight.detach().permute(0, 2, 1).reshape(C, k * C).to(torch.bfloat16).contiguous()    # (C_out, k * C_in), tap-major
        y32, _, _, _ = simt.linear_small(wk, conv.bias.detach().float(), x=x.view(T * Q, C), taps=k, tap_pad=k // 2, tap_period=Q,
                                         tap_len=T, out_f32=True, out_bf16=False)
        conv_bf = torch.nn.Conv1d(C, C, k, padding="same", padding_mode="replicate")
        conv_bf.weight.data = conv.weight.detach().to(torch.bfloat16).float()
        conv_bf.bias.data = conv.bias.detach()
        ref = conv_bf(x.float().permute(1, 2, 0)).permute(2, 0, 1).reshape(T * Q, C)                # (q, c, t) -> rows t*Q+q
        assert (y32 - ref).abs().max() < 1e-4, k


def test_linear_small_64_row_tiles():
    torch.manual_seed(4)
    M, N, K = 530, 64, 128                                   # M > 512 selects the 64-row variant (refiner sizes)
    x = torch.randn(M, K).to(torch.bfloat16)
    w = (torch.randn(N, K) / K ** 0.5).to(torch.bfloat16)
    y32, _, _, _ = simt.linear_small(w, None, x=x, out_f32=True, out_bf16=False)
    assert (y32 - x.float() @ w.float().t()).abs().max() < 1e-4
    g = torch.rand(K) + 0.5
    src0 = torch.randn(M, K)
    y32, _, _, s1 = simt.linear_small(w, None, src0=src0, ln1=(g, torch.zeros(K)), want_side1=True, out_f32=True, out_bf16=False)
    r = _ln(src0, g, torch.zeros(K))
    assert (s1 - r).abs().max() < 1e-5 and (y32 - _bf(r) @ w.float().t()).abs().max() < 2e-3


def test_linear_small_rejects_bad_arguments():
    w = torch.zeros(64, 96, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="K %% 64|K % 64"):
        simt.linear_small(w, None, x=torch.zeros(4, 96, dtype=torch.bfloat16))
    w = torch.zeros(64, 640, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="prologue needs"):
        simt.linear_small(w, None, src0=torch.zeros(4, 640))


def test_linear_small_split_k_is_deterministic_and_exact():
    """K >= 1024 over a narrow output: 4-way split-K, partial tiles added in split order by the last CTA (no float atomics)"""
    torch.manual_seed(6)
    M, N, K = 40, 64, 1024
    x = torch.randn(M, K).to(torch.bfloat16)
    w = (torch.randn(N, K) / K ** 0.5).to(torch.bfloat16)
    b, res = torch.randn(N), torch.randn(M, N)
    y32, _, _, _ = simt.linear_small(w, b, x=x, residual=res, out_f32=True, out_bf16=False)
    ref = x.float() @ w.float().t() + b + res
    assert (y32 - ref).abs().max() < 1e-4
    again, _, _, _ = simt.linear_small(w, b, x=x, residual=res, out_f32=True, out_bf16=False)
    assert torch.equal(again, y32)                                    # the counters were left at zero, the sum order is fixed


@pytest.mark.parametrize("K", [128, 1024])
def test_linear_small_ln_epilogue_by_last_cta(K):
    """out-proj / FFN2 form: y = x W^T + b + residual; e1 = LN(y); e2 = LN(e1 + src1) -- normalised by the last CTA of each
    32-row block (K = 1024 also goes through the 4-way split-K)"""
    torch.manual_seed(K)
    M, N = 45, 256
    x = torch.randn(M, K).to(torch.bfloat16)
    w = (torch.randn(N, K) / K ** 0.5).to(torch.bfloat16)
    b, res = torch.randn(N), torch.randn(M, N)
    g1, b1, g2, b2 = torch.rand(N) + 0.5, torch.randn(N) * 0.1, torch.rand(N) + 0.5, torch.randn(N) * 0.1
    src1 = torch.randn(M, N).to(torch.bfloat16)
    e1_32, e1_16, e2_32, e2_16 = simt.linear_small_ln(w, b, x, res, (g1, b1), src1=src1, ln2=(g2, b2))
    y = x.float() @ w.float().t() + b + res
    r1 = _ln(y, g1, b1)
    r2 = _ln(r1 + src1.float(), g2, b2)
    assert (e1_32 - r1).abs().max() < 1e-4 and (e2_32 - r2).abs().max() < 1e-4
    assert torch.equal(e1_16, e1_32.to(torch.bfloat16)) and torch.equal(e2_16, e2_32.to(torch.bfloat16))
    # single LayerNorm, ReLU, no residual; and the counters were left at zero (second call gives the same bits)
    e1_32b, _, e2n, _ = simt.linear_small_ln(w, b, x, None, (g1, b1), relu=True)
    assert e2n is None and (e1_32b - _ln(torch.relu(x.float() @ w.float().t() + b), g1, b1)).abs().max() < 1e-4
    again = simt.linear_small_ln(w, b, x, res, (g1, b1), src1=src1, ln2=(g2, b2))
    assert torch.equal(again[0], e1_32) and torch.equal(again[2], e2_32)
