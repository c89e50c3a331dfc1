"""Torch-tensor front ends of the C ABI (include/dvis_b200.h).

`ms_deform_attn_forward` / `ms_deform_attn_backward` have the exact signatures of the reference's compiled
module `MultiScaleDeformableAttention` (OPS/src/vision.cpp:18-21), including its argument checks
(OPS/src/cuda/ms_deform_attn_cuda.cu:33-57): `install_as_reference_extension()` registers this module under
that name so `OPS/functions/ms_deform_attn_func.py:22` imports it unchanged.
"""
import sys
import types

import torch

from . import _lib
from ._lib import DVIS_BF16, DVIS_F32, DVIS_F64

_DTYPE = {torch.float32: DVIS_F32, torch.float64: DVIS_F64, torch.bfloat16: DVIS_BF16}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check_cuda_contig(**tensors):
    for name, t in tensors.items():
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
        if not t.is_cuda:
            # the reference dispatcher raises exactly this for CPU tensors (OPS/src/ms_deform_attn.h:43)
            raise RuntimeError("Not implemented on the CPU" if name == "value" else f"{name} must be a CUDA tensor")


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step,
                           item_order=None):
    """value (N,S,M,D); spatial_shapes (L,2) int64 cuda; level_start_index (L,) int64 cuda;
    sampling_loc (N,Lq,M,L,P,2); attn_weight (N,Lq,M,L,P) -> (N, Lq, M*D).  fp32 or fp64."""
    _check_cuda_contig(value=value, spatial_shapes=spatial_shapes, level_start_index=level_start_index,
                       sampling_loc=sampling_loc, attn_weight=attn_weight)
    N, S, M, D = value.shape
    L = spatial_shapes.shape[0]
    Lq, P = sampling_loc.shape[1], sampling_loc.shape[4]
    step = min(N, int(im2col_step))
    if step <= 0:
        raise RuntimeError(f"im2col_step({im2col_step}) and batch({N}) must be positive")
    if N % step != 0:
        raise RuntimeError(f"batch({N}) must divide im2col_step({step})")
    if value.dtype not in (torch.float32, torch.float64):
        raise RuntimeError(f'"ms_deform_attn_forward_cuda" not implemented for \'{value.dtype}\'')
    if sampling_loc.dtype != value.dtype or attn_weight.dtype != value.dtype:
        raise RuntimeError("value, sampling_loc and attn_weight must have the same dtype")
    if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
        raise RuntimeError("spatial_shapes and level_start_index must be int64")
    if sampling_loc.shape != (N, Lq, M, L, P, 2) or attn_weight.shape != (N, Lq, M, L, P):
        raise RuntimeError("sampling_loc / attn_weight shapes do not match value / spatial_shapes")
    out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
    if out.numel() == 0:
        return out
    order_ptr = None
    if item_order is not None:
        assert item_order.dtype == torch.int32 and item_order.is_cuda and item_order.numel() == Lq * M
        order_ptr = item_order.data_ptr()
    with torch.cuda.device(value.device):
        _lib.call("dvis_msda_forward", value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                  sampling_loc.data_ptr(), attn_weight.data_ptr(), N, S, M, D, L, Lq, P, _DTYPE[value.dtype],
                  order_ptr, out.data_ptr(), _stream())
    return out


def _check_item_order(item_order, n_items, device):
    """item_order: optional int32 CUDA permutation of the Lq*M (query, head) items (locality.tiled_item_order)."""
    if item_order is None:
        return None
    assert item_order.dtype == torch.int32 and item_order.is_cuda and item_order.device == device and item_order.is_contiguous()
    assert item_order.numel() == n_items, f"item_order has {item_order.numel()} entries, expected Lq*M = {n_items}"
    return item_order.data_ptr()


def msda_fused_forward(value, spatial_shapes, level_start_index, offsets, logits, reference_points, num_heads,
                       num_levels, num_points, item_order=None, out_dtype=None):
    """Fused softmax + location arithmetic + gather (dvis_msda_fused_forward).

    value (N,S,M,D) f32|bf16; offsets (N,Lq,M*L*P*2) and logits (N,Lq,M*L*P) f32|bf16 -- possibly column slices of one
    (N,Lq,M*L*P*3) tensor (last dim contiguous); reference_points (N,Lq,L,2|4) f32.
    """
    N, S, M, D = value.shape
    Lq = offsets.shape[1]
    assert value.is_contiguous() and reference_points.is_contiguous()
    assert offsets.dtype == logits.dtype and offsets.dtype in (torch.float32, torch.bfloat16)
    assert reference_points.dtype == torch.float32
    assert offsets.stride(-1) == 1 and logits.stride(-1) == 1
    assert offsets.stride(0) == offsets.stride(1) * Lq and logits.stride(0) == logits.stride(1) * Lq
    out_dtype = out_dtype or value.dtype
    out = torch.empty((N, Lq, M * D), dtype=out_dtype, device=value.device)
    order_ptr = _check_item_order(item_order, Lq * M, value.device)
    with torch.cuda.device(value.device):
        _lib.call("dvis_msda_fused_forward", value.data_ptr(), _DTYPE[value.dtype], spatial_shapes.data_ptr(),
                  level_start_index.data_ptr(), offsets.data_ptr(), offsets.stride(1), logits.data_ptr(),
                  logits.stride(1), _DTYPE[offsets.dtype], reference_points.data_ptr(), reference_points.shape[-1],
                  N, S, M, D, num_levels,
                  Lq, num_points, order_ptr, out.data_ptr(), _DTYPE[out_dtype], _stream())
    return out


def msda_fused_forward_hm(value_hm, spatial_shapes, level_start_index, offsets, logits, reference_points, num_levels,
                          num_points, item_order=None):
    """msda_fused_forward on a head-major value tensor (N, M, S, 32) bf16 -> (N, Lq, M*32) bf16 (dvis_msda_fused_forward_hm)."""
    N, M, S, D = value_hm.shape
    Lq = offsets.shape[1]
    assert value_hm.dtype == torch.bfloat16 and D == 32 and value_hm.is_contiguous() and reference_points.is_contiguous()
    assert offsets.dtype == logits.dtype and offsets.dtype in (torch.float32, torch.bfloat16)
    assert reference_points.dtype == torch.float32
    assert offsets.stride(-1) == 1 and logits.stride(-1) == 1
    assert offsets.stride(0) == offsets.stride(1) * Lq and logits.stride(0) == logits.stride(1) * Lq
    out = torch.empty((N, Lq, M * D), dtype=torch.bfloat16, device=value_hm.device)
    order_ptr = _check_item_order(item_order, Lq * M, value_hm.device)
    with torch.cuda.device(value_hm.device):
        _lib.call("dvis_msda_fused_forward_hm", value_hm.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                  offsets.data_ptr(), offsets.stride(1), logits.data_ptr(), logits.stride(1), _DTYPE[offsets.dtype],
                  reference_points.data_ptr(), reference_points.shape[-1], N, S, M, D, num_levels, Lq, num_points,
                  order_ptr, out.data_ptr(), _stream())
    return out


def msda_pair_forward(value, spatial_shapes, level_start_index, offsets, logits, reference_points, num_heads,
                      num_levels, num_points, item_order=None):
    """Pair-packed bf16 fused forward: packs `value` (N,S,M,32) bf16 into x-adjacent corner pairs, then gathers 2 lines
    per sampling point (dvis_msda_pack_pairs + dvis_msda_pair_forward).  Same contract as msda_fused_forward."""
    N, S, M, D = value.shape
    Lq = offsets.shape[1]
    assert value.dtype == torch.bfloat16 and D == 32 and value.is_contiguous() and reference_points.is_contiguous()
    assert reference_points.dtype == torch.float32, "reference points are read as float32"
    assert offsets.dtype == logits.dtype and offsets.dtype in (torch.float32, torch.bfloat16)
    assert offsets.stride(-1) == 1 and logits.stride(-1) == 1
    assert offsets.stride(0) == offsets.stride(1) * Lq and logits.stride(0) == logits.stride(1) * Lq
    _check_item_order(item_order, Lq * M, value.device)
    pairs = torch.empty((N, S + 1, M, 2, D), dtype=torch.bfloat16, device=value.device)
    out = torch.empty((N, Lq, M * D), dtype=torch.bfloat16, device=value.device)
    order_ptr = item_order.data_ptr() if item_order is not None else None
    with torch.cuda.device(value.device):
        _lib.call("dvis_msda_pack_pairs", value.data_ptr(), N, S, M, D, pairs.data_ptr(), _stream())
        _lib.call("dvis_msda_pair_forward", pairs.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                  offsets.data_ptr(), offsets.stride(1), logits.data_ptr(), logits.stride(1), _DTYPE[offsets.dtype],
                  reference_points.data_ptr(), reference_points.shape[-1], N, S, M, D, num_levels, Lq, num_points,
                  order_ptr, out.data_ptr(), _stream())
    return out


def install_as_reference_extension():
    """Register this module's op functions as the importable module `MultiScaleDeformableAttention`
    (OPS/setup.py:60) so the reference's `import MultiScaleDeformableAttention as MSDA` binds to them."""
    m = types.ModuleType("MultiScaleDeformableAttention")
    m.ms_deform_attn_forward = lambda v, s, l, loc, a, step: ms_deform_attn_forward(v, s, l, loc, a, step)
    m.ms_deform_attn_backward = ms_deform_attn_backward
    m.__doc__ = "dvis_plus_b200 drop-in for the reference's MultiScaleDeformableAttention extension"
    sys.modules["MultiScaleDeformableAttention"] = m
    return m


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step):
    """-> [grad_value, grad_sampling_loc, grad_attn_weight]  (OPS/src/vision.cpp:20, ms_deform_attn_cuda.cu:88-158)."""
    _check_cuda_contig(value=value, spatial_shapes=spatial_shapes, level_start_index=level_start_index,
                       sampling_loc=sampling_loc, attn_weight=attn_weight, grad_output=grad_output)
    N, S, M, D = value.shape
    L = spatial_shapes.shape[0]
    Lq, P = sampling_loc.shape[1], sampling_loc.shape[4]
    step = min(N, int(im2col_step))
    if step <= 0:
        raise RuntimeError(f"im2col_step({im2col_step}) and batch({N}) must be positive")
    if N % step != 0:
        raise RuntimeError(f"batch({N}) must divide im2col_step({step})")
    if value.dtype not in (torch.float32, torch.float64):
        raise RuntimeError(f'"ms_deform_attn_backward_cuda" not implemented for \'{value.dtype}\'')
    grad_value = torch.zeros_like(value)
    grad_loc = torch.zeros_like(sampling_loc)
    grad_attn = torch.zeros_like(attn_weight)
    if value.numel() and sampling_loc.numel():
        with torch.cuda.device(value.device):
            _lib.call("dvis_msda_backward", value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                      sampling_loc.data_ptr(), attn_weight.data_ptr(), grad_output.data_ptr(), N, S, M, D, L, Lq, P,
                      _DTYPE[value.dtype], grad_value.data_ptr(), grad_loc.data_ptr(), grad_attn.data_ptr(), _stream())
    return [grad_value, grad_loc, grad_attn]


_MASK_OPERAND = [torch.bfloat16]


def set_mask_operand_dtype(dtype):
    """Operand type of the mask-head GEMMs when the caller does not say: torch.bfloat16 (default: the reference evaluates this einsum
    in half precision under autocast, P/train_net_video.py:259) or torch.float32 = TF32 tensor-core operands (the fp32 tier;
    modules.precision.set_precision("fp32") selects it)."""
    assert dtype in (torch.bfloat16, torch.float32)
    _MASK_OPERAND[0] = dtype


def mask_logits(mask_embed, mask_features, out_dtype=torch.float32, operand_dtype=None):
    """out[b,q,h,w] = sum_c mask_embed[b,q,c] * mask_features[b,c,h,w] on the tcgen05 tensor path (dvis_mask_logits).

    operand_dtype=torch.float32: fp32 operands multiplied as TF32 (dvis_mask_logits_tf32; fp32 output only) -- the fp32 tier;
    None: the process-wide setting of set_mask_operand_dtype (bf16 unless the fp32 tier is selected).

    mask_embed (B,Q,C) any float dtype (cast to bf16; Q*C elements -- negligible); mask_features (B,C,H,W) bf16 in
    torch.channels_last memory format (that is what the pixel decoder drop-in emits); any other layout / dtype is
    converted once here.  Q > 256 is processed in slices of 256 queries.
    Replaces torch.einsum("bqc,bchw->bqhw") (P/dvis_Plus/video_mask2former_transformer_decoder.py:363).
    """
    B, Q, C = mask_embed.shape
    Bf, Cf, H, W = mask_features.shape
    if Bf != B or Cf != C:
        raise RuntimeError(f"mask_logits: shape mismatch {tuple(mask_embed.shape)} vs {tuple(mask_features.shape)}")
    if not mask_features.is_cuda:
        raise RuntimeError("mask_logits: CUDA tensors required (there is no CPU path)")
    feat = mask_features
    if operand_dtype is None:
        operand_dtype = _MASK_OPERAND[0] if out_dtype == torch.float32 else torch.bfloat16
    if operand_dtype == torch.float32:
        if out_dtype != torch.float32:
            raise RuntimeError("mask_logits: TF32 operands produce fp32 logits")
        if feat.dtype != torch.float32 or not feat.is_contiguous(memory_format=torch.channels_last):
            feat = feat.to(dtype=torch.float32, memory_format=torch.channels_last)
        emb = mask_embed.float().contiguous()
        out = torch.empty((B, Q, H, W), dtype=torch.float32, device=feat.device)
        with torch.cuda.device(feat.device):
            _lib.call("dvis_mask_logits_tf32", emb.data_ptr(), feat.data_ptr(), B, Q, C, H * W, out.data_ptr(), _stream())
        return out
    if feat.dtype != torch.bfloat16 or not feat.is_contiguous(memory_format=torch.channels_last):
        feat = feat.to(dtype=torch.bfloat16, memory_format=torch.channels_last)
    emb = mask_embed.to(torch.bfloat16).contiguous()
    out = torch.empty((B, Q, H, W), dtype=out_dtype, device=feat.device)
    with torch.cuda.device(feat.device):
        if Q <= 256:
            _lib.call("dvis_mask_logits", emb.data_ptr(), feat.data_ptr(), B, Q, C, H * W, out.data_ptr(),
                      _DTYPE[out_dtype], _stream())
        else:
            # query slices of <= 256 (a multiple of 8 so the slice starts stay 16-byte aligned), computed in place
            es_in, es_out = emb.element_size(), out.element_size()
            for q0 in range(0, Q, 256):
                q1 = min(Q, q0 + 256)
                _lib.call("dvis_mask_logits_strided", emb.data_ptr() + q0 * C * es_in, Q * C, feat.data_ptr(), B, q1 - q0, C,
                          H * W, out.data_ptr() + q0 * H * W * es_out, Q * H * W, _DTYPE[out_dtype], _stream())
    return out


def mask_logits_clip(mask_embed, mask_features, out_dtype=torch.float32):
    """Mask logits of a clip, written query-major (dvis_mask_logits_clip): mask_embed (T, Q, C), mask_features (T, C, H, W) bf16
    channels_last -> (Q, T, H, W): the "q t h w" layout the meta-architecture keeps, without a transposition pass."""
    T, Q, C = mask_embed.shape
    Tf, Cf, H, W = mask_features.shape
    assert Tf == T and Cf == C and Q <= 256 and mask_features.is_cuda
    feat = mask_features
    if feat.dtype != torch.bfloat16 or not feat.is_contiguous(memory_format=torch.channels_last):
        feat = feat.to(dtype=torch.bfloat16, memory_format=torch.channels_last)
    emb = mask_embed.to(torch.bfloat16).contiguous()
    out = torch.empty((Q, T, H, W), dtype=out_dtype, device=feat.device)
    with torch.cuda.device(feat.device):
        _lib.call("dvis_mask_logits_clip", emb.data_ptr(), feat.data_ptr(), T, Q, C, H * W, out.data_ptr(), _DTYPE[out_dtype], _stream())
    return out


def add_layernorm(x, residual, weight, bias, eps=1e-5, *, want_f32=True, lp_dtype=None, pos=None):
    """LayerNorm(x + residual) in one pass (dvis_add_layernorm).

    x, residual: (..., C) f32|bf16 contiguous (residual may be None).  Returns (y_f32 | None, y_lp | None, y_lp_pos | None)
    where y_lp is a copy in `lp_dtype` and y_lp_pos = y + pos (pos: (rows_p, C) f32 broadcast over leading rows).
    """
    C = x.shape[-1]
    rows = x.numel() // C
    assert x.is_cuda and x.is_contiguous() and (residual is None or (residual.is_contiguous() and residual.shape == x.shape))
    w = weight if weight.dtype == torch.float32 else weight.float()
    b = bias if bias.dtype == torch.float32 else bias.float()
    out_f32 = torch.empty(x.shape, dtype=torch.float32, device=x.device) if want_f32 else None
    out_lp = torch.empty(x.shape, dtype=lp_dtype, device=x.device) if lp_dtype is not None else None
    out_lp_pos = None
    pos_rows = 0
    if pos is not None:
        assert lp_dtype is not None and pos.dtype == torch.float32 and pos.is_contiguous() and pos.shape[-1] == C
        pos_rows = pos.numel() // C
        assert rows % pos_rows == 0
        out_lp_pos = torch.empty(x.shape, dtype=lp_dtype, device=x.device)
    ptr = lambda t: t.data_ptr() if t is not None else None
    with torch.cuda.device(x.device):
        _lib.call("dvis_add_layernorm", x.data_ptr(), _DTYPE[x.dtype], ptr(residual),
                  _DTYPE[residual.dtype] if residual is not None else DVIS_F32, w.data_ptr(), b.data_ptr(), ptr(pos),
                  pos_rows, rows, C, float(eps), ptr(out_f32), ptr(out_lp), ptr(out_lp_pos),
                  _DTYPE[lp_dtype] if lp_dtype is not None else DVIS_F32, _stream())
    return out_f32, out_lp, out_lp_pos


def _check_linear_tc(x, w, bias):
    K = x.shape[-1]
    assert x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.is_contiguous() and w.is_contiguous()
    assert w.shape[1] == K and (bias is None or (bias.dtype == torch.float32 and bias.numel() == w.shape[0]))
    return x.numel() // K, w.shape[0], K


def linear_tc(x, w, bias=None, relu=False):
    """nn.Linear on the tcgen05 path (dvis_linear_tc): x (..., K) bf16, w (N, K) bf16, bias (N,) f32 -> (..., N) bf16."""
    rows, N, K = _check_linear_tc(x, w, bias)
    y = torch.empty(x.shape[:-1] + (N,), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("dvis_linear_tc", x.data_ptr(), K, w.data_ptr(), bias.data_ptr() if bias is not None else None, int(relu),
                  rows, N, K, y.data_ptr(), N, _stream())
    return y


def linear_tc_heads(x, w, bias=None, row_mask=None):
    """MSDeformAttn.value_proj with a head-major result (dvis_linear_tc_heads): x (B, S, K) bf16 -> (B, N/32, S, 32) bf16."""
    rows, N, K = _check_linear_tc(x, w, bias)
    B, S = x.shape[0], x.shape[1]
    assert x.dim() == 3
    if row_mask is not None:
        row_mask = row_mask.reshape(B * S).to(torch.uint8).contiguous()
    out = torch.empty((B, N // 32, S, 32), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("dvis_linear_tc_heads", x.data_ptr(), K, w.data_ptr(), bias.data_ptr() if bias is not None else None, B, S, N, K,
                  row_mask.data_ptr() if row_mask is not None else None, out.data_ptr(), _stream())
    return out


def linear_tc_add_layernorm(x, w, bias, residual, gamma, beta, eps=1e-5, *, want_f32=True, want_lp=True, pos=None):
    """LayerNorm(residual + x . w^T + bias) in one kernel (dvis_linear_tc_add_ln); outputs as add_layernorm:
    (y_f32 | None, y_bf16 | None, (y + pos)_bf16 | None).  residual (..., N) f32; gamma / beta (N,) f32; pos (rows_p, N) f32."""
    rows, N, K = _check_linear_tc(x, w, bias)
    shape = x.shape[:-1] + (N,)
    assert residual.dtype == torch.float32 and residual.is_contiguous() and residual.shape == shape
    assert gamma.dtype == torch.float32 and beta.dtype == torch.float32
    out_f32 = torch.empty(shape, dtype=torch.float32, device=x.device) if want_f32 else None
    out_lp = torch.empty(shape, dtype=torch.bfloat16, device=x.device) if want_lp else None
    out_lp_pos, pos_rows = None, 0
    if pos is not None:
        assert pos.dtype == torch.float32 and pos.is_contiguous() and pos.shape[-1] == N
        pos_rows = pos.numel() // N
        assert rows % pos_rows == 0
        out_lp_pos = torch.empty(shape, dtype=torch.bfloat16, device=x.device)
    ptr = lambda t: t.data_ptr() if t is not None else None
    with torch.cuda.device(x.device):
        _lib.call("dvis_linear_tc_add_ln", x.data_ptr(), K, w.data_ptr(), ptr(bias), residual.data_ptr(), gamma.data_ptr(),
                  beta.data_ptr(), float(eps), rows, N, K, ptr(pos), pos_rows, ptr(out_f32), ptr(out_lp), ptr(out_lp_pos), _stream())
    return out_f32, out_lp, out_lp_pos


def groupnorm_nhwc(x, num_groups, weight, bias, eps=1e-5, *, relu=False, up=None, up_hw=None, hw=None, pos=None,
                   out_f32=None, out_lp=None, out_lp_pos=None):
    """GroupNorm over a channels-last (N, HW, C) map with fused upsample-add / ReLU / casts (dvis_groupnorm_nhwc).

    x: (N, HW, C) f32|bf16, rows contiguous (batch stride free).  `up`: optional (N, h*w, C) f32 low-res map to add after
    bilinear upsampling from `up_hw` = (h, w) to `hw` = (H, W).  Outputs are caller-provided (N, HW, C) views (any batch
    stride, e.g. slices of an (N, S, C) token buffer): out_f32 f32, out_lp / out_lp_pos in their own dtype (same for
    both); out_lp_pos additionally adds pos (HW, C) f32.
    """
    N, HW, C = x.shape
    assert x.is_cuda and x.stride(2) == 1 and x.stride(1) == C
    outs = [o for o in (out_f32, out_lp, out_lp_pos) if o is not None]
    assert outs and all(o.shape == (N, HW, C) and o.stride(2) == 1 and o.stride(1) == C for o in outs)
    obs = outs[0].stride(0)
    assert all(o.stride(0) == obs for o in outs)
    lp = out_lp if out_lp is not None else out_lp_pos
    lp_dtype = lp.dtype if lp is not None else torch.float32
    if out_lp is not None and out_lp_pos is not None:
        assert out_lp.dtype == out_lp_pos.dtype
    ws = torch.empty(2 * N * num_groups * (1 + (HW + 255) // 256), dtype=torch.float64, device=x.device)
    uh = uw = H = W = 0
    if up is not None:
        (uh, uw), (H, W) = up_hw, hw
        assert up.dtype == torch.float32 and up.shape == (N, uh * uw, C) and up.stride(2) == 1 and up.stride(1) == C
    if pos is not None:
        assert pos.dtype == torch.float32 and pos.shape == (HW, C) and pos.is_contiguous()
    w = weight if weight.dtype == torch.float32 else weight.float()
    b = bias if bias.dtype == torch.float32 else bias.float()
    ptr = lambda t: t.data_ptr() if t is not None else None
    with torch.cuda.device(x.device):
        _lib.call("dvis_groupnorm_nhwc", x.data_ptr(), _DTYPE[x.dtype], x.stride(0), N, HW, C, num_groups, w.data_ptr(),
                  b.data_ptr(), float(eps), int(relu), ws.data_ptr(), ptr(up), up.stride(0) if up is not None else 0,
                  uh, uw, H, W, ptr(pos), ptr(out_f32), ptr(out_lp), ptr(out_lp_pos), _DTYPE[lp_dtype], obs, _stream())


def lap_chain(cost, idx_init=None):
    """Batched Hungarian matching + index chain on the device (dvis_lap_chain).
    cost (T, n, n) f32 with rows = reference items, cols = current items -> (sigma (T,n), idx (T,n)) int64."""
    assert cost.is_cuda and cost.dtype == torch.float32 and cost.dim() == 3 and cost.shape[1] == cost.shape[2]
    cost = cost.contiguous()
    T, n, _ = cost.shape
    sigma = torch.empty((T, n), dtype=torch.int64, device=cost.device)
    idx = torch.empty((T, n), dtype=torch.int64, device=cost.device)
    if idx_init is not None:
        assert idx_init.dtype == torch.int64 and idx_init.numel() == n and idx_init.is_cuda
        idx_init = idx_init.contiguous()
    with torch.cuda.device(cost.device):
        _lib.call("dvis_lap_chain", cost.data_ptr(), T, n, idx_init.data_ptr() if idx_init is not None else None,
                  sigma.data_ptr(), idx.data_ptr(), _stream())
    return sigma, idx


def lap_rect(cost):
    """Rectangular Hungarian matching on the device (dvis_lap_rect): cost (rows, cols) or (B, rows, cols) f32 ->
    row_to_col int64 of shape (rows,) / (B, rows): the assigned column per row, -1 where a row stays unmatched."""
    assert cost.is_cuda and cost.dtype == torch.float32 and cost.dim() in (2, 3)
    c = cost.contiguous()
    B = 1 if c.dim() == 2 else c.shape[0]
    rows, cols = c.shape[-2:]
    out = torch.empty(c.shape[:-1], dtype=torch.int64, device=c.device)
    with torch.cuda.device(c.device):
        _lib.call("dvis_lap_rect", c.data_ptr(), B, rows, cols, out.data_ptr(), _stream())
    return out


def resize_bilinear_nhwc(x, size):
    """F.interpolate(x, size, mode="bilinear", align_corners=False) for a bf16 channels_last (N, C, h, w) map; returns a
    bf16 channels_last (N, C, H, W) map (dvis_resize_bilinear_nhwc)."""
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous(memory_format=torch.channels_last)
    N, C, h, w = x.shape
    H, W = int(size[0]), int(size[1])
    out = torch.empty((N, C, H, W), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    with torch.cuda.device(x.device):
        _lib.call("dvis_resize_bilinear_nhwc", x.data_ptr(), N, h, w, C, out.data_ptr(), H, W, _stream())
    return out


def attn_bias_from_logits(logits, dtype=torch.bfloat16):
    """(..., hw) f32 mask logits -> additive attention bias of the same shape: -inf where sigmoid(logit) < 0.5, rows that
    would be fully masked reset to 0 (dvis_attn_bias_from_logits)."""
    assert logits.is_cuda and logits.dtype == torch.float32 and logits.is_contiguous()
    hw = logits.shape[-1]
    bias = torch.empty(logits.shape, dtype=dtype, device=logits.device)
    with torch.cuda.device(logits.device):
        _lib.call("dvis_attn_bias_from_logits", logits.data_ptr(), logits.numel() // hw, hw, bias.data_ptr(), _DTYPE[dtype], _stream())
    return bias


def mask_attn_bias(mask_embed, level_features, dtype=torch.bfloat16):
    """Attention bias of the masked-attention decoder straight from the tcgen05 mask GEMM (dvis_mask_attn_bias):
    mask_embed (B,Q,C); level_features (B,C,h,w) bf16 channels_last = mask_features resized to the attention level.
    -> (B, Q, h*w) `dtype`: -inf where sigmoid(E @ F) < 0.5, rows that would be fully masked reset to 0."""
    B, Q, C = mask_embed.shape
    _, _, h, w = level_features.shape
    assert level_features.is_contiguous(memory_format=torch.channels_last)
    if level_features.dtype == torch.float32:                      # fp32 tier: TF32 operands, fp32 bias
        assert dtype == torch.float32
        emb = mask_embed.float().contiguous()
        bias = torch.empty((B, Q, h * w), dtype=torch.float32, device=emb.device)
        ws = torch.empty(B * Q, dtype=torch.int32, device=emb.device)
        with torch.cuda.device(emb.device):
            _lib.call("dvis_mask_attn_bias_tf32", emb.data_ptr(), level_features.data_ptr(), B, Q, C, h * w, bias.data_ptr(),
                      ws.data_ptr(), _stream())
        return bias
    assert level_features.dtype == torch.bfloat16
    emb = mask_embed.to(torch.bfloat16).contiguous()
    if Q > 256:
        return attn_bias_from_logits(mask_logits(mask_embed, level_features, torch.float32).flatten(2), dtype)
    bias = torch.empty((B, Q, h * w), dtype=dtype, device=emb.device)
    ws = torch.empty(B * Q, dtype=torch.int32, device=emb.device)
    with torch.cuda.device(emb.device):
        _lib.call("dvis_mask_attn_bias", emb.data_ptr(), level_features.data_ptr(), B, Q, C, h * w, bias.data_ptr(),
                  _DTYPE[dtype], ws.data_ptr(), _stream())
    return bias


def level_tokens(x, level_embed, pos=None):
    """The decoder's memory of one attention level in one pass (dvis_level_tokens): x (B, C, h, w) f32|bf16 with
    channels-last memory (any batch stride), level_embed (C,) f32, pos (h*w, C) f32 or None.
    -> (tok (B, h*w, C) bf16 = x + level_embed,  key (B, h*w, C) bf16 = tok + pos  | None)."""
    B, C, h, w = x.shape
    if not x.is_cuda:
        raise RuntimeError("level_tokens: CUDA tensors required (there is no CPU path)")
    assert x.stride(1) == 1 and x.stride(3) == C and x.stride(2) == w * C, "channels-last rows expected"
    assert x.dtype in (torch.float32, torch.bfloat16) and level_embed.dtype == torch.float32 and level_embed.is_contiguous()
    tok = torch.empty((B, h * w, C), dtype=torch.bfloat16, device=x.device)
    key = None
    if pos is not None:
        assert pos.dtype == torch.float32 and pos.is_contiguous() and pos.shape == (h * w, C)
        key = torch.empty((B, h * w, C), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("dvis_level_tokens", x.data_ptr(), _DTYPE[x.dtype], x.stride(0), level_embed.data_ptr(),
                  pos.data_ptr() if pos is not None else None, B, h * w, C, tok.data_ptr(), key.data_ptr() if key is not None else None,
                  _stream())
    return tok, key


def mask_attn_bits(mask_embed, level_features):
    """The masked-attention decoder's attention mask as packed bits, straight from the tcgen05 mask GEMM
    (dvis_mask_attn_bits): mask_embed (B,Q,C); level_features (B,C,h,w) bf16 channels_last.
    -> uint8 (B, Q, ceil(h*w / 64) * 8): bit (p % 8) of byte (p / 8) set where sigmoid(E @ F)[p] < 0.5 (query may not attend
    pixel p); rows that would be fully masked are cleared.  Feed to `flash_attn(..., mask_bits=)`."""
    B, Q, C = mask_embed.shape
    _, _, h, w = level_features.shape
    if not level_features.is_cuda:
        raise RuntimeError("mask_attn_bits: CUDA tensors required (there is no CPU path)")
    assert level_features.dtype == torch.bfloat16 and level_features.is_contiguous(memory_format=torch.channels_last)
    emb = mask_embed.to(torch.bfloat16).contiguous()
    row = (h * w + 63) // 64 * 8
    bits = torch.empty((B, Q, row), dtype=torch.uint8, device=emb.device)
    ws = torch.empty(B * Q, dtype=torch.int32, device=emb.device)
    with torch.cuda.device(emb.device):
        _lib.call("dvis_mask_attn_bits", emb.data_ptr(), level_features.data_ptr(), B, Q, C, h * w, bits.data_ptr(), row,
                  ws.data_ptr(), _stream())
    return bits


def flash_attn(q, k, v, scale, mask_bits=None, out=None):
    """softmax(scale * q k^T [masked]) v on the warp tensor cores (dvis_flash_attn).

    q (B, Lq, H, Dh), k / v (B, Lk, H, Dh): bf16 views with a contiguous head dim (any row / batch / head stride, e.g.
    slices of a packed QKV projection).  mask_bits: optional uint8 (B, Lq, row_bytes) with row_bytes % 8 == 0 and
    row_bytes >= ceil(Lk / 64) * 8: bit j set = key j masked for that query (all heads), see `ops.mask_attn_bits`.
    out: optional (B, Lq, H*Dh) bf16 view to write into (any batch / row stride, e.g. a transposed buffer).
    Returns `out` or a new contiguous (B, Lq, H*Dh) bf16 tensor."""
    B, Lq, H, Dh = q.shape
    Lk = k.shape[1]
    for t in (q, k, v):
        if not t.is_cuda:
            raise RuntimeError("flash_attn: CUDA tensors required (there is no CPU path)")
        assert t.dtype == torch.bfloat16 and t.stride(3) == 1, "bf16 with a contiguous head dim"
    assert k.shape == (B, Lk, H, Dh) and v.shape == (B, Lk, H, Dh)
    if out is None:
        out = torch.empty((B, Lq, H * Dh), dtype=torch.bfloat16, device=q.device)
    assert out.shape == (B, Lq, H * Dh) and out.dtype == torch.bfloat16 and out.stride(2) == 1
    mrow = mbatch = 0
    if mask_bits is not None:
        assert mask_bits.dtype == torch.uint8 and mask_bits.is_cuda and mask_bits.dim() == 3 and mask_bits.stride(2) == 1
        assert mask_bits.shape[0] == B and mask_bits.shape[1] == Lq
        mrow, mbatch = mask_bits.stride(1), mask_bits.stride(0)
    with torch.cuda.device(q.device):
        _lib.call("dvis_flash_attn", q.data_ptr(), q.stride(1), q.stride(0), q.stride(2), k.data_ptr(), k.stride(1), k.stride(0),
                  k.stride(2), v.data_ptr(), v.stride(1), v.stride(0), v.stride(2), out.data_ptr(), out.stride(1), out.stride(0),
                  mask_bits.data_ptr() if mask_bits is not None else None, mrow, mbatch, B, Lq, Lk, H, Dh, float(scale), _stream())
    return out


def set_pdl(enabled):
    """Programmatic dependent launch for the temporal-stage kernels (dvis_set_pdl)."""
    _lib.lib().dvis_set_pdl(int(bool(enabled)))


def linear_small(w, bias=None, *, x=None, src0=None, ln0=None, src1=None, ln1=None, eps=1e-5, want_side0=False,
                 want_side1=False, residual=None, relu=False, out_f32=False, out_bf16=True, taps=1, tap_pad=0,
                 tap_period=0, tap_len=0):
    """One dependent step of a tracker / refiner block (dvis_linear_small):  Y = act(A @ w^T + bias) [+ residual].

    w (N, K) or (B, N, K) bf16; bias (N,) / (B, N) f32 or None.
    A = x: bf16 (M, K/taps) or (B, M, K/taps), last dim contiguous (taps > 1: Conv1d over time as a GEMM, see the header), or
    A = LN1(LN0(src0) + src1) built in the kernel: src0 (M, K) f32, ln0 / ln1 = (gamma, beta, ...) f32 or None, src1 (M, K)
    f32 | bf16 or None.  residual (M, N) f32.  -> (y_f32 | None, y_bf16 | None, side0 | None, side1 | None)."""
    assert w.is_cuda and w.dtype == torch.bfloat16 and w.is_contiguous()
    batched = w.dim() == 3
    B = w.shape[0] if batched else 1
    N, K = w.shape[-2], w.shape[-1]
    dev = w.device
    ptr = lambda t: t.data_ptr() if t is not None else None
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.shape[-1] == N
    if x is not None:
        if not x.is_cuda:
            raise RuntimeError("linear_small: CUDA tensors required (there is no CPU path)")
        assert x.dtype == torch.bfloat16 and x.stride(-1) == 1 and x.shape[-1] * taps == K
        M = x.shape[-2]
        ldx, xb = x.stride(-2), (x.stride(0) if batched else 0)
    else:
        assert not batched and src0.dtype == torch.float32 and src0.is_contiguous() and src0.shape[-1] == K
        M = src0.numel() // K
        ldx = xb = 0
        if src1 is not None:
            assert src1.is_contiguous() and src1.numel() == src0.numel() and src1.dtype in (torch.float32, torch.bfloat16)
    for ln in (ln0, ln1):
        assert ln is None or (ln[0].dtype == torch.float32 and ln[1].dtype == torch.float32 and ln[0].numel() == K)
    shape = (B, M, N) if batched else (M, N)
    y32 = torch.empty(shape, dtype=torch.float32, device=dev) if out_f32 else None
    y16 = torch.empty(shape, dtype=torch.bfloat16, device=dev) if out_bf16 else None
    side0 = torch.empty((M, K), dtype=torch.float32, device=dev) if want_side0 else None
    side1 = torch.empty((M, K), dtype=torch.float32, device=dev) if want_side1 else None
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.stride(-1) == 1 and residual.numel() == M * N and not batched
    ws = cnt = None
    if x is not None and not batched and taps == 1 and K >= 1024 and M <= 512:        # split-K scratch (see the header)
        ws = torch.empty(4 * M * N, dtype=torch.float32, device=dev)
        cnt = _splitk_counters(dev, ((M + 31) // 32) * ((N + 31) // 32))
    with torch.cuda.device(dev):
        _lib.call("dvis_linear_small", ptr(x), ldx, xb, taps, tap_pad, tap_period, tap_len, ptr(src0),
                  ptr(ln0[0]) if ln0 else None, ptr(ln0[1]) if ln0 else None, ptr(src1),
                  _DTYPE[src1.dtype] if src1 is not None else DVIS_F32, ptr(ln1[0]) if ln1 else None,
                  ptr(ln1[1]) if ln1 else None, float(eps), ptr(side0), ptr(side1), w.data_ptr(), N * K if batched else 0,
                  ptr(bias), N if (batched and bias is not None and bias.dim() == 2) else 0, ptr(residual),
                  residual.stride(-2) if residual is not None else 0, int(relu), ptr(y32), ptr(y16), N, M * N, B, M, N, K,
                  ptr(ws), ptr(cnt), _stream())
    return y32, y16, side0, side1


def linear_small_ln(w, bias, x, residual, ln, *, src1=None, ln2=None, relu=False, eps=1e-5, want_f32=True):
    """Linear step with the post-norm block's LayerNorm(s) in the producer's epilogue (dvis_linear_small_ln):
        y = act(x @ w^T + bias) + residual;  e1 = LN(y; ln);  e2 = LN(e1 + src1; ln2)  (when ln2 is given)
    x (M, K) bf16 (last dim contiguous), w (N, K) bf16, residual (M, N) f32 | None, ln / ln2 = (gamma, beta) f32.
    -> (e1_f32 | None, e1_bf16, e2_f32 | None, e2_bf16 | None)."""
    if not x.is_cuda:
        raise RuntimeError("linear_small_ln: CUDA tensors required (there is no CPU path)")
    assert w.dtype == torch.bfloat16 and w.is_contiguous() and w.dim() == 2 and x.dtype == torch.bfloat16 and x.stride(-1) == 1
    N, K = w.shape
    M = x.shape[0]
    dev = w.device
    ptr = lambda t: t.data_ptr() if t is not None else None
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.stride(-1) == 1 and residual.shape == (M, N)
    if src1 is not None:
        assert src1.is_contiguous() and src1.shape == (M, N) and src1.dtype in (torch.float32, torch.bfloat16)
    y = torch.empty((M, N), dtype=torch.float32, device=dev)
    e1_32 = torch.empty((M, N), dtype=torch.float32, device=dev) if want_f32 else None
    e1_16 = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
    e2_32 = torch.empty((M, N), dtype=torch.float32, device=dev) if (ln2 is not None and want_f32) else None
    e2_16 = torch.empty((M, N), dtype=torch.bfloat16, device=dev) if ln2 is not None else None
    ws = torch.empty(4 * M * N, dtype=torch.float32, device=dev) if K >= 1024 else None
    cnt = _splitk_counters(dev, 2 * (((M + 31) // 32) * ((N + 31) // 32) + (M + 31) // 32))
    tiles = ((M + 31) // 32) * ((N + 31) // 32)
    with torch.cuda.device(dev):
        _lib.call("dvis_linear_small_ln", x.data_ptr(), x.stride(0), w.data_ptr(), ptr(bias), ptr(residual),
                  residual.stride(0) if residual is not None else 0, int(relu), M, N, K, y.data_ptr(), ln[0].data_ptr(), ln[1].data_ptr(),
                  float(eps), ptr(src1), _DTYPE[src1.dtype] if src1 is not None else DVIS_F32, ptr(ln2[0]) if ln2 else None,
                  ptr(ln2[1]) if ln2 else None, ptr(e1_32), e1_16.data_ptr(), ptr(e2_32), ptr(e2_16), ptr(ws),
                  cnt.data_ptr() if ws is not None else None, cnt.data_ptr() + 4 * tiles, _stream())
    return e1_32, e1_16, e2_32, e2_16


_SPLITK_CNT = {}


def _splitk_counters(dev, n):
    """Arrival counters of the split-K tiles, one buffer per device: zeroed once, left zero by every launch.  Created by the
    first eager call (the runners warm up before they capture), so CUDA-graph captures only ever see an existing buffer.
    Split-K launches of one process run on one stream at a time (the temporal stage is a single dependent chain); concurrent
    use from two streams would need a buffer per stream."""
    key = str(dev)
    c = _SPLITK_CNT.get(key)
    if c is None or c.numel() < n:
        c = torch.zeros(max(n, 4096), dtype=torch.int32, device=dev)
        _SPLITK_CNT[key] = c
    return c


# ---- video post-processing (csrc/postproc.cu; P/dvis_Plus/meta_architecture.py:818-979) --------------------------------

def _mask_view(pred_masks):
    """(Q, T, h, w) f32|bf16 view with contiguous (h, w) planes -> (tensor, q_stride, t_stride); copies only if the
    planes themselves are strided."""
    if not pred_masks.is_cuda:
        raise RuntimeError("post-processing: CUDA tensors required (there is no CPU path)")
    if pred_masks.dim() != 4:
        raise RuntimeError(f"post-processing: pred_masks must be (Q, T, h, w), got {tuple(pred_masks.shape)}")
    if pred_masks.dtype not in (torch.float32, torch.bfloat16):
        pred_masks = pred_masks.float()
    if pred_masks.stride(3) != 1 or pred_masks.stride(2) != pred_masks.shape[3]:
        pred_masks = pred_masks.contiguous()
    return pred_masks, pred_masks.stride(0), pred_masks.stride(1)


def _geom_args(pred_masks, first_resize_size, img_size, out_size):
    h, w = pred_masks.shape[-2:]
    return [int(v) for v in (h, w, first_resize_size[0], first_resize_size[1], img_size[0], img_size[1], out_size[0], out_size[1])]


def class_scores(pred_cls, aux_pred_cls=None):
    """(Q, K+1) logits -> (Q, K+1) f32 softmax scores, the K object columns max'ed with softmax(aux) (dvis_class_scores)."""
    cls = pred_cls.float().contiguous()
    aux = aux_pred_cls.float().contiguous() if aux_pred_cls is not None else None
    if not cls.is_cuda:
        raise RuntimeError("class_scores: CUDA tensors required (there is no CPU path)")
    Q, K1 = cls.shape
    scores = torch.empty_like(cls)
    with torch.cuda.device(cls.device):
        _lib.call("dvis_class_scores", cls.data_ptr(), aux.data_ptr() if aux is not None else None, Q, K1, scores.data_ptr(),
                  _stream())
    return scores


def vis_topk(pred_cls, max_num, aux_pred_cls=None):
    """Top-`max_num` (score, label, query) triples of softmax(pred_cls)[:, :-1] (max'ed with aux), score descending
    (dvis_vis_topk).  -> (scores f32, labels i64, query indices i64), all (max_num,) on the device."""
    cls = pred_cls.float().contiguous()
    aux = aux_pred_cls.float().contiguous() if aux_pred_cls is not None else None
    if not cls.is_cuda:
        raise RuntimeError("vis_topk: CUDA tensors required (there is no CPU path)")
    Q, K1 = cls.shape
    ws = torch.empty((Q, K1), dtype=torch.float32, device=cls.device)
    scores = torch.empty(max_num, dtype=torch.float32, device=cls.device)
    labels = torch.empty(max_num, dtype=torch.int64, device=cls.device)
    query = torch.empty(max_num, dtype=torch.int64, device=cls.device)
    with torch.cuda.device(cls.device):
        _lib.call("dvis_vis_topk", cls.data_ptr(), aux.data_ptr() if aux is not None else None, Q, K1, int(max_num),
                  ws.data_ptr(), scores.data_ptr(), labels.data_ptr(), query.data_ptr(), _stream())
    return scores, labels, query


def vis_masks(pred_masks, sel, first_resize_size, img_size, out_size, out=None, packed=False):
    """bool masks (n, T, Ho, Wo) = resize_chain(pred_masks[sel]) > 0 straight from the stride-4 logits (dvis_vis_masks).
    pred_masks (Q, T, h, w) f32|bf16 view (any query / frame strides); sel (n,) int64 device tensor or None (all).
    packed=True: one bit per pixel instead, uint8 (n, T, Ho, ceil(Wo/8)), little bit order (dvis_vis_masks_packed; see
    unpack_masks) -- 8x less to write and to copy to the host."""
    m, qs, ts = _mask_view(pred_masks)
    n = m.shape[0] if sel is None else sel.numel()
    T = m.shape[1]
    Ho, Wo = int(out_size[0]), int(out_size[1])
    row = (Wo + 7) // 8 if packed else Wo
    if out is None:
        out = torch.empty((n, T, Ho, row), dtype=torch.uint8 if packed else torch.bool, device=m.device)
    assert out.dtype == (torch.uint8 if packed else torch.bool) and out.is_contiguous() and out.shape == (n, T, Ho, row)
    if n == 0:
        return out
    if sel is not None:
        assert sel.dtype == torch.int64 and sel.is_cuda
        sel = sel.contiguous()
    geom = _geom_args(m, first_resize_size, img_size, out_size)
    per_call = max(1, 65535 // T)                     # grid.z limit: n_sel * frames <= 65535 per launch
    entry = "dvis_vis_masks_packed" if packed else "dvis_vis_masks"
    with torch.cuda.device(m.device):
        for n0 in range(0, n, per_call):
            n1 = min(n, n0 + per_call)
            if sel is not None:
                base, sel_ptr = m.data_ptr(), sel.data_ptr() + 8 * n0
            else:
                base, sel_ptr = m.data_ptr() + n0 * qs * m.element_size(), None
            _lib.call(entry, base, _DTYPE[m.dtype], qs, ts, sel_ptr, n1 - n0, T, *geom, out.data_ptr() + n0 * T * Ho * row, _stream())
    return out


def unpack_masks(packed, width):
    """Host side of vis_masks(packed=True): (..., Ho, ceil(Wo/8)) uint8 tensor -> (..., Ho, Wo) bool CPU tensor."""
    import numpy as np
    assert packed.dtype == torch.uint8
    bits = np.unpackbits(packed.cpu().numpy(), axis=-1, bitorder="little")[..., :width]
    return torch.from_numpy(bits).bool()


def vps_argmax(pred_masks, keep_idx, keep_score, first_resize_size, img_size, out_size):
    """Per-pixel arg-max of keep_score[k] * sigmoid(resize_chain(pred_masks[keep_idx[k]])) + the segment-filter counts
    (dvis_vps_argmax).  -> win (T, Ho, Wo) int32 (k, or ~k if the winner's probability < 0.5), areas (3, n_keep) int64."""
    m, qs, ts = _mask_view(pred_masks)
    T = m.shape[1]
    n = keep_idx.numel()
    assert keep_idx.dtype == torch.int64 and keep_idx.is_cuda and keep_score.dtype == torch.float32 and keep_score.numel() == n
    Ho, Wo = int(out_size[0]), int(out_size[1])
    win = torch.empty((T, Ho, Wo), dtype=torch.int32, device=m.device)
    areas = torch.empty((3, n), dtype=torch.int64, device=m.device)
    with torch.cuda.device(m.device):
        _lib.call("dvis_vps_argmax", m.data_ptr(), _DTYPE[m.dtype], qs, ts, keep_idx.contiguous().data_ptr(),
                  keep_score.contiguous().data_ptr(), n, T, *_geom_args(m, first_resize_size, img_size, out_size),
                  win.data_ptr(), areas.data_ptr(), _stream())
    return win, areas


def vps_paint(win, seg_of_k):
    """panoptic = seg_of_k[win] where win >= 0 else 0 (dvis_vps_paint); seg_of_k (n_keep,) int32 on the device."""
    assert win.dtype == torch.int32 and win.is_contiguous() and seg_of_k.dtype == torch.int32 and seg_of_k.is_cuda
    out = torch.empty_like(win)
    with torch.cuda.device(win.device):
        _lib.call("dvis_vps_paint", win.data_ptr(), seg_of_k.contiguous().data_ptr(), win.numel(), out.data_ptr(), _stream())
    return out


def vss_argmax(pred_masks, mask_cls, first_resize_size, img_size, out_size):
    """argmax_c sum_q mask_cls[q, c] * sigmoid(resize_chain(pred_masks[q])) -> (T, Ho, Wo) int64 (dvis_vss_argmax).
    mask_cls (Q, K) f32 with unit column stride (e.g. class_scores(...)[:, :-1])."""
    m, qs, ts = _mask_view(pred_masks)
    Q, T = m.shape[:2]
    assert mask_cls.dtype == torch.float32 and mask_cls.is_cuda and mask_cls.shape[0] == Q and mask_cls.stride(1) == 1
    K = mask_cls.shape[1]
    Ho, Wo = int(out_size[0]), int(out_size[1])
    out = torch.empty((T, Ho, Wo), dtype=torch.int64, device=m.device)
    with torch.cuda.device(m.device):
        _lib.call("dvis_vss_argmax", m.data_ptr(), _DTYPE[m.dtype], qs, ts, mask_cls.data_ptr(), mask_cls.stride(0), Q, K, T,
                  *_geom_args(m, first_resize_size, img_size, out_size), out.data_ptr(), _stream())
    return out
