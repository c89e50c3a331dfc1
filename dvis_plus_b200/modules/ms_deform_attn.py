"""Drop-in for OPS/modules/ms_deform_attn.py and OPS/functions/ms_deform_attn_func.py.

Same constructor signature, attributes (`im2col_step`, `d_model`, `n_levels`, `n_heads`, `n_points`), parameter
names (`sampling_offsets`, `attention_weights`, `value_proj`, `output_proj`) and `_reset_parameters` as the
reference (OPS/modules/ms_deform_attn.py:35-80), so reference checkpoints load unchanged.

Differences, all deliberate:
  * no bare `except` that silently switches to a PyTorch implementation (py:116-121): a failing kernel raises;
  * CPU tensors raise RuntimeError("Not implemented on the CPU") like the reference extension itself does;
  * without autograd the forward uses the fused kernel (softmax + location arithmetic + gather in one pass).
"""
import math
import warnings

import torch
import torch.nn.functional as F
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.init import constant_, xavier_uniform_

from .. import ops
from ..locality import tiled_item_order
from .precision import gemm_dtype


class MSDeformAttnFunction(Function):
    """OPS/functions/ms_deform_attn_func.py:32-49."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        output = ops.ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                                            attention_weights, ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, loc, attn = ctx.saved_tensors
        gv, gl, ga = ops.ms_deform_attn_backward(value, shapes, lsi, loc, attn, grad_output.contiguous(), ctx.im2col_step)
        return gv, None, None, gl, ga, None


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


class MSDeformAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4, ratio=1.0):
        # `ratio` is accepted and ignored exactly like the reference (py:35; ViT-Adapter passes deform_ratio)
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError("d_model must be divisible by n_heads, but got {} and {}".format(d_model, n_heads))
        if not _is_power_of_2(d_model // n_heads):
            warnings.warn("You'd better set d_model in MSDeformAttn to make the dimension of each attention head a power of 2 "
                          "which is more efficient in our CUDA implementation.")
        self.im2col_step = 128
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()
        self._fused_cache = None
        self.use_tc_linear = True      # bf16 inference: value_proj on csrc/linear_tc.cu (head-major epilogue) + head-major gather
        # output_proj + residual + norm1 in one tcgen05 kernel (dvis_linear_tc_add_ln): correct and tested, but measured 226 us
        # against 196 us for cuBLAS + add_layernorm at 16 x 19 320 tokens (profiles/r2_linear_tc.md) -- opt-in until it wins
        self.fuse_output_norm = False

    def _reset_parameters(self):
        # py:66-80: zero offset weights, ring-shaped offset bias scaled by the point index, zero attention
        # weights, Xavier value / output projections
        constant_(self.sampling_offsets.weight.data, 0.)
        thetas = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        grid_init = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid_init = (grid_init / grid_init.abs().max(-1, keepdim=True)[0]).view(self.n_heads, 1, 1, 2)
        grid_init = grid_init.repeat(1, self.n_levels, self.n_points, 1)
        for i in range(self.n_points):
            grid_init[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(grid_init.view(-1))
        constant_(self.attention_weights.weight.data, 0.)
        constant_(self.attention_weights.bias.data, 0.)
        xavier_uniform_(self.value_proj.weight.data)
        constant_(self.value_proj.bias.data, 0.)
        xavier_uniform_(self.output_proj.weight.data)
        constant_(self.output_proj.bias.data, 0.)
        self._fused_cache = None

    # -- inference fast path ----------------------------------------------------------------------
    def _fused_weights(self, dtype):
        """[sampling_offsets ; attention_weights] as ONE linear so offsets and logits come out of a single GEMM."""
        so, aw = self.sampling_offsets, self.attention_weights
        key = (dtype, so.weight._version, aw.weight._version, so.bias._version, aw.bias._version, so.weight.device,
               self.value_proj.weight._version, self.output_proj.weight._version)
        if self._fused_cache is None or self._fused_cache[0] != key:
            w = torch.cat([so.weight, aw.weight], 0).detach().to(dtype).contiguous()
            b = torch.cat([so.bias, aw.bias], 0).detach().to(dtype).contiguous()
            vw, vb = self.value_proj.weight.detach().to(dtype), self.value_proj.bias.detach().to(dtype)
            ow, ob = self.output_proj.weight.detach().to(dtype), self.output_proj.bias.detach().to(dtype)
            self._fused_cache = (key, w, b, vw, vb, ow, ob)
        return self._fused_cache[1:]

    def tc_path_ok(self, dt, S):
        """The tcgen05 projections + head-major gather are built for the production shape: bf16, 32 channels per head,
        d_model <= 256 (csrc/linear_tc.cu, dvis_msda_fused_forward_hm)."""
        return (self.use_tc_linear and dt == torch.bfloat16 and self.d_model // self.n_heads == 32 and self.d_model <= 256
                and self.d_model % 64 == 0 and S >= 2 and 2 <= self.n_levels * self.n_points <= 32)

    def forward_core(self, query, reference_points, input_flatten, host_spatial_shapes, spatial_shapes_dev, level_start_dev,
                     input_padding_mask=None):
        """Everything of the inference path up to (not including) output_proj: (N, Lq, C) in the GEMM dtype."""
        N, Lq, _ = query.shape
        S = input_flatten.shape[1]
        dt = query.dtype
        w, b, vw, vb, ow, ob = self._fused_weights(dt)
        ol = F.linear(query, w, b)                                   # (N, Lq, M*L*P*3): [offsets | logits]
        n_off = self.n_heads * self.n_levels * self.n_points * 2
        order = None
        if Lq == S and sum(h * w_ for h, w_ in host_spatial_shapes) == S:
            order = tiled_item_order(host_spatial_shapes, self.n_heads, query.device)
        ref = reference_points.float().contiguous()
        if self.tc_path_ok(dt, S) and all(w_ >= 2 for _, w_ in host_spatial_shapes):
            # value projection on tcgen05 with the head-major epilogue: (N, M, S, 32), x-adjacent corners in one line
            value_hm = ops.linear_tc_heads(input_flatten, vw, self.value_proj.bias.detach().float(), row_mask=input_padding_mask)
            return ops.msda_fused_forward_hm(value_hm, spatial_shapes_dev, level_start_dev, ol[..., :n_off], ol[..., n_off:], ref,
                                             self.n_levels, self.n_points, item_order=order)
        value = F.linear(input_flatten, vw, vb)
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None], 0.0)
        value = value.view(N, S, self.n_heads, self.d_model // self.n_heads)
        return ops.msda_fused_forward(value, spatial_shapes_dev, level_start_dev, ol[..., :n_off], ol[..., n_off:], ref,
                                      self.n_heads, self.n_levels, self.n_points, item_order=order, out_dtype=dt)

    def forward_fused(self, query, reference_points, input_flatten, host_spatial_shapes, spatial_shapes_dev,
                      level_start_dev, input_padding_mask=None, out_dtype=None):
        """Inference path: query / input_flatten (N, L, C) in the GEMM dtype; returns output_proj(...) in `out_dtype`."""
        core = self.forward_core(query, reference_points, input_flatten, host_spatial_shapes, spatial_shapes_dev, level_start_dev,
                                 input_padding_mask)
        _, _, _, _, ow, ob = self._fused_weights(query.dtype)
        out = F.linear(core, ow, ob)
        return out if out_dtype is None else out.to(out_dtype)

    # -- reference-compatible forward ---------------------------------------------------------------
    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None):
        """Same contract as the reference forward (py:82-125).  (N, Len_q, C) -> (N, Len_q, C)."""
        N, Len_q, _ = query.shape
        N, Len_in, _ = input_flatten.shape
        if not query.is_cuda:
            raise RuntimeError("Not implemented on the CPU")      # OPS/src/ms_deform_attn.h:43
        if reference_points.shape[-1] not in (2, 4):
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(reference_points.shape[-1]))
        needs_grad = torch.is_grad_enabled() and (query.requires_grad or input_flatten.requires_grad or
                                                  any(p.requires_grad for p in self.parameters()))
        if not needs_grad and self.n_points == 4 and (self.d_model // self.n_heads) in (16, 32, 64):
            host_shapes = _host_shapes(input_spatial_shapes)
            assert sum(h * w for h, w in host_shapes) == Len_in
            dt = gemm_dtype()
            return self.forward_fused(query.to(dt), reference_points, input_flatten.to(dt), host_shapes,
                                      input_spatial_shapes, input_level_start_index, input_padding_mask,
                                      out_dtype=query.dtype)
        # autograd / odd-geometry path: the reference's own formulation around the op (py:98-118)
        assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == Len_in
        value = self.value_proj(input_flatten)
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None], float(0))
        value = value.view(N, Len_in, self.n_heads, self.d_model // self.n_heads)
        sampling_offsets = self.sampling_offsets(query).view(N, Len_q, self.n_heads, self.n_levels, self.n_points, 2)
        attention_weights = self.attention_weights(query).view(N, Len_q, self.n_heads, self.n_levels * self.n_points)
        attention_weights = F.softmax(attention_weights, -1).view(N, Len_q, self.n_heads, self.n_levels, self.n_points)
        if reference_points.shape[-1] == 2:
            offset_normalizer = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
            sampling_locations = reference_points[:, :, None, :, None, :] \
                + sampling_offsets / offset_normalizer[None, None, None, :, None, :]
        else:
            sampling_locations = reference_points[:, :, None, :, None, :2] \
                + sampling_offsets / self.n_points * reference_points[:, :, None, :, None, 2:] * 0.5
        output = MSDeformAttnFunction.apply(value.float().contiguous(), input_spatial_shapes, input_level_start_index,
                                            sampling_locations.float().contiguous(), attention_weights.float().contiguous(),
                                            self.im2col_step)
        return self.output_proj(output.to(query.dtype))


_shape_cache = {}


def _host_shapes(spatial_shapes):
    """(L,2) int64 device tensor -> tuple of (H, W) on the host; one device->host copy per distinct tensor."""
    key = (spatial_shapes.data_ptr(), spatial_shapes._version, tuple(spatial_shapes.shape), str(spatial_shapes.device))
    hit = _shape_cache.get(key)
    if hit is not None and hit[0] is spatial_shapes:
        return hit[1]
    shapes = tuple((int(h), int(w)) for h, w in spatial_shapes.tolist())
    if len(_shape_cache) > 64:
        _shape_cache.clear()
    _shape_cache[key] = (spatial_shapes, shapes)
    return shapes
