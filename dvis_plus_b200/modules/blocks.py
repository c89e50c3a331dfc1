"""Attention / FFN building blocks shared by the segmenter predictor, the tracker and the refiner.

Mirrors SelfAttentionLayer, CrossAttentionLayer, FFNLayer, MLP
(P/mask2former_video/modeling/transformer_decoder/video_mask2former_transformer_decoder.py:18-73,76-136,139-179,193-205)
and ReferringCrossAttentionLayer (P/dvis_Plus/tracker.py:8-92) with identical parameter names, so the reference's
state_dicts load unchanged.  `MultiheadAttention` here is parameter-compatible with torch.nn.MultiheadAttention
(in_proj_weight, in_proj_bias, out_proj.weight, out_proj.bias) but always returns the output tensor only.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from .precision import gemm_dtype


def _fast_path(x):
    return x.is_cuda and not torch.is_grad_enabled()


def add_norm(norm: nn.LayerNorm, x, residual):
    """LayerNorm(x + residual) -> fp32; fused kernel on the inference path."""
    if _fast_path(x):
        # device inference: the kernel or an error -- never a silent torch substitute (C % 4 == 0, C <= 2048: every DVIS width)
        to = lambda t: t if t.dtype in (torch.float32, torch.bfloat16) else t.float()
        return ops.add_layernorm(to(x).contiguous(), to(residual).contiguous(), norm.weight, norm.bias, norm.eps)[0]
    return F.layer_norm(x.float() + residual.float(), norm.normalized_shape, norm.weight, norm.bias, norm.eps)


def linear(layer: nn.Linear, x, relu=False):
    """x @ W^T + b in the configured GEMM dtype (inputs rounded, fp32 accumulate); returns that dtype."""
    dt = gemm_dtype() if _fast_path(x) else x.dtype
    w, b = layer.weight, layer.bias
    if w.dtype != dt:
        w, b = _cast_cached(layer, dt)
    x = x.to(dt)
    if relu and b is not None and x.is_cuda and not torch.is_grad_enabled():
        # bias + ReLU in the GEMM epilogue (cuBLASLt RELU_BIAS) instead of a separate elementwise kernel
        return torch._addmm_activation(b, x.reshape(-1, x.shape[-1]), w.t()).view(*x.shape[:-1], w.shape[0])
    y = F.linear(x, w, b)
    return F.relu(y) if relu else y


def _cast_cached(layer, dt):
    """Low-precision copies of a layer's weight / bias, refreshed when the parameters change."""
    key = (dt, layer.weight._version, layer.weight.data_ptr(), None if layer.bias is None else layer.bias._version)
    c = getattr(layer, "_dvis_cast", None)
    if c is None or c[0] != key:
        c = (key, layer.weight.detach().to(dt), None if layer.bias is None else layer.bias.detach().to(dt))
        layer._dvis_cast = c
    return c[1], c[2]


def flash_attn_wins(B, H, Lq):
    """Where dvis_flash_attn measured faster than cuDNN SDPA on a B200 (profiles/r2_temporal_kernels.md): short query
    sequences (<= 16 rows: attention over time) and small batches of 200-query problems; 16 frames x 8 heads x 200 queries
    run 2x faster on the library kernel."""
    return Lq <= 16 or B * H * ((Lq + 15) // 16) <= 296


class MultiheadAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, dropout=0.0):
        super().__init__()
        assert embed_dim % num_heads == 0 and dropout == 0.0
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.out_proj.bias, 0.0)

    def _weights(self, dt):
        key = (dt, self.in_proj_weight._version, self.in_proj_bias._version, self.in_proj_weight.data_ptr())
        c = getattr(self, "_dvis_cast", None)
        if c is None or c[0] != key:
            c = (key, self.in_proj_weight.detach().to(dt), self.in_proj_bias.detach().to(dt))
            self._dvis_cast = c
        return c[1], c[2]

    def forward(self, query, key, value, attn_mask=None):
        """query (Lq,B,E), key/value (Lk,B,E); attn_mask bool (B*H,Lq,Lk) or (Lq,Lk), True = may NOT attend.
        Returns (Lq,B,E) in the GEMM dtype."""
        Lq, B, E = query.shape
        Lk = key.shape[0]
        H, dh = self.num_heads, self.head_dim
        fast = _fast_path(query)
        dt = gemm_dtype() if fast else query.dtype
        if fast:
            w, b = self._weights(dt)
        else:
            w, b = self.in_proj_weight, self.in_proj_bias
        same_qk, same_kv = query is key, key is value
        query = query.to(dt)
        key = query if same_qk else key.to(dt)
        value = key if same_kv else value.to(dt)
        if same_qk and same_kv:
            q, k, v = F.linear(query, w, b).chunk(3, dim=-1)
        elif same_kv:
            q = F.linear(query, w[:E], b[:E])
            k, v = F.linear(key, w[E:], b[E:]).chunk(2, dim=-1)
        elif same_qk:
            q, k = F.linear(query, w[:2 * E], b[:2 * E]).chunk(2, dim=-1)
            v = F.linear(value, w[2 * E:], b[2 * E:])
        else:
            q = F.linear(query, w[:E], b[:E])
            k = F.linear(key, w[E:2 * E], b[E:2 * E])
            v = F.linear(value, w[2 * E:], b[2 * E:])
        # (L, B, H*dh) -> (B, H, L, dh)
        q = q.reshape(Lq, B, H, dh).permute(1, 2, 0, 3)
        k = k.reshape(Lk, B, H, dh).permute(1, 2, 0, 3)
        v = v.reshape(Lk, B, H, dh).permute(1, 2, 0, 3)
        mask = None
        if attn_mask is not None:
            mask = ~attn_mask.reshape(B, H, Lq, Lk) if attn_mask.dim() == 3 else ~attn_mask
        if fast and attn_mask is None and dt == torch.bfloat16 and dh in (32, 64) and flash_attn_wins(B, H, Lq):
            # latency-bound shapes (one 200-query problem, or the refiner's 200 x 8 attentions over 16 frames): the
            # hand-written core (csrc/flash_attn.cu) beats the library kernel; big batches stay on cuDNN's flash kernel
            o = ops.flash_attn(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3), 1.0 / math.sqrt(dh))   # (B, Lq, E)
            return linear(self.out_proj, o.transpose(0, 1))
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, scale=1.0 / math.sqrt(dh))
        o = o.permute(2, 0, 1, 3).reshape(Lq, B, E)
        return linear(self.out_proj, o)


def _with_pos(t, pos):
    return t if pos is None else t + pos


def _xavier_reset(module):
    for p in module.parameters():
        if p.dim() > 1:
            nn.init.xavier_uniform_(p)


class SelfAttentionLayer(nn.Module):
    def __init__(self, d_model, nhead, dropout=0.0, activation="relu", normalize_before=False):
        super().__init__()
        self.self_attn = MultiheadAttention(d_model, nhead, dropout=dropout)
        self.norm = nn.LayerNorm(d_model)
        self.normalize_before = normalize_before
        _xavier_reset(self)

    def forward(self, tgt, tgt_mask=None, tgt_key_padding_mask=None, query_pos=None):
        assert tgt_key_padding_mask is None
        if self.normalize_before:
            t2 = self.norm(tgt)
            qk = _with_pos(t2, query_pos)
            return tgt + self.self_attn(qk, qk, t2, attn_mask=tgt_mask)
        qk = _with_pos(tgt, query_pos)
        if query_pos is None:
            t2 = self.self_attn(tgt, tgt, tgt, attn_mask=tgt_mask)
        else:
            t2 = self.self_attn(qk, qk, tgt, attn_mask=tgt_mask)
        return add_norm(self.norm, t2, tgt)


class CrossAttentionLayer(nn.Module):
    def __init__(self, d_model, nhead, dropout=0.0, activation="relu", normalize_before=False):
        super().__init__()
        self.multihead_attn = MultiheadAttention(d_model, nhead, dropout=dropout)
        self.norm = nn.LayerNorm(d_model)
        self.normalize_before = normalize_before
        _xavier_reset(self)

    def forward(self, tgt, memory, memory_mask=None, memory_key_padding_mask=None, pos=None, query_pos=None):
        assert memory_key_padding_mask is None
        if self.normalize_before:
            t2 = self.norm(tgt)
            return tgt + self.multihead_attn(_with_pos(t2, query_pos), _with_pos(memory, pos), memory, attn_mask=memory_mask)
        key = memory if pos is None else memory + pos
        t2 = self.multihead_attn(_with_pos(tgt, query_pos), key, memory, attn_mask=memory_mask)
        return add_norm(self.norm, t2, tgt)


class ReferringCrossAttentionLayer(nn.Module):
    """P/dvis_Plus/tracker.py:8-92: residual comes from `indentify`, query from `tgt`, key != value allowed."""

    def __init__(self, d_model, nhead, dropout=0.0, activation="relu", normalize_before=False):
        super().__init__()
        self.multihead_attn = MultiheadAttention(d_model, nhead, dropout=dropout)
        self.norm = nn.LayerNorm(d_model)
        self.normalize_before = normalize_before
        _xavier_reset(self)

    def forward(self, indentify, tgt, key, memory, memory_mask=None, memory_key_padding_mask=None, pos=None, query_pos=None):
        assert memory_key_padding_mask is None
        if self.normalize_before:
            t2 = self.norm(tgt)
            return indentify + self.multihead_attn(_with_pos(t2, query_pos), _with_pos(key, pos), memory, attn_mask=memory_mask)
        k = key if pos is None else key + pos
        if pos is None and key is memory:
            k = memory
        t2 = self.multihead_attn(_with_pos(tgt, query_pos), k, memory, attn_mask=memory_mask)
        return add_norm(self.norm, t2, indentify)


class FFNLayer(nn.Module):
    def __init__(self, d_model, dim_feedforward=2048, dropout=0.0, activation="relu", normalize_before=False):
        super().__init__()
        assert activation == "relu", "DVIS configs use relu"
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm = nn.LayerNorm(d_model)
        self.normalize_before = normalize_before
        _xavier_reset(self)

    def forward(self, tgt):
        if self.normalize_before:
            return tgt + linear(self.linear2, linear(self.linear1, self.norm(tgt), relu=True))
        t2 = linear(self.linear2, linear(self.linear1, tgt, relu=True))
        return add_norm(self.norm, t2, tgt)


class MLP(nn.Module):
    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))

    def forward(self, x):
        for i, layer in enumerate(self.layers):
            x = linear(layer, x, relu=i < self.num_layers - 1)
        return x


def sine_position_embedding(H, W, num_pos_feats, device, temperature=10000.0):
    """PositionEmbeddingSine(normalize=True) for an all-valid (H, W) map
    (P/mask2former/modeling/transformer_decoder/position_encoding.py:29-52).  -> (2*num_pos_feats, H, W) fp32."""
    scale, eps = 2 * math.pi, 1e-6
    y = torch.arange(1, H + 1, dtype=torch.float32, device=device) / (H + eps) * scale
    x = torch.arange(1, W + 1, dtype=torch.float32, device=device) / (W + eps) * scale
    i = torch.arange(num_pos_feats, dtype=torch.float32, device=device)
    dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode="floor") / num_pos_feats)

    def enc(v):
        a = v[:, None] / dim_t
        return torch.stack((a[:, 0::2].sin(), a[:, 1::2].cos()), dim=2).flatten(1)

    py = enc(y)[:, None, :].expand(H, W, num_pos_feats)
    px = enc(x)[None, :, :].expand(H, W, num_pos_feats)
    return torch.cat((py, px), dim=2).permute(2, 0, 1).contiguous()
