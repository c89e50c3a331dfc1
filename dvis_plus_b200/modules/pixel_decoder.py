"""Drop-in for P/mask2former/modeling/pixel_decoder/msdeformattn.py (MSDeformAttnPixelDecoder and its encoder).

Same class names, constructor keywords, module / parameter names (so `state_dict`s are interchangeable) and the
same `forward_features(features) -> (mask_features, out[0], multi_scale_features[:3])` contract (py:314-358).

B200-first differences in HOW it runs (results match the reference within the stated tolerances):
  * tokens stay channel-last end to end: (N, S, C) encoder tokens, channels_last conv maps, and `mask_features`
    is emitted in torch.channels_last memory format (logical shape still (N, C, H, W)) -- the K-major layout the
    tcgen05 mask GEMM streams with TMA;
  * per-shape constants (sine position embeddings + level embedding, reference points, level tables, the L1
    locality schedule) are cached instead of being rebuilt every forward (py:62-84,141-153);
  * each encoder layer is: one GEMM for value, ONE GEMM for [offsets | logits], the fused MSDA kernel (softmax +
    location arithmetic + gather), output projection, and fused residual + LayerNorm kernels that also emit the
    low-precision operands of the next GEMMs.
"""
import math
from typing import Callable, Dict, List, Optional, Union

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.init import normal_

from .. import ops
from .blocks import _fast_path, sine_position_embedding
from .ms_deform_attn import MSDeformAttn
from .precision import gemm_dtype

try:  # optional: when Detectron2 is installed the class registers itself like the reference (py:164)
    from detectron2.config import configurable
    from detectron2.modeling import SEM_SEG_HEADS_REGISTRY
    _HAVE_D2 = True
except Exception:  # pragma: no cover - Detectron2 is not part of this image
    _HAVE_D2 = False

    def configurable(f=None, **kw):
        return f if f is not None else (lambda g: g)


class ShapeSpec:
    """Minimal stand-in for detectron2.layers.ShapeSpec (channels, stride)."""

    def __init__(self, channels=None, height=None, width=None, stride=None):
        self.channels, self.height, self.width, self.stride = channels, height, width, stride


class ConvNorm(nn.Conv2d):
    """detectron2.layers.Conv2d equivalent: conv -> optional norm (child module `norm`) -> optional activation."""

    def __init__(self, *args, norm=None, activation=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation

    def forward(self, x):
        dt = gemm_dtype() if _fast_path(x) else x.dtype
        w = self.weight if self.weight.dtype == dt else self.weight.to(dt)
        b = None if self.bias is None else (self.bias if self.bias.dtype == dt else self.bias.to(dt))
        x = F.conv2d(x.to(dt), w, b, self.stride, self.padding)
        if self.norm is not None:
            x = F.group_norm(x.float(), self.norm.num_groups, self.norm.weight, self.norm.bias, self.norm.eps)
        if self.activation is not None:
            x = self.activation(x)
        return x


def _c2_xavier_fill(m):
    nn.init.kaiming_uniform_(m.weight, a=1)
    if m.bias is not None:
        nn.init.constant_(m.bias, 0)


def _get_clones(module, n):
    import copy
    return nn.ModuleList([copy.deepcopy(module) for _ in range(n)])


class MSDeformAttnTransformerEncoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        assert activation == "relu"
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, padding_mask=None):
        """Reference-style forward (py:122-131); used under autograd."""
        src2 = self.self_attn(src if pos is None else src + pos, reference_points, src, spatial_shapes, level_start_index, padding_mask)
        src = self.norm1(src + self.dropout1(src2))
        src2 = self.linear2(self.dropout2(F.relu(self.linear1(src))))
        return self.norm2(src + self.dropout3(src2))

    def _lp(self, dt):
        key = (dt, self.linear1.weight._version, self.linear2.weight._version, self.linear1.weight.data_ptr())
        c = getattr(self, "_dvis_cast", None)
        if c is None or c[0] != key:
            c = (key, self.linear1.weight.detach().to(dt), self.linear1.bias.detach().to(dt),
                 self.linear2.weight.detach().to(dt), self.linear2.bias.detach().to(dt))
            self._dvis_cast = c
        return c[1:]

    def forward_fused(self, src32, src_lp, q_lp, pos, ctx, next_needs_q=True):
        """Inference path.  src32: fp32 residual stream (N,S,C); src_lp / q_lp: GEMM-dtype copies of src and src+pos.
        Returns (src32, src_lp, q_lp) for the next layer."""
        dt = src_lp.dtype
        w1, b1, w2, b2 = self._lp(dt)
        attn = self.self_attn
        if attn.fuse_output_norm and attn.tc_path_ok(dt, src_lp.shape[1]):
            # output_proj + residual + norm1 in ONE tcgen05 kernel (csrc/linear_tc.cu): the projection never reaches HBM
            core = attn.forward_core(q_lp, ctx["ref"], src_lp, ctx["shapes"], ctx["shapes_dev"], ctx["lsi_dev"])
            ow = attn._fused_weights(dt)[4]
            src32, src_lp, _ = ops.linear_tc_add_layernorm(core, ow, attn.output_proj.bias.detach().float(), src32, self.norm1.weight,
                                                           self.norm1.bias, self.norm1.eps)
        else:
            a = attn.forward_fused(q_lp, ctx["ref"], src_lp, ctx["shapes"], ctx["shapes_dev"], ctx["lsi_dev"])
            src32, src_lp, _ = ops.add_layernorm(a, src32, self.norm1.weight, self.norm1.bias, self.norm1.eps, lp_dtype=dt)
        h = torch._addmm_activation(b1, src_lp.view(-1, src_lp.shape[-1]), w1.t())      # ReLU in the GEMM epilogue
        f = F.linear(h, w2, b2).view(src_lp.shape)
        return ops.add_layernorm(f, src32, self.norm2.weight, self.norm2.bias, self.norm2.eps, lp_dtype=dt,
                                 pos=pos if next_needs_q else None)


class MSDeformAttnTransformerEncoder(nn.Module):
    def __init__(self, encoder_layer, num_layers):
        super().__init__()
        self.layers = _get_clones(encoder_layer, num_layers)
        self.num_layers = num_layers

    @staticmethod
    def get_reference_points(spatial_shapes, valid_ratios, device):
        """py:141-153."""
        pts = []
        for lvl, (H_, W_) in enumerate(spatial_shapes):
            H_, W_ = int(H_), int(W_)
            ref_y, ref_x = torch.meshgrid(torch.linspace(0.5, H_ - 0.5, H_, dtype=torch.float32, device=device),
                                          torch.linspace(0.5, W_ - 0.5, W_, dtype=torch.float32, device=device), indexing="ij")
            ref_y = ref_y.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H_)
            ref_x = ref_x.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W_)
            pts.append(torch.stack((ref_x, ref_y), -1))
        reference_points = torch.cat(pts, 1)
        return reference_points[:, :, None] * valid_ratios[:, None]

    def forward(self, src, spatial_shapes, level_start_index, valid_ratios, pos=None, padding_mask=None):
        output = src
        reference_points = self.get_reference_points(spatial_shapes, valid_ratios, device=src.device)
        for layer in self.layers:
            output = layer(output, pos, reference_points, spatial_shapes, level_start_index, padding_mask)
        return output


class MSDeformAttnTransformerEncoderOnly(nn.Module):
    def __init__(self, d_model=256, nhead=8, num_encoder_layers=6, dim_feedforward=1024, dropout=0.1, activation="relu",
                 num_feature_levels=4, enc_n_points=4):
        super().__init__()
        self.d_model, self.nhead = d_model, nhead
        layer = MSDeformAttnTransformerEncoderLayer(d_model, dim_feedforward, dropout, activation, num_feature_levels,
                                                    nhead, enc_n_points)
        self.encoder = MSDeformAttnTransformerEncoder(layer, num_encoder_layers)
        self.level_embed = nn.Parameter(torch.Tensor(num_feature_levels, d_model))
        self._reset_parameters()
        self._ctx_cache = {}

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()
        normal_(self.level_embed)

    def _context(self, shapes, device):
        """Per-shape constants: level-embedded sine position embedding (S, C), reference points, level tables."""
        key = (shapes, str(device), self.level_embed._version, self.level_embed.data_ptr())
        ctx = self._ctx_cache.get(key)
        if ctx is None:
            C = self.d_model
            pos = torch.cat([sine_position_embedding(h, w, C // 2, device).flatten(1).t() + self.level_embed[l].detach().view(1, -1)
                             for l, (h, w) in enumerate(shapes)], 0).contiguous()
            shapes_dev = torch.as_tensor(shapes, dtype=torch.long, device=device)
            lsi_dev = torch.cat((shapes_dev.new_zeros((1,)), shapes_dev.prod(1).cumsum(0)[:-1]))
            ones = torch.ones(1, len(shapes), 2, device=device)
            ref = MSDeformAttnTransformerEncoder.get_reference_points(shapes, ones, device)   # (1, S, L, 2)
            ctx = dict(pos=pos, shapes=shapes, shapes_dev=shapes_dev, lsi_dev=lsi_dev, ref1=ref.contiguous())
            if len(self._ctx_cache) > 8:
                self._ctx_cache.clear()
            self._ctx_cache[key] = ctx
        return ctx

    def forward(self, srcs, pos_embeds=None):
        """srcs: list of (N, C, H, W) maps (low -> high resolution).  Returns (memory (N,S,C) fp32, spatial_shapes,
        level_start_index) like the reference (py:61-89).  `pos_embeds` is ignored on the inference path because the
        embedding depends on the shapes only (all-valid masks, py:62) and is cached."""
        shapes = tuple((int(s.shape[2]), int(s.shape[3])) for s in srcs)
        N = srcs[0].shape[0]
        ctx = self._context(shapes, srcs[0].device)
        src = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)          # (N, S, C)
        if not _fast_path(src):
            pos = ctx["pos"][None].expand(N, -1, -1)
            valid = torch.ones(N, len(shapes), 2, device=src.device)
            memory = self.encoder(src.float(), ctx["shapes_dev"], ctx["lsi_dev"], valid, pos, None)
            return memory, ctx["shapes_dev"], ctx["lsi_dev"]
        dt = gemm_dtype()
        src32 = src.float().contiguous()
        src_lp = src32.to(dt)
        q_lp = (src32 + ctx["pos"][None]).to(dt)
        return self.run_layers(src32, src_lp, q_lp, ctx), ctx["shapes_dev"], ctx["lsi_dev"]

    def run_layers(self, src32, src_lp, q_lp, ctx):
        """The encoder stack on prepared inputs: fp32 stream, GEMM-dtype copy, GEMM-dtype (src + pos).  -> (N,S,C) fp32."""
        N = src32.shape[0]
        run = dict(ctx)
        key = ("ref", N)
        if key not in ctx:
            ctx[key] = ctx["ref1"].expand(N, -1, -1, -1).contiguous()
        run["ref"] = ctx[key]
        n_layers = len(self.encoder.layers)
        for i, layer in enumerate(self.encoder.layers):
            src32, src_lp, q_lp = layer.forward_fused(src32, src_lp, q_lp, ctx["pos"], run, next_needs_q=i + 1 < n_layers)
        return src32


class MSDeformAttnPixelDecoder(nn.Module):
    @configurable
    def __init__(self, input_shape: Dict[str, ShapeSpec], *, transformer_dropout: float, transformer_nheads: int,
                 transformer_dim_feedforward: int, transformer_enc_layers: int, conv_dim: int, mask_dim: int,
                 norm: Optional[Union[str, Callable]] = None, transformer_in_features: List[str], common_stride: int):
        super().__init__()
        transformer_input_shape = {k: v for k, v in input_shape.items() if k in transformer_in_features}
        input_shape = sorted(input_shape.items(), key=lambda x: x[1].stride)
        self.in_features = [k for k, v in input_shape]
        self.feature_strides = [v.stride for k, v in input_shape]
        self.feature_channels = [v.channels for k, v in input_shape]
        transformer_input_shape = sorted(transformer_input_shape.items(), key=lambda x: x[1].stride)
        self.transformer_in_features = [k for k, v in transformer_input_shape]
        transformer_in_channels = [v.channels for k, v in transformer_input_shape]
        self.transformer_feature_strides = [v.stride for k, v in transformer_input_shape]
        self.transformer_num_feature_levels = len(self.transformer_in_features)

        chans = transformer_in_channels[::-1] if self.transformer_num_feature_levels > 1 else [transformer_in_channels[-1]]
        self.input_proj = nn.ModuleList([nn.Sequential(nn.Conv2d(c, conv_dim, kernel_size=1), nn.GroupNorm(32, conv_dim))
                                         for c in chans])
        for proj in self.input_proj:
            nn.init.xavier_uniform_(proj[0].weight, gain=1)
            nn.init.constant_(proj[0].bias, 0)

        self.transformer = MSDeformAttnTransformerEncoderOnly(
            d_model=conv_dim, dropout=transformer_dropout, nhead=transformer_nheads,
            dim_feedforward=transformer_dim_feedforward, num_encoder_layers=transformer_enc_layers,
            num_feature_levels=self.transformer_num_feature_levels)
        self.conv_dim = conv_dim
        self.mask_dim = mask_dim
        self.mask_features = ConvNorm(conv_dim, mask_dim, kernel_size=1, stride=1, padding=0)
        _c2_xavier_fill(self.mask_features)
        self.maskformer_num_feature_levels = 3
        self.common_stride = common_stride
        stride = min(self.transformer_feature_strides)
        self.num_fpn_levels = int(np.log2(stride) - np.log2(self.common_stride))

        lateral_convs, output_convs = [], []
        use_bias = norm == ""
        for idx, in_channels in enumerate(self.feature_channels[:self.num_fpn_levels]):
            assert norm in ("GN", "", None), norm
            mk = (lambda: nn.GroupNorm(32, conv_dim)) if norm == "GN" else (lambda: None)
            lateral_conv = ConvNorm(in_channels, conv_dim, kernel_size=1, bias=use_bias, norm=mk())
            output_conv = ConvNorm(conv_dim, conv_dim, kernel_size=3, stride=1, padding=1, bias=use_bias, norm=mk(),
                                   activation=F.relu)
            _c2_xavier_fill(lateral_conv)
            _c2_xavier_fill(output_conv)
            self.add_module("adapter_{}".format(idx + 1), lateral_conv)
            self.add_module("layer_{}".format(idx + 1), output_conv)
            lateral_convs.append(lateral_conv)
            output_convs.append(output_conv)
        self.lateral_convs = lateral_convs[::-1]
        self.output_convs = output_convs[::-1]

    @classmethod
    def from_config(cls, cfg, input_shape):
        ret = {}
        ret["input_shape"] = {k: v for k, v in input_shape.items() if k in cfg.MODEL.SEM_SEG_HEAD.IN_FEATURES}
        ret["conv_dim"] = cfg.MODEL.SEM_SEG_HEAD.CONVS_DIM
        ret["mask_dim"] = cfg.MODEL.SEM_SEG_HEAD.MASK_DIM
        ret["norm"] = cfg.MODEL.SEM_SEG_HEAD.NORM
        ret["transformer_dropout"] = cfg.MODEL.MASK_FORMER.DROPOUT
        ret["transformer_nheads"] = cfg.MODEL.MASK_FORMER.NHEADS
        ret["transformer_dim_feedforward"] = 1024  # the reference hard-codes 1024 for the deformable encoder (py:307)
        ret["transformer_enc_layers"] = cfg.MODEL.SEM_SEG_HEAD.TRANSFORMER_ENC_LAYERS
        ret["transformer_in_features"] = cfg.MODEL.SEM_SEG_HEAD.DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES
        ret["common_stride"] = cfg.MODEL.SEM_SEG_HEAD.COMMON_STRIDE
        return ret

    # -- inference path: channels-last tokens end to end ------------------------------------------------------
    def _lp_weights(self, dt):
        params = list(self.input_proj.parameters()) + [p for c in self.lateral_convs + self.output_convs for p in c.parameters()] \
            + list(self.mask_features.parameters())
        key = (dt, tuple(p._version for p in params), params[0].data_ptr())
        c = getattr(self, "_dvis_lp", None)
        if c is None or c[0] != key:
            d = lambda t: None if t is None else t.detach().to(dt).contiguous()
            c = (key, dict(
                inproj=[(d(seq[0].weight.flatten(1)), d(seq[0].bias)) for seq in self.input_proj],
                lateral=[(d(cv.weight.flatten(1)), d(cv.bias)) for cv in self.lateral_convs],
                output=[(cv.weight.detach().to(dt).contiguous(memory_format=torch.channels_last), d(cv.bias)) for cv in self.output_convs],
                mask=(d(self.mask_features.weight.flatten(1)), d(self.mask_features.bias))))
            self._dvis_lp = c
        return c[1]

    @staticmethod
    def _tokens(x, dt):
        """(N, C, H, W) map -> (N, H*W, C) channels-last token view in the GEMM dtype (no copy when the producer already
        emits channels_last, as Swin / ViT backbones natively do)."""
        if x.dtype != dt or not x.is_contiguous(memory_format=torch.channels_last):
            x = x.to(dtype=dt, memory_format=torch.channels_last)
        N, C, H, W = x.shape
        return x.permute(0, 2, 3, 1).reshape(N, H * W, C)

    def _fused_ok(self):
        tf = self.transformer
        attn = tf.encoder.layers[0].self_attn
        return (self.conv_dim % 128 == 0 and attn.n_points == 4 and (attn.d_model // attn.n_heads) in (16, 32, 64)
                and all(isinstance(seq[1], nn.GroupNorm) for seq in self.input_proj)
                and all(isinstance(cv.norm, nn.GroupNorm) for cv in self.lateral_convs + self.output_convs)
                and self.transformer_num_feature_levels >= self.maskformer_num_feature_levels)

    def _forward_features_fused(self, features):
        dt = gemm_dtype()
        tf, C = self.transformer, self.conv_dim
        lp = self._lp_weights(dt)
        maps = [features[f] for f in self.transformer_in_features[::-1]]            # low -> high resolution
        N, dev = maps[0].shape[0], maps[0].device
        shapes = tuple((int(m.shape[2]), int(m.shape[3])) for m in maps)
        ctx = tf._context(shapes, dev)
        S = sum(h * w for h, w in shapes)
        src32 = torch.empty(N, S, C, dtype=torch.float32, device=dev)
        src_lp = torch.empty(N, S, C, dtype=dt, device=dev)
        q_lp = torch.empty(N, S, C, dtype=dt, device=dev)
        off, offsets = 0, []
        for idx, (m, (h, w)) in enumerate(zip(maps, shapes)):
            gn = self.input_proj[idx][1]
            y = F.linear(self._tokens(m, dt), *lp["inproj"][idx])                    # 1x1 conv == per-pixel linear
            sl = slice(off, off + h * w)
            ops.groupnorm_nhwc(y, gn.num_groups, gn.weight, gn.bias, gn.eps, pos=ctx["pos"][sl],
                               out_f32=src32[:, sl], out_lp=src_lp[:, sl], out_lp_pos=q_lp[:, sl])
            offsets.append(off)
            off += h * w
        memory = tf.run_layers(src32, src_lp, q_lp, ctx)                              # (N, S, C) fp32
        out = [memory[:, o:o + h * w].view(N, h, w, C).permute(0, 3, 1, 2) for o, (h, w) in zip(offsets, shapes)]
        cur, cur_hw = memory[:, offsets[-1]:], shapes[-1]                             # highest-resolution encoder level
        last_lp = None
        for idx, f in enumerate(self.in_features[:self.num_fpn_levels][::-1]):
            x = features[f]
            H, W = int(x.shape[2]), int(x.shape[3])
            lat, outc = self.lateral_convs[idx], self.output_convs[idx]
            y = F.linear(self._tokens(x, dt), *lp["lateral"][idx])
            z = torch.empty(N, H * W, C, dtype=dt, device=dev)
            ops.groupnorm_nhwc(y, lat.norm.num_groups, lat.norm.weight, lat.norm.bias, lat.norm.eps,
                               up=cur, up_hw=cur_hw, hw=(H, W), out_lp=z)               # GN(lateral) + upsample(top-down)
            w3, b3 = lp["output"][idx]
            c = F.conv2d(z.view(N, H, W, C).permute(0, 3, 1, 2), w3, b3, padding=1)
            c = c.permute(0, 2, 3, 1).reshape(N, H * W, C)
            last_lp = torch.empty(N, H * W, C, dtype=dt, device=dev)
            more = idx + 1 < self.num_fpn_levels
            nxt = torch.empty(N, H * W, C, dtype=torch.float32, device=dev) if more else None
            ops.groupnorm_nhwc(c, outc.norm.num_groups, outc.norm.weight, outc.norm.bias, outc.norm.eps, relu=True,
                               out_lp=last_lp, out_f32=nxt)
            cur, cur_hw = nxt, (H, W)
        if last_lp is None:                                                           # no FPN level: mask features from the encoder
            H, W = shapes[-1]
            last_lp = memory[:, offsets[-1]:].to(dt)
        mf = F.linear(last_lp, *lp["mask"])                                           # (N, H*W, mask_dim)
        mask_features = mf.view(N, H, W, self.mask_dim).permute(0, 3, 1, 2)           # NCHW shape, channels_last memory
        return mask_features, out[0], out[:self.maskformer_num_feature_levels]

    def forward_features(self, features):
        """-> (mask_features (N, mask_dim, H/4, W/4), out[0], multi_scale_features[:3])  (py:314-358).
        On the inference path the returned maps are channels_last; mask_features is in the GEMM dtype."""
        if _fast_path(next(iter(features.values()))) and self._fused_ok():
            with torch.autocast("cuda", enabled=False):
                return self._forward_features_fused(features)
        with torch.autocast("cuda", enabled=False):
            fast = _fast_path(next(iter(features.values())))
            dt = gemm_dtype() if fast else torch.float32
            srcs = []
            for idx, f in enumerate(self.transformer_in_features[::-1]):
                x = features[f]
                conv, gn = self.input_proj[idx][0], self.input_proj[idx][1]
                if fast:
                    x = x.to(dtype=dt, memory_format=torch.channels_last)
                    y = F.conv2d(x, conv.weight.to(dt), conv.bias.to(dt))
                    srcs.append(F.group_norm(y.float(), gn.num_groups, gn.weight, gn.bias, gn.eps))
                else:
                    srcs.append(gn(conv(x.float())))
            y, spatial_shapes, level_start_index = self.transformer(srcs)
            bs = y.shape[0]
            out = []
            start = 0
            for (h, w) in [(int(s.shape[2]), int(s.shape[3])) for s in srcs]:
                # (N, h*w, C) token block viewed as an (N, C, h, w) map: channels_last strides, no copy
                out.append(y[:, start:start + h * w].reshape(bs, h, w, -1).permute(0, 3, 1, 2))
                start += h * w
            for idx, f in enumerate(self.in_features[:self.num_fpn_levels][::-1]):
                x = features[f]
                x = x.to(dtype=dt, memory_format=torch.channels_last) if fast else x.float()
                cur_fpn = self.lateral_convs[idx](x)
                y = cur_fpn + F.interpolate(out[-1].float(), size=cur_fpn.shape[-2:], mode="bilinear", align_corners=False)
                out.append(self.output_convs[idx](y))
            multi_scale_features = out[:self.maskformer_num_feature_levels]
            mask_features = self.mask_features(out[-1])
            if fast:
                mask_features = mask_features.contiguous(memory_format=torch.channels_last)
            return mask_features, out[0], multi_scale_features


if _HAVE_D2:  # pragma: no cover
    MSDeformAttnPixelDecoder = SEM_SEG_HEADS_REGISTRY.register()(MSDeformAttnPixelDecoder)
