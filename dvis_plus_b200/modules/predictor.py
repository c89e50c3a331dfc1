"""Drop-in for VideoMultiScaleMaskedTransformerDecoder_dvisPlus
(P/dvis_Plus/video_mask2former_transformer_decoder.py:174-374; base class
P/mask2former_video/modeling/transformer_decoder/video_mask2former_transformer_decoder.py:208-474).

Same constructor keywords, parameter names and output dictionary.  The mask head (`forward_prediction_heads`,
py:358-374) runs on the tcgen05 mask GEMM.  At inference the full-resolution logits of the intermediate layers
are never materialised (the meta-architecture deletes them, P/dvis_Plus/meta_architecture.py:787-788,1465-1467):
bilinear resizing is linear, so  interpolate(E @ F) == E @ interpolate(F)  and  sigmoid(x) < 0.5 <=> x < 0;
the attention mask of each layer is therefore  (E @ F_l) < 0  with F_l = mask_features resized once per level.
Set `materialize_aux_masks=True` (or run under autograd) to get the reference's every-layer `aux_outputs` masks.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from .blocks import (MLP, CrossAttentionLayer, FFNLayer, SelfAttentionLayer, _fast_path, add_norm, flash_attn_wins, linear,
                     sine_position_embedding)
from .precision import gemm_dtype
from .pixel_decoder import ConvNorm, _c2_xavier_fill, configurable

try:  # optional registration, mirrors py:174
    from mask2former.modeling.transformer_decoder.maskformer_transformer_decoder import TRANSFORMER_DECODER_REGISTRY
    _HAVE_REG = True
except Exception:  # pragma: no cover
    _HAVE_REG = False


class VideoMultiScaleMaskedTransformerDecoder_dvisPlus(nn.Module):
    _version = 2

    @configurable
    def __init__(self, in_channels, mask_classification=True, *, num_classes: int, hidden_dim: int, num_queries: int,
                 nheads: int, dim_feedforward: int, dec_layers: int, pre_norm: bool, mask_dim: int,
                 enforce_input_project: bool, num_frames: int, num_reid_head_layers, reid_hidden_dim):
        super().__init__()
        assert mask_classification, "Only support mask classification model"
        self.mask_classification = mask_classification
        self.num_frames = num_frames
        self.num_heads = nheads
        self.num_layers = dec_layers
        self.hidden_dim = hidden_dim
        self.pre_norm = pre_norm
        self.transformer_self_attention_layers = nn.ModuleList()
        self.transformer_cross_attention_layers = nn.ModuleList()
        self.transformer_ffn_layers = nn.ModuleList()
        for _ in range(self.num_layers):
            self.transformer_self_attention_layers.append(SelfAttentionLayer(hidden_dim, nheads, 0.0, normalize_before=pre_norm))
            self.transformer_cross_attention_layers.append(CrossAttentionLayer(hidden_dim, nheads, 0.0, normalize_before=pre_norm))
            self.transformer_ffn_layers.append(FFNLayer(hidden_dim, dim_feedforward, 0.0, normalize_before=pre_norm))
        self.decoder_norm = nn.LayerNorm(hidden_dim)
        self.num_queries = num_queries
        self.query_feat = nn.Embedding(num_queries, hidden_dim)
        self.query_embed = nn.Embedding(num_queries, hidden_dim)
        self.num_feature_levels = 3
        self.level_embed = nn.Embedding(self.num_feature_levels, hidden_dim)
        self.input_proj = nn.ModuleList()
        for _ in range(self.num_feature_levels):
            if in_channels != hidden_dim or enforce_input_project:
                self.input_proj.append(ConvNorm(in_channels, hidden_dim, kernel_size=1))
                _c2_xavier_fill(self.input_proj[-1])
            else:
                self.input_proj.append(nn.Sequential())
        self.class_embed = nn.Linear(hidden_dim, num_classes + 1)
        self.mask_embed = MLP(hidden_dim, hidden_dim, mask_dim, 3)
        if num_reid_head_layers > 0:
            self.reid_embed = MLP(hidden_dim, reid_hidden_dim, hidden_dim, num_reid_head_layers)
            for layer in self.reid_embed.layers:
                _c2_xavier_fill(layer)
        else:
            self.reid_embed = nn.Identity()
        self.materialize_aux_masks = False
        # bf16 mode: masked cross-attention and query self-attention on csrc/flash_attn.cu; the attention mask is one BIT per
        # (query, pixel) written by the tcgen05 mask GEMM's epilogue (ops.mask_attn_bits) instead of a dense additive bias
        self.use_fused_attention = True
        self._pos_cache = {}

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        # key migration of the reference base class (…/video_mask2former_transformer_decoder.py:213-234)
        version = local_metadata.get("version", None)
        if version is None or version < 2:
            for k in list(state_dict.keys()):
                if k.startswith(prefix) and "static_query" in k:
                    state_dict[k.replace("static_query", "query_feat")] = state_dict.pop(k)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    @classmethod
    def from_config(cls, cfg, in_channels, mask_classification):
        ret = {"in_channels": in_channels, "mask_classification": mask_classification}
        ret["num_classes"] = cfg.MODEL.SEM_SEG_HEAD.NUM_CLASSES
        ret["hidden_dim"] = cfg.MODEL.MASK_FORMER.HIDDEN_DIM
        ret["num_queries"] = cfg.MODEL.MASK_FORMER.NUM_OBJECT_QUERIES
        ret["nheads"] = cfg.MODEL.MASK_FORMER.NHEADS
        ret["dim_feedforward"] = cfg.MODEL.MASK_FORMER.DIM_FEEDFORWARD
        assert cfg.MODEL.MASK_FORMER.DEC_LAYERS >= 1
        ret["dec_layers"] = cfg.MODEL.MASK_FORMER.DEC_LAYERS - 1
        ret["pre_norm"] = cfg.MODEL.MASK_FORMER.PRE_NORM
        ret["enforce_input_project"] = cfg.MODEL.MASK_FORMER.ENFORCE_INPUT_PROJ
        ret["mask_dim"] = cfg.MODEL.SEM_SEG_HEAD.MASK_DIM
        ret["num_frames"] = cfg.INPUT.SAMPLING_FRAME_NUM
        ret["reid_hidden_dim"] = cfg.MODEL.MASK_FORMER.REID_HIDDEN_DIM
        ret["num_reid_head_layers"] = cfg.MODEL.MASK_FORMER.NUM_REID_HEAD_LAYERS
        return ret

    def _pos(self, H, W, device):
        key = (H, W, str(device))
        if key not in self._pos_cache:
            self._pos_cache[key] = sine_position_embedding(H, W, self.hidden_dim // 2, device).flatten(1).t().contiguous()[:, None, :]
        return self._pos_cache[key]          # (H*W, 1, C)

    # ---- mask head ----------------------------------------------------------------------------------
    def forward_prediction_heads(self, output, mask_features, attn_mask_target_size):
        """Reference contract (py:358-374): (Q,B,C) -> class (B,Q,K+1), masks (B,Q,H,W), bool attn_mask (B*h,Q,hw)."""
        decoder_output = self.decoder_norm(output.float()).transpose(0, 1)
        outputs_class = linear(self.class_embed, decoder_output).float()
        mask_embed = self.mask_embed(decoder_output)
        if _fast_path(mask_features):
            outputs_mask = ops.mask_logits(mask_embed, mask_features, torch.float32)
        else:
            outputs_mask = torch.einsum("bqc,bchw->bqhw", mask_embed.float(), mask_features.float())
        attn_mask = F.interpolate(outputs_mask, size=attn_mask_target_size, mode="bilinear", align_corners=False)
        attn_mask = (attn_mask.sigmoid().flatten(2).unsqueeze(1).repeat(1, self.num_heads, 1, 1).flatten(0, 1) < 0.5).bool()
        return outputs_class, outputs_mask, attn_mask.detach()

    def _heads_lowres(self, output, level_feat):
        """Inference mask head without full-resolution logits: class logits + (B, Q, hw) bool attention mask."""
        decoder_output = self.decoder_norm(output.float()).transpose(0, 1)
        mask_embed = self.mask_embed(decoder_output)
        logits = ops.mask_logits(mask_embed, level_feat, torch.float32)       # (B, Q, h, w) on the level's own grid
        return (logits < 0).flatten(2)

    # ---- inference path ------------------------------------------------------------------------------
    def _stacked(self, dt):
        """GEMM-dtype weights grouped the way the fast path consumes them: the keys / values of the 3 layers that attend
        to the same level come out of ONE GEMM each (their input, the level memory, is the same)."""
        params = list(self.parameters())
        key = (dt, tuple(p._version for p in params), params[0].data_ptr())
        c = getattr(self, "_dvis_fast", None)
        if c is None or c[0] != key:
            C, L, nl = self.hidden_dim, self.num_layers, self.num_feature_levels
            d = lambda t: t.detach().to(dt).contiguous()
            ca = [m.multihead_attn for m in self.transformer_cross_attention_layers]
            f = dict(wk=[], bk=[], wv=[], bv=[], layers_of_level=[])
            for l in range(nl):
                ids = list(range(l, L, nl))
                f["layers_of_level"].append(ids)
                f["wk"].append(d(torch.cat([ca[i].in_proj_weight[C:2 * C] for i in ids], 0)) if ids else None)
                f["bk"].append(d(torch.cat([ca[i].in_proj_bias[C:2 * C] for i in ids], 0)) if ids else None)
                f["wv"].append(d(torch.cat([ca[i].in_proj_weight[2 * C:] for i in ids], 0)) if ids else None)
                f["bv"].append(d(torch.cat([ca[i].in_proj_bias[2 * C:] for i in ids], 0)) if ids else None)
            f["wq"] = [d(m.in_proj_weight[:C]) for m in ca]
            f["bq"] = [d(m.in_proj_bias[:C]) for m in ca]
            c = (key, f)
            self._dvis_fast = c
        return c[1]

    def _mask_embed_of(self, output):
        return self.mask_embed(self.decoder_norm(output))

    def _forward_fast(self, x, mask_features):
        """Batch-first (B, L, C) formulation of forward() for inference: same arithmetic, but level memories are cast /
        position-embedded once, K / V of the layers sharing a level are projected together, the attention mask stays a
        (B, 1, Q, hw) additive bias (no per-head copies) and only the last layer's masks are produced at full resolution."""
        dt = gemm_dtype()
        f = self._stacked(dt)
        C, H, L, nl = self.hidden_dim, self.num_heads, self.num_layers, self.num_feature_levels
        dh = C // H
        scale = 1.0 / math.sqrt(dh)
        B = x[0].shape[0]
        sizes, k_all, v_all, level_feats, kv_rows = [], [], [], [], []
        own_attn = self.use_fused_attention and dt == torch.bfloat16 and dh in (32, 64)
        mask_dt = torch.bfloat16 if dt == torch.bfloat16 else torch.float32       # fp32 tier: TF32 operands for the mask head
        mf_lp = mask_features
        if mf_lp.dtype != mask_dt or not mf_lp.is_contiguous(memory_format=torch.channels_last):
            mf_lp = mf_lp.to(dtype=mask_dt, memory_format=torch.channels_last)

        def level_features(h, w):      # mask features resized once to a level's grid: interpolate(E @ F) == E @ interpolate(F)
            if mask_dt == torch.bfloat16:
                return ops.resize_bilinear_nhwc(mf_lp, (h, w))
            return F.interpolate(mf_lp, size=(h, w), mode="bilinear", align_corners=False).contiguous(memory_format=torch.channels_last)
        for l in range(nl):
            xl = self.input_proj[l](x[l])   # empty nn.Sequential is the identity
            h, w = xl.shape[-2:]
            sizes.append((h, w))
            n_of = len(f["layers_of_level"][l])
            cl = xl.stride(1) == 1 and xl.stride(3) == C and xl.stride(2) == w * C and xl.dtype in (torch.float32, torch.bfloat16)
            if n_of and dt == torch.bfloat16 and cl and C % 4 == 0:
                # src + level_embed and (src + level_embed) + pos, both straight to bf16, in one pass over the level
                tok_lp, key_in = ops.level_tokens(xl, self.level_embed.weight[l].detach().float().contiguous(),
                                                  self._pos(h, w, xl.device)[:, 0].contiguous())
                k = F.linear(key_in, f["wk"][l], f["bk"][l]).view(B, h * w, n_of, H, dh)
                v = F.linear(tok_lp, f["wv"][l], f["bv"][l]).view(B, h * w, n_of, H, dh)
                k_all.append(k.permute(2, 0, 3, 1, 4))
                v_all.append(v.permute(2, 0, 3, 1, 4))
                kv_rows.append((k, v))
                level_feats.append(level_features(h, w))
                continue
            tok = xl.permute(0, 2, 3, 1).reshape(B, h * w, C).float() + self.level_embed.weight[l]      # (B, hw, C)
            if n_of:
                key_in = (tok + self._pos(h, w, tok.device)[:, 0][None]).to(dt)
                k = F.linear(key_in, f["wk"][l], f["bk"][l]).view(B, h * w, n_of, H, dh)
                v = F.linear(tok.to(dt), f["wv"][l], f["bv"][l]).view(B, h * w, n_of, H, dh)
                k_all.append(k.permute(2, 0, 3, 1, 4))                                                # (n_of, B, H, hw, dh) views
                v_all.append(v.permute(2, 0, 3, 1, 4))
                kv_rows.append((k, v))                                                                # (B, hw, n_of, H, dh)
            else:
                k_all.append(None)
                v_all.append(None)
                kv_rows.append(None)
            level_feats.append(level_features(h, w))
        query_embed = self.query_embed.weight.detach().float().contiguous()                            # (Q, C)
        Q = query_embed.shape[0]
        dn = self.decoder_norm
        out32 = self.query_feat.weight.detach().float()[None].expand(B, -1, -1).contiguous()           # (B, Q, C) fp32 stream
        out_lp = out32.to(dt)
        out_q = (out32 + query_embed[None]).to(dt)                                                     # with_pos_embed(tgt, query_pos)

        def attn_bias(level):
            normed_lp = ops.add_layernorm(out32, None, dn.weight, dn.bias, dn.eps, want_f32=False, lp_dtype=dt)[1]
            if own_attn:
                return ops.mask_attn_bits(self.mask_embed(normed_lp), level_feats[level])            # (B, Q, bytes): 1 bit / pixel
            return ops.mask_attn_bias(self.mask_embed(normed_lp), level_feats[level], dt)[:, None]   # (B, 1, Q, hw)

        def ln(norm, x, want_lp=True, want_q=True):
            return ops.add_layernorm(x, out32, norm.weight, norm.bias, norm.eps, lp_dtype=dt if (want_lp or want_q) else None,
                                     pos=query_embed if want_q else None)

        bias = attn_bias(0)
        for i in range(L):
            l = i % nl
            j = i // nl
            sa, ff, ca = self.transformer_self_attention_layers[i], self.transformer_ffn_layers[i], self.transformer_cross_attention_layers[i]
            # masked cross-attention to level l
            q = F.linear(out_q, f["wq"][i], f["bq"][i]).view(B, Q, H, dh)
            if own_attn:
                o = ops.flash_attn(q, kv_rows[l][0][:, :, j], kv_rows[l][1][:, :, j], scale, mask_bits=bias)
            else:
                o = F.scaled_dot_product_attention(q.transpose(1, 2), k_all[l][j], v_all[l][j], attn_mask=bias, scale=scale)
                o = o.transpose(1, 2).reshape(B, Q, C)
            o = linear(ca.multihead_attn.out_proj, o)
            out32, out_lp, out_q = ln(ca.norm, o)
            # self-attention over the queries
            m = sa.self_attn
            w, b = m._weights(dt)
            qk = F.linear(out_q, w[:2 * C], b[:2 * C]).view(B, Q, 2, H, dh)
            v = F.linear(out_lp, w[2 * C:], b[2 * C:]).view(B, Q, H, dh)
            if own_attn and flash_attn_wins(B, H, Q):
                o = ops.flash_attn(qk[:, :, 0], qk[:, :, 1], v, scale)
            else:
                o = F.scaled_dot_product_attention(qk[:, :, 0].transpose(1, 2), qk[:, :, 1].transpose(1, 2), v.transpose(1, 2), scale=scale)
                o = o.transpose(1, 2).reshape(B, Q, C)
            o = linear(m.out_proj, o)
            out32, out_lp, _ = ln(sa.norm, o, want_q=False)
            # FFN
            out32, out_lp, out_q = ln(ff.norm, linear(ff.linear2, linear(ff.linear1, out_lp, relu=True)))
            if i + 1 < L:
                bias = attn_bias((i + 1) % nl)
        output = out32
        normed = self.decoder_norm(output)
        cls = linear(self.class_embed, normed).float()                                                 # (B, Q, K+1)
        masks = ops.mask_logits(self.mask_embed(normed), mf_lp, torch.float32, operand_dtype=mask_dt)   # (B, Q, H, W)
        reid = self.reid_embed(normed).float()
        b = B // self.num_frames if self.training else 1
        t = B // b
        to_bctq = lambda z: z.reshape(b, t, Q, z.shape[-1]).permute(0, 3, 1, 2)                        # (b t) q c -> b c t q
        pe, re_, nn_ = to_bctq(normed), to_bctq(reid), to_bctq(output)
        return {
            "pred_logits": cls.reshape(b, t, Q, -1),
            "pred_masks": masks.reshape(b, t, Q, *masks.shape[-2:]).permute(0, 2, 1, 3, 4),
            "aux_outputs": [],
            "pred_embds": torch.cat([pe, re_], dim=1),
            "pred_embds_without_norm": torch.cat([nn_, re_], dim=1),
            "pred_reid_embed": re_,
            "mask_features": mask_features,
        }

    def forward(self, x, mask_features, mask=None):
        assert len(x) == self.num_feature_levels
        del mask
        # the fused path is written for the post-norm blocks every DVIS config uses; pre-norm takes the generic loop
        if _fast_path(mask_features) and not self.materialize_aux_masks and self.hidden_dim % 128 == 0 and not self.pre_norm:
            return self._forward_fast(x, mask_features)
        fast = _fast_path(mask_features) and not self.materialize_aux_masks
        src, pos, size_list = [], [], []
        for i in range(self.num_feature_levels):
            H, W = x[i].shape[-2:]
            size_list.append((H, W))
            pos.append(self._pos(H, W, x[i].device))
            s = self.input_proj[i](x[i])
            s = s.flatten(2).float() + self.level_embed.weight[i][None, :, None]
            src.append(s.permute(2, 0, 1))                                    # (hw, B, C)
        bs = src[0].shape[1]
        query_embed = self.query_embed.weight.unsqueeze(1).repeat(1, bs, 1)
        output = self.query_feat.weight.unsqueeze(1).repeat(1, bs, 1)
        predictions_class, predictions_mask = [], []

        if fast:
            # mask features resized once to each level grid (3 small maps) for the attention masks
            level_feats = [F.interpolate(mask_features.float(), size=sz, mode="bilinear", align_corners=False)
                           .to(ops._MASK_OPERAND[0]).contiguous(memory_format=torch.channels_last) for sz in size_list]
            attn_mask = self._heads_lowres(output, level_feats[0])
        else:
            c, m, attn_mask = self.forward_prediction_heads(output, mask_features, size_list[0])
            predictions_class.append(c)
            predictions_mask.append(m)

        for i in range(self.num_layers):
            level_index = i % self.num_feature_levels
            if fast:
                am = attn_mask.clone()
                am[am.all(-1)] = False                                          # py:297
                am = am[:, None].expand(-1, self.num_heads, -1, -1).flatten(0, 1)
            else:
                am = attn_mask
                am[torch.where(am.sum(-1) == am.shape[-1])] = False
            output = self.transformer_cross_attention_layers[i](output, src[level_index], memory_mask=am,
                                                                memory_key_padding_mask=None, pos=pos[level_index],
                                                                query_pos=query_embed)
            output = self.transformer_self_attention_layers[i](output, tgt_mask=None, tgt_key_padding_mask=None,
                                                               query_pos=query_embed)
            output = self.transformer_ffn_layers[i](output)
            nxt = (i + 1) % self.num_feature_levels
            if fast:
                if i + 1 < self.num_layers:
                    attn_mask = self._heads_lowres(output, level_feats[nxt])
            else:
                c, m, attn_mask = self.forward_prediction_heads(output, mask_features, size_list[nxt])
                predictions_class.append(c)
                predictions_mask.append(m)

        if fast:   # final heads at full resolution (the only masks the meta-architecture keeps)
            decoder_output = self.decoder_norm(output.float()).transpose(0, 1)
            predictions_class.append(linear(self.class_embed, decoder_output).float())
            predictions_mask.append(ops.mask_logits(self.mask_embed(decoder_output), mask_features, torch.float32))

        bt = predictions_mask[-1].shape[0]
        b = bt // self.num_frames if self.training else 1
        t = bt // b
        masks = [m.reshape(b, t, *m.shape[1:]).permute(0, 2, 1, 3, 4) for m in predictions_mask]      # b q t h w
        classes = [c.reshape(b, t, *c.shape[1:]) for c in predictions_class]                          # b t q c
        output = output.float()
        normed = self.decoder_norm(output)
        reid = self.reid_embed(normed).float()
        to_bctq = lambda z: z.reshape(z.shape[0], b, t, z.shape[-1]).permute(1, 3, 2, 0)            # q (b t) c -> b c t q
        pred_embds, reid_e, no_norm = to_bctq(normed), to_bctq(reid), to_bctq(output)
        return {
            "pred_logits": classes[-1],
            "pred_masks": masks[-1],
            "aux_outputs": [{"pred_logits": a, "pred_masks": m} for a, m in zip(classes[:-1], masks[:-1])],
            "pred_embds": torch.cat([pred_embds, reid_e], dim=1),
            "pred_embds_without_norm": torch.cat([no_norm, reid_e], dim=1),
            "pred_reid_embed": reid_e,
            "mask_features": mask_features,
        }


if _HAVE_REG:  # pragma: no cover
    VideoMultiScaleMaskedTransformerDecoder_dvisPlus = TRANSFORMER_DECODER_REGISTRY.register()(
        VideoMultiScaleMaskedTransformerDecoder_dvisPlus)
