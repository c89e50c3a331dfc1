"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Torch-CPU restatement of the reference's module-level hot path.

Every function is *functional*: it takes a flat ``sd`` dict of weights keyed by the reference's own
parameter names (so a reference ``state_dict()`` plugs in unchanged) plus a ``prefix``, and uses only
basic torch ops on whatever device/dtype the inputs live on (CPU fp32 in practice).  It is the checker
for the CUDA path and the `cpu_baseline` / `--impl reference` arm of bench.py -- the same algorithm
the reference executes on CPU (its `ms_deform_attn_core_pytorch` F.grid_sample path, einsum mask head,
multi-head attention blocks).  `dvis_plus_b200/` never imports this file.

Parity pin: tests/test_oracle.py compares each function with golden outputs of the unmodified reference
modules (tests/golden/*.pt, produced by tests/golden/make_golden.py in the build container).

Citations: P = /root/reference/DVIS_Plus, OPS = P/mask2former/modeling/pixel_decoder/ops.
"""
import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------
# small building blocks
# --------------------------------------------------------------------------------------------------


def linear(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def layer_norm(sd, prefix, x, eps=1e-5):
    w = sd[prefix + ".weight"]
    return F.layer_norm(x, (w.numel(),), w, sd[prefix + ".bias"], eps)


def mlp(sd, prefix, x, num_layers=3):
    """MLP with ReLU between layers (P/mask2former_video/.../video_mask2former_transformer_decoder.py:193-205)."""
    for i in range(num_layers):
        x = linear(sd, f"{prefix}.layers.{i}", x)
        if i < num_layers - 1:
            x = torch.relu(x)
    return x


def multihead_attention(sd, prefix, query, key, value, nheads, attn_mask=None):
    """What nn.MultiheadAttention computes for (L, B, E) inputs, dropout 0, returning output [0].

    Used by SelfAttentionLayer / CrossAttentionLayer / ReferringCrossAttentionLayer
    (…/video_mask2former_transformer_decoder.py:23,46,81,104; P/dvis_Plus/tracker.py:19,45).
    attn_mask: optional bool (B*nheads, Lq, Lk), True = may NOT attend.
    """
    Lq, B, E = query.shape
    Lk = key.shape[0]
    dh = E // nheads
    w, b = sd[prefix + ".in_proj_weight"], sd[prefix + ".in_proj_bias"]
    q = F.linear(query, w[:E], b[:E])
    k = F.linear(key, w[E:2 * E], b[E:2 * E])
    v = F.linear(value, w[2 * E:], b[2 * E:])
    q = q.reshape(Lq, B * nheads, dh).transpose(0, 1) * (1.0 / math.sqrt(dh))
    k = k.reshape(Lk, B * nheads, dh).transpose(0, 1)
    v = v.reshape(Lk, B * nheads, dh).transpose(0, 1)
    logits = torch.bmm(q, k.transpose(1, 2))
    if attn_mask is not None:
        logits = logits.masked_fill(attn_mask, float("-inf"))
    p = torch.softmax(logits, dim=-1)
    o = torch.bmm(p, v).transpose(0, 1).reshape(Lq, B, E)
    return F.linear(o, sd[prefix + ".out_proj.weight"], sd[prefix + ".out_proj.bias"])


def self_attention_layer(sd, prefix, tgt, nheads, query_pos=None):
    """Post-norm SelfAttentionLayer.forward_post (…/video_mask2former_transformer_decoder.py:40-50)."""
    qk = tgt if query_pos is None else tgt + query_pos
    t2 = multihead_attention(sd, prefix + ".self_attn", qk, qk, tgt, nheads)
    return layer_norm(sd, prefix + ".norm", tgt + t2)


def cross_attention_layer(sd, prefix, tgt, memory, nheads, memory_mask=None, pos=None, query_pos=None, identity=None):
    """Post-norm CrossAttentionLayer.forward_post (:98-111); with `identity` given it is the tracker's
    ReferringCrossAttentionLayer.forward_post (P/dvis_Plus/tracker.py:34-53) where key != value."""
    q = tgt if query_pos is None else tgt + query_pos
    if isinstance(memory, tuple):
        key, val = memory
    else:
        key = val = memory
    k = key if pos is None else key + pos
    t2 = multihead_attention(sd, prefix + ".multihead_attn", q, k, val, nheads, attn_mask=memory_mask)
    res = tgt if identity is None else identity
    return layer_norm(sd, prefix + ".norm", res + t2)


def ffn_layer(sd, prefix, tgt):
    """Post-norm FFNLayer.forward_post (:163-167), ReLU."""
    t2 = linear(sd, prefix + ".linear2", torch.relu(linear(sd, prefix + ".linear1", tgt)))
    return layer_norm(sd, prefix + ".norm", tgt + t2)


def sine_position_embedding(x, num_pos_feats, temperature=10000.0):
    """PositionEmbeddingSine(normalize=True) on an all-valid map
    (P/mask2former/modeling/transformer_decoder/position_encoding.py:29-52).  x: (N, C, H, W)."""
    N, _, H, W = x.shape
    scale, eps = 2 * math.pi, 1e-6
    ys = torch.arange(1, H + 1, dtype=torch.float32, device=x.device) / (H + eps) * scale
    xs = torch.arange(1, W + 1, dtype=torch.float32, device=x.device) / (W + eps) * scale
    i = torch.arange(num_pos_feats, dtype=torch.float32, device=x.device)
    dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode="floor") / num_pos_feats)

    def enc(v):  # (n,) -> (n, num_pos_feats): sin on even slots, cos on odd slots
        a = v[:, None] / dim_t
        return torch.stack((a[:, 0::2].sin(), a[:, 1::2].cos()), dim=2).flatten(1)

    py = enc(ys)[:, None, :].expand(H, W, num_pos_feats)
    px = enc(xs)[None, :, :].expand(H, W, num_pos_feats)
    return torch.cat((py, px), dim=2).permute(2, 0, 1)[None].expand(N, -1, -1, -1)


# --------------------------------------------------------------------------------------------------
# multi-scale deformable attention
# --------------------------------------------------------------------------------------------------


def msda_core(value, spatial_shapes, sampling_locations, attention_weights):
    """The reference's CPU path for the op: per level F.grid_sample(bilinear, zeros, align_corners=False)
    then a weighted sum over L*P (OPS/functions/ms_deform_attn_func.py:52-72).

    value (N,S,M,D), spatial_shapes list[(H,W)], sampling_locations (N,Lq,M,L,P,2) in [0,1] (x,y),
    attention_weights (N,Lq,M,L,P) -> (N, Lq, M*D).
    """
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    shapes = [(int(h), int(w)) for h, w in spatial_shapes]
    grids = 2 * sampling_locations - 1
    start, taps = 0, []
    for lvl, (H, W) in enumerate(shapes):
        v = value[:, start:start + H * W].permute(0, 2, 3, 1).reshape(N * M, D, H, W)
        g = grids[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(N * M, Lq, P, 2)
        taps.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))
        start += H * W
    taps = torch.stack(taps, dim=3).reshape(N * M, D, Lq, L * P)
    w = attention_weights.permute(0, 2, 1, 3, 4).reshape(N * M, 1, Lq, L * P)
    out = (taps * w).sum(-1).reshape(N, M * D, Lq)
    return out.transpose(1, 2).contiguous()


def ms_deform_attn(sd, prefix, query, reference_points, input_flatten, spatial_shapes, n_heads, n_levels, n_points,
                   input_padding_mask=None, return_parts=False):
    """MSDeformAttn.forward (OPS/modules/ms_deform_attn.py:82-125)."""
    N, Lq, C = query.shape
    S = input_flatten.shape[1]
    value = linear(sd, prefix + ".value_proj", input_flatten)
    if input_padding_mask is not None:
        value = value.masked_fill(input_padding_mask[..., None], 0.0)
    value = value.view(N, S, n_heads, C // n_heads)
    off = linear(sd, prefix + ".sampling_offsets", query).view(N, Lq, n_heads, n_levels, n_points, 2)
    aw = linear(sd, prefix + ".attention_weights", query).view(N, Lq, n_heads, n_levels * n_points)
    aw = torch.softmax(aw, -1).view(N, Lq, n_heads, n_levels, n_points)
    shapes_t = torch.as_tensor([[int(h), int(w)] for h, w in spatial_shapes], dtype=query.dtype, device=query.device)
    if reference_points.shape[-1] == 2:
        normalizer = torch.stack([shapes_t[:, 1], shapes_t[:, 0]], -1)  # (W, H) per level
        loc = reference_points[:, :, None, :, None, :] + off / normalizer[None, None, None, :, None, :]
    elif reference_points.shape[-1] == 4:
        loc = reference_points[:, :, None, :, None, :2] + off / n_points * reference_points[:, :, None, :, None, 2:] * 0.5
    else:
        raise ValueError("Last dim of reference_points must be 2 or 4")
    core = msda_core(value, spatial_shapes, loc, aw)
    out = linear(sd, prefix + ".output_proj", core)
    if return_parts:
        return out, dict(value=value, sampling_locations=loc, attention_weights=aw, core=core)
    return out


def encoder_reference_points(spatial_shapes, N, device, dtype=torch.float32):
    """get_reference_points with all-ones valid ratios (P/mask2former/modeling/pixel_decoder/msdeformattn.py:141-153):
    pixel centres, normalised, identical for every level.  -> (N, S, L, 2) as (x, y)."""
    pts = []
    for H, W in spatial_shapes:
        H, W = int(H), int(W)
        ys = (torch.arange(H, dtype=dtype, device=device) + 0.5) / H
        xs = (torch.arange(W, dtype=dtype, device=device) + 0.5) / W
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        pts.append(torch.stack((gx.reshape(-1), gy.reshape(-1)), -1))
    pts = torch.cat(pts, 0)
    return pts[None, :, None, :].expand(N, -1, len(spatial_shapes), -1)


def encoder_layer(sd, prefix, src, pos, reference_points, spatial_shapes, n_heads, n_levels, n_points):
    """MSDeformAttnTransformerEncoderLayer.forward (msdeformattn.py:116-131), dropout 0."""
    a = ms_deform_attn(sd, prefix + ".self_attn", src + pos, reference_points, src, spatial_shapes, n_heads, n_levels, n_points)
    src = layer_norm(sd, prefix + ".norm1", src + a)
    f = linear(sd, prefix + ".linear2", torch.relu(linear(sd, prefix + ".linear1", src)))
    return layer_norm(sd, prefix + ".norm2", src + f)


def _conv_gn(sd, prefix_conv, prefix_norm, x, padding=0, groups=32, relu=False):
    x = F.conv2d(x, sd[prefix_conv + ".weight"], sd.get(prefix_conv + ".bias"), padding=padding)
    x = F.group_norm(x, groups, sd[prefix_norm + ".weight"], sd[prefix_norm + ".bias"], 1e-5)
    return torch.relu(x) if relu else x


def pixel_decoder_forward_features(sd, features, *, transformer_in_features=("res3", "res4", "res5"),
                                   in_features=("res2", "res3", "res4", "res5"), strides=(4, 8, 16, 32),
                                   n_heads=8, n_points=4, num_layers=6, common_stride=4, prefix=""):
    """MSDeformAttnPixelDecoder.forward_features (msdeformattn.py:314-358) incl. the encoder-only transformer
    (:61-89,155-161).  Returns (mask_features, out[0], multi_scale_features[:3])."""
    conv_dim = sd[prefix + "mask_features.weight"].shape[1]
    L = len(transformer_in_features)
    srcs, poss = [], []
    for idx, f in enumerate(transformer_in_features[::-1]):
        x = features[f].float()
        srcs.append(_conv_gn(sd, f"{prefix}input_proj.{idx}.0", f"{prefix}input_proj.{idx}.1", x))
        poss.append(sine_position_embedding(x, conv_dim // 2))
    shapes = [(s.shape[2], s.shape[3]) for s in srcs]
    N = srcs[0].shape[0]
    src = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)
    pos = torch.cat([p.flatten(2).transpose(1, 2) + sd[prefix + "transformer.level_embed"][lvl].view(1, 1, -1)
                     for lvl, p in enumerate(poss)], 1)
    ref = encoder_reference_points(shapes, N, src.device)
    out = src
    for i in range(num_layers):
        out = encoder_layer(sd, f"{prefix}transformer.encoder.layers.{i}", out, pos, ref, shapes, n_heads, L, n_points)
    outs, start = [], 0
    for (H, W) in shapes:
        outs.append(out[:, start:start + H * W].transpose(1, 2).reshape(N, conv_dim, H, W))
        start += H * W
    tf_strides = [strides[in_features.index(f)] for f in transformer_in_features]
    num_fpn = int(math.log2(min(tf_strides)) - math.log2(common_stride))
    fpn_feats = list(in_features[:num_fpn])[::-1]
    for idx, f in enumerate(fpn_feats):
        k = num_fpn - idx  # module names adapter_k / layer_k were created low-stride first (msdeformattn.py:262-286)
        x = features[f].float()
        lat = _conv_gn(sd, f"{prefix}adapter_{k}", f"{prefix}adapter_{k}.norm", x)
        y = lat + F.interpolate(outs[-1], size=lat.shape[-2:], mode="bilinear", align_corners=False)
        outs.append(_conv_gn(sd, f"{prefix}layer_{k}", f"{prefix}layer_{k}.norm", y, padding=1, relu=True))
    mask_features = F.conv2d(outs[-1], sd[prefix + "mask_features.weight"], sd[prefix + "mask_features.bias"])
    return mask_features, outs[0], outs[:3]


# --------------------------------------------------------------------------------------------------
# mask head / segmenter predictor
# --------------------------------------------------------------------------------------------------


def mask_logits(mask_embed, mask_features):
    """einsum "bqc,bchw->bqhw" (P/dvis_Plus/video_mask2former_transformer_decoder.py:363)."""
    return torch.einsum("bqc,bchw->bqhw", mask_embed, mask_features)


def prediction_heads(sd, prefix, output, mask_features, target_size, num_heads):
    """forward_prediction_heads (decoder.py:358-374). output (Q,B,C) -> class (B,Q,K+1), masks (B,Q,H,W),
    bool attn_mask (B*heads, Q, h*w) where True = may not attend."""
    d = layer_norm(sd, prefix + "decoder_norm", output).transpose(0, 1)
    cls = linear(sd, prefix + "class_embed", d)
    me = mlp(sd, prefix + "mask_embed", d)
    masks = mask_logits(me, mask_features)
    am = F.interpolate(masks, size=target_size, mode="bilinear", align_corners=False)
    am = (am.sigmoid().flatten(2).unsqueeze(1).repeat(1, num_heads, 1, 1).flatten(0, 1) < 0.5)
    return cls, masks, am


def predictor_forward(sd, x, mask_features, *, num_heads=8, num_layers=9, prefix=""):
    """VideoMultiScaleMaskedTransformerDecoder_dvisPlus.forward at eval (decoder.py:258-356).
    x: list of 3 feature maps (B,C,h,w) low->high res; returns the reference's output dict (b=1, t=B)."""
    C = sd[prefix + "query_feat.weight"].shape[1]
    src, pos, sizes = [], [], []
    for i in range(3):
        sizes.append(tuple(x[i].shape[-2:]))
        pos.append(sine_position_embedding(x[i], C // 2).flatten(2).permute(2, 0, 1))
        s = x[i]
        if (prefix + f"input_proj.{i}.weight") in sd:
            s = F.conv2d(s, sd[prefix + f"input_proj.{i}.weight"], sd[prefix + f"input_proj.{i}.bias"])
        s = s.flatten(2) + sd[prefix + "level_embed.weight"][i][None, :, None]
        src.append(s.permute(2, 0, 1))
    B = src[0].shape[1]
    query_embed = sd[prefix + "query_embed.weight"].unsqueeze(1).repeat(1, B, 1)
    output = sd[prefix + "query_feat.weight"].unsqueeze(1).repeat(1, B, 1)
    classes, masks = [], []
    c, m, am = prediction_heads(sd, prefix, output, mask_features, sizes[0], num_heads)
    classes.append(c)
    masks.append(m)
    for i in range(num_layers):
        lvl = i % 3
        am = am.clone()
        am[am.all(-1)] = False  # fully-masked query rows attend everywhere (decoder.py:297)
        output = cross_attention_layer(sd, f"{prefix}transformer_cross_attention_layers.{i}", output, src[lvl], num_heads,
                                       memory_mask=am, pos=pos[lvl], query_pos=query_embed)
        output = self_attention_layer(sd, f"{prefix}transformer_self_attention_layers.{i}", output, num_heads, query_pos=query_embed)
        output = ffn_layer(sd, f"{prefix}transformer_ffn_layers.{i}", output)
        c, m, am = prediction_heads(sd, prefix, output, mask_features, sizes[(i + 1) % 3], num_heads)
        classes.append(c)
        masks.append(m)
    normed = layer_norm(sd, prefix + "decoder_norm", output)
    if (prefix + "reid_embed.layers.0.weight") in sd:
        n_reid = sum(1 for k in sd if k.startswith(prefix + "reid_embed.layers.") and k.endswith(".weight"))
        reid = mlp(sd, prefix + "reid_embed", normed, n_reid)
    else:
        reid = normed
    to_bctq = lambda t: t.permute(2, 1, 0)[None].permute(0, 1, 2, 3)  # (q, t, c) -> (1, c, t, q)
    return {
        "pred_logits": classes[-1][None],                           # (1, t, q, K+1)
        "pred_masks": masks[-1].permute(1, 0, 2, 3)[None],          # (1, q, t, h, w)
        "all_masks": masks,
        "all_logits": classes,
        "pred_embds": torch.cat([to_bctq(normed), to_bctq(reid)], 1),
        "pred_embds_without_norm": torch.cat([to_bctq(output), to_bctq(reid)], 1),
        "pred_reid_embed": to_bctq(reid),
        "mask_features": mask_features,
    }


# --------------------------------------------------------------------------------------------------
# tracker / refiner (eval mode)
# --------------------------------------------------------------------------------------------------


def hungarian_match(ref_embds, cur_embds):
    """Noiser.match_embds (P/dvis_Plus/noiser.py:43-56): cosine cost, scipy linear_sum_assignment on the host."""
    from scipy.optimize import linear_sum_assignment
    r, c = ref_embds.detach()[:, 0, :], cur_embds.detach()[:, 0, :]
    r = r / (r.norm(dim=1)[:, None] + 1e-6)
    c = c / (c.norm(dim=1)[:, None] + 1e-6)
    cost = (1 - torch.mm(c, r.transpose(0, 1))).cpu()
    cost = torch.where(torch.isnan(cost), torch.zeros_like(cost), cost)
    return linear_sum_assignment(cost.transpose(0, 1))[1]


def tracker_forward(sd, frame_embeds, mask_features, frame_embeds_no_norm, *, num_heads=8, num_layers=6, prefix="",
                    state=None, with_masks=True):
    """ReferringTracker_noiser.forward at eval, noise off (P/dvis_Plus/tracker.py:187-357,368-380).

    frame_embeds / frame_embeds_no_norm (b=1, c, t, q); mask_features (1, t, c, h, w).
    `state` carries (last_outputs, last_frame_embeds) across windows (resume=True semantics, tracker.py:225,277).
    """
    fe = frame_embeds.permute(2, 3, 0, 1)           # (t, q, b, c)
    fn = frame_embeds_no_norm.permute(2, 3, 0, 1)
    T = fe.shape[0]
    last_outputs, last_frame_embeds = state if state is not None else (None, None)
    outs, refs, all_indices = [], [], []
    ca = lambda j: f"{prefix}transformer_cross_attention_layers.{j}"
    sa = lambda j: f"{prefix}transformer_self_attention_layers.{j}"
    ff = lambda j: f"{prefix}transformer_ffn_layers.{j}"
    for i in range(T):
        cur, cur_nn = fe[i], fn[i]
        first = last_outputs is None
        if first:
            indices = hungarian_match(cur, cur)
            reference = None
        else:
            reference = mlp(sd, prefix + "ref_proj", last_outputs[-1])
            indices = hungarian_match(last_frame_embeds, cur)
        all_indices.append(indices)
        idx = torch.as_tensor(indices, device=cur.device, dtype=torch.long)
        ms = [cur_nn[idx]]
        last_frame_embeds = cur[idx]
        for j in range(num_layers):
            if first:
                tgt = mlp(sd, prefix + "ref_proj", cur_nn if j == 0 else ms[-1])
            else:
                tgt = reference
            o = cross_attention_layer(sd, ca(j), tgt, (cur_nn, cur_nn), num_heads, identity=ms[-1])
            o = self_attention_layer(sd, sa(j), o, num_heads)
            o = ffn_layer(sd, ff(j), o)
            ms.append(o)
        refs.append(mlp(sd, prefix + "ref_proj", cur_nn) if first else reference)
        last_outputs = torch.stack(ms, 0)
        outs.append(last_outputs[-1])
    outputs = torch.stack(outs, 0)                      # (t, q, b, c)   (last layer only, eval)
    references = torch.stack(refs, 0)                   # (t, q, b, c)
    dec = layer_norm(sd, prefix + "decoder_norm", outputs).permute(2, 0, 1, 3)   # (b, t, q, c)
    cls = linear(sd, prefix + "class_embed", torch.cat([references.permute(2, 0, 1, 3), dec], -1))  # (b, t, q, K+1)
    res = {
        "pred_logits": cls,
        "pred_embds": outputs.permute(2, 3, 0, 1),        # (b, c, t, q)
        "pred_references": references.permute(2, 3, 0, 1),
        "indices": all_indices,
        "state": (last_outputs, last_frame_embeds),
    }
    if with_masks:
        mf = mask_features
        b, t = mf.shape[:2]
        mf = F.conv2d(mf.flatten(0, 1), sd[prefix + "mask_feature_proj.weight"], sd[prefix + "mask_feature_proj.bias"]).reshape(mf.shape)
        me = mlp(sd, prefix + "mask_embed", dec)
        res["pred_masks"] = torch.einsum("btqc,btchw->bqthw", me, mf)
    return res


def refiner_forward(sd, instance_embeds, frame_embeds, mask_features, *, num_heads=8, num_layers=6, prefix="", with_masks=True):
    """TemporalRefiner.forward at eval (P/dvis_Plus/refiner.py:91-158,169-210).
    instance_embeds, frame_embeds (b, c, t, q); mask_features (b, t, c, h, w)."""
    b, c, t, q = instance_embeds.shape
    out = instance_embeds
    mem = frame_embeds.permute(3, 0, 2, 1).flatten(1, 2)      # (q, b*t, c)
    for i in range(num_layers):
        x = out.permute(2, 0, 3, 1).flatten(1, 2)             # (t, b*q, c)
        x = self_attention_layer(sd, f"{prefix}transformer_time_self_attention_layers.{i}", x, num_heads)
        x = x.permute(1, 2, 0)                                # (b*q, c, t)
        cp = f"{prefix}conv_short_aggregate_layers.{i}"
        y = F.conv1d(F.pad(x, (2, 2), mode="replicate"), sd[cp + ".0.weight"], sd[cp + ".0.bias"])
        y = F.conv1d(F.pad(torch.relu(y), (1, 1), mode="replicate"), sd[cp + ".2.weight"], sd[cp + ".2.bias"])
        x = layer_norm(sd, f"{prefix}conv_norms.{i}", (y + x).transpose(1, 2)).transpose(1, 2)
        x = x.reshape(b, q, c, t).permute(1, 0, 3, 2).flatten(1, 2)   # (q, b*t, c)
        x = self_attention_layer(sd, f"{prefix}transformer_obj_self_attention_layers.{i}", x, num_heads)
        x = cross_attention_layer(sd, f"{prefix}transformer_cross_attention_layers.{i}", x, mem, num_heads)
        x = ffn_layer(sd, f"{prefix}transformer_ffn_layers.{i}", x)
        out = x.reshape(q, b, t, c).permute(1, 3, 2, 0)       # (b, c, t, q)
    dec = layer_norm(sd, prefix + "decoder_norm", out.permute(0, 2, 3, 1))        # (b, t, q, c)
    act = linear(sd, prefix + "activation_proj", dec).softmax(dim=1)             # softmax over t (refiner.py:205)
    fused = (dec * act).sum(dim=1, keepdim=True).expand(-1, t, -1, -1)
    res = {
        "pred_logits": linear(sd, prefix + "class_embed", fused),                # (b, t, q, K+1)
        "pred_embds": dec.permute(0, 3, 1, 2),                                    # (b, c, t, q)
    }
    if with_masks:
        me = mlp(sd, prefix + "mask_embed", dec)
        res["pred_masks"] = torch.einsum("btqc,btchw->bqthw", me, mask_features)
    return res
