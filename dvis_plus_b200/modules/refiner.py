"""Drop-in for TemporalRefiner (P/dvis_Plus/refiner.py:6-226): same constructor keywords, parameter names and
output dictionary.  At inference the windowed mask prediction (py:169-194) keeps mask features on the device and runs
the tcgen05 mask GEMM; the reference's host round trips (`.to(device)` py:188, `.cpu()` py:191) exist only to save
GPU memory on 16-40 GB parts and are not reproduced (180 GB HBM per B200)."""
import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from .blocks import MLP, CrossAttentionLayer, FFNLayer, SelfAttentionLayer, _fast_path, add_norm, linear
from .precision import gemm_dtype


class TemporalRefiner(nn.Module):
    def __init__(self, hidden_channel=256, feedforward_channel=2048, num_head=8, decoder_layer_num=6, mask_dim=256,
                 class_num=25, windows=5):
        super().__init__()
        self.windows = windows
        self.num_heads = num_head
        self.num_layers = decoder_layer_num
        self.transformer_obj_self_attention_layers = nn.ModuleList()
        self.transformer_time_self_attention_layers = nn.ModuleList()
        self.transformer_cross_attention_layers = nn.ModuleList()
        self.transformer_ffn_layers = nn.ModuleList()
        self.conv_short_aggregate_layers = nn.ModuleList()
        self.conv_norms = nn.ModuleList()
        for _ in range(self.num_layers):
            self.transformer_time_self_attention_layers.append(SelfAttentionLayer(hidden_channel, num_head, 0.0))
            self.conv_short_aggregate_layers.append(nn.Sequential(
                nn.Conv1d(hidden_channel, hidden_channel, kernel_size=5, stride=1, padding="same", padding_mode="replicate"),
                nn.ReLU(inplace=True),
                nn.Conv1d(hidden_channel, hidden_channel, kernel_size=3, stride=1, padding="same", padding_mode="replicate")))
            self.conv_norms.append(nn.LayerNorm(hidden_channel))
            self.transformer_obj_self_attention_layers.append(SelfAttentionLayer(hidden_channel, num_head, 0.0))
            self.transformer_cross_attention_layers.append(CrossAttentionLayer(hidden_channel, num_head, 0.0))
            self.transformer_ffn_layers.append(FFNLayer(hidden_channel, feedforward_channel, 0.0))
        self.decoder_norm = nn.LayerNorm(hidden_channel)
        self.class_embed = nn.Linear(hidden_channel, class_num + 1)
        self.mask_embed = MLP(hidden_channel, hidden_channel, mask_dim, 3)
        self.activation_proj = nn.Linear(hidden_channel, 1)

    def _short_conv(self, i, x):
        """x (bq, c, t) -> conv(k5) -> ReLU -> conv(k3), replicate padding (py:44-52,116-119)."""
        c5, c3 = self.conv_short_aggregate_layers[i][0], self.conv_short_aggregate_layers[i][2]
        dt = gemm_dtype() if _fast_path(x) else x.dtype
        x = x.to(dt)
        rep = lambda z, k: torch.cat([z[..., :1].expand(-1, -1, k), z, z[..., -1:].expand(-1, -1, k)], dim=-1)   # replicate pad
        y = F.conv1d(rep(x, 2), c5.weight.to(dt), c5.bias.to(dt))
        y = F.conv1d(rep(F.relu(y), 1), c3.weight.to(dt), c3.bias.to(dt))
        return y

    def _short_conv_rows(self, i, xt):
        """The same two convolutions over time as GEMMs on rows (bq, t): xt (bq, t, c) in the GEMM dtype -> (bq, t, c).
        A[(q, t), (k, ci)] = x[q, clamp(t + k - pad), ci] is one gather per convolution (replicate padding = clamped indices)
        and the convolution one cuBLAS GEMM with the ReLU in its epilogue -- instead of cat (padding) + cuDNN's NCHW <-> NHWC
        conversion kernels around each Conv1d (15 -> 7 launches per layer; py:44-52,116-119)."""
        c5, c3 = self.conv_short_aggregate_layers[i][0], self.conv_short_aggregate_layers[i][2]
        dt, T = xt.dtype, xt.shape[1]
        key = (dt, c5.weight._version, c3.weight._version, c5.weight.data_ptr(), str(xt.device), T)
        cache = self.__dict__.setdefault("_conv_rows_cache", {})
        if cache.get(i, (None,))[0] != key:
            rows = lambda conv: conv.weight.detach().permute(0, 2, 1).reshape(conv.out_channels, -1).to(dt).contiguous()   # (co, k*ci)
            idx = lambda k: (torch.arange(T, device=xt.device)[:, None] + torch.arange(k, device=xt.device)[None] - k // 2).clamp_(0, T - 1)
            cache[i] = (key, rows(c5), c5.bias.detach().to(dt), rows(c3), c3.bias.detach().to(dt), idx(5), idx(3))
        _, w5, b5, w3, b3, i5, i3 = cache[i]
        bq, _, C = xt.shape
        h = torch._addmm_activation(b5, xt[:, i5].reshape(bq * T, 5 * C), w5.t())            # ReLU in the GEMM epilogue
        return torch.addmm(b3, h.view(bq, T, -1)[:, i3].reshape(bq * T, 3 * h.shape[-1]), w3.t()).view(bq, T, -1)

    # opt-in (bf16 mode, batch 1): the layers run on csrc/small_linear.cu + csrc/flash_attn.cu only (14 launches each).
    # Parity-green but slower than the library GEMMs at T*Q = 3 200 rows (4.7 ms vs 1.8 ms per clip): off by default
    use_fused_kernels = False
    _fast = None

    def _stacked(self):
        """bf16 weights (nn.Linear layout; Conv1d as (C_out, k*C_in), tap-major), fp32 biases / LayerNorm parameters."""
        params = list(self.parameters())
        key = (tuple(p._version for p in params), params[0].data_ptr())
        if self._fast is None or self._fast["key"] != key:
            w = lambda t: t.detach().to(torch.bfloat16).contiguous()
            v = lambda t: t.detach().float().contiguous()
            ln = lambda n: (v(n.weight), v(n.bias))
            conv = lambda c: (w(c.weight.permute(0, 2, 1).reshape(c.weight.shape[0], -1)), v(c.bias))
            C = self.decoder_norm.normalized_shape[0]
            ta, oa, ca, ff = (self.transformer_time_self_attention_layers, self.transformer_obj_self_attention_layers,
                              self.transformer_cross_attention_layers, self.transformer_ffn_layers)
            L = self.num_layers
            f = dict(key=key, C=C)
            f["w_kv"] = w(torch.cat([ca[i].multihead_attn.in_proj_weight[C:] for i in range(L)], 0))      # (L*2C, C)
            f["b_kv"] = v(torch.cat([ca[i].multihead_attn.in_proj_bias[C:] for i in range(L)], 0))
            f["layers"] = [dict(
                t_qkv=(w(ta[i].self_attn.in_proj_weight), v(ta[i].self_attn.in_proj_bias)),
                t_o=(w(ta[i].self_attn.out_proj.weight), v(ta[i].self_attn.out_proj.bias)), ln_t=ln(ta[i].norm),
                c5=conv(self.conv_short_aggregate_layers[i][0]), c3=conv(self.conv_short_aggregate_layers[i][2]),
                ln_c=ln(self.conv_norms[i]),
                o_qkv=(w(oa[i].self_attn.in_proj_weight), v(oa[i].self_attn.in_proj_bias)),
                o_o=(w(oa[i].self_attn.out_proj.weight), v(oa[i].self_attn.out_proj.bias)), ln_o=ln(oa[i].norm),
                x_q=(w(ca[i].multihead_attn.in_proj_weight[:C]), v(ca[i].multihead_attn.in_proj_bias[:C])),
                x_o=(w(ca[i].multihead_attn.out_proj.weight), v(ca[i].multihead_attn.out_proj.bias)), ln_x=ln(ca[i].norm),
                f1=(w(ff[i].linear1.weight), v(ff[i].linear1.bias)), f2=(w(ff[i].linear2.weight), v(ff[i].linear2.bias)),
                ln_f=ln(ff[i].norm)) for i in range(L)]
            self._fast = f
        return self._fast

    def _refine_fused(self, instance_embeds, frame_embeds):
        """py:104-145 for batch 1 on libdvis_b200 kernels only, tokens kept as (t, q, c) rows throughout (no permutes): per
        layer QKV / attention-over-time / out-proj, LayerNorm, Conv1d k5 + ReLU and k3 as tap-gather GEMMs, QKV /
        attention-over-objects / out-proj, Q-proj / cross-attention to the frame queries / out-proj, FFN1, FFN2 -- 14
        launches.  LayerNorms are folded into the consumers' prologues (csrc/small_linear.cu) except the one the
        convolution reads through its time taps."""
        f = self._stacked()
        L, C, H = self.num_layers, f["C"], self.num_heads
        dh = C // H
        _, _, T, Q = instance_embeds.shape
        scale = 1.0 / (dh ** 0.5)
        eps = self.decoder_norm.eps
        x0 = instance_embeds[0].permute(1, 2, 0).float().contiguous().view(T * Q, C)              # rows t*Q + q
        mem = frame_embeds[0].permute(1, 2, 0).to(torch.bfloat16).contiguous().view(T * Q, C)
        kv = ops.linear_small(f["w_kv"], f["b_kv"], x=mem)[1].view(T, Q, L, 2, H, dh)              # cross-attention keys / values
        src, pending, outs = x0, None, []
        for i in range(L):
            p = f["layers"][i]
            # time self-attention: every query attends over its own T frames (py:105-113)
            _, qkv, _, x = ops.linear_small(*p["t_qkv"], src0=src, ln1=pending, eps=eps, want_side1=pending is not None)
            if pending is None:
                x = src
            else:
                outs.append(x)                                                                   # layer i-1's output
            qkv = qkv.view(T, Q, 3, H, dh).permute(1, 0, 2, 3, 4)                                  # (q, t, 3, H, dh) view
            o = torch.empty((T, Q, C), dtype=torch.bfloat16, device=x0.device)
            ops.flash_attn(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], scale, out=o.permute(1, 0, 2))
            pre, _, _, _ = ops.linear_small(*p["t_o"], x=o.view(T * Q, C), residual=x, out_f32=True, out_bf16=False)
            # short-term convolution over time (py:116-119): the taps read neighbouring frames, so this LN is materialised
            x32, x16, _ = ops.add_layernorm(pre, None, *p["ln_t"], eps, lp_dtype=torch.bfloat16)
            _, h, _, _ = ops.linear_small(*p["c5"], x=x16, taps=5, tap_pad=2, tap_period=Q, tap_len=T, relu=True)
            pre, _, _, _ = ops.linear_small(*p["c3"], x=h, taps=3, tap_pad=1, tap_period=Q, tap_len=T, residual=x32, out_f32=True,
                                            out_bf16=False)
            # object self-attention within each frame (py:125-129)
            _, qkv, _, x = ops.linear_small(*p["o_qkv"], src0=pre, ln1=p["ln_c"], eps=eps, want_side1=True)
            qkv = qkv.view(T, Q, 3, H, dh)
            o = ops.flash_attn(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], scale)
            pre, _, _, _ = ops.linear_small(*p["o_o"], x=o.view(T * Q, C), residual=x, out_f32=True, out_bf16=False)
            # cross-attention to the same frame's segmenter queries (py:132-137)
            _, q, _, x = ops.linear_small(*p["x_q"], src0=pre, ln1=p["ln_o"], eps=eps, want_side1=True)
            o = ops.flash_attn(q.view(T, Q, H, dh), kv[:, :, i, 0], kv[:, :, i, 1], scale)
            pre, _, _, _ = ops.linear_small(*p["x_o"], x=o.view(T * Q, C), residual=x, out_f32=True, out_bf16=False)
            # FFN (py:140-142)
            _, h, _, x = ops.linear_small(*p["f1"], src0=pre, ln1=p["ln_x"], eps=eps, want_side1=True, relu=True)
            src, _, _, _ = ops.linear_small(*p["f2"], x=h, residual=x, out_f32=True, out_bf16=False)
            pending = p["ln_f"]
        outs.append(ops.add_layernorm(src, None, *pending, eps)[0])
        return torch.stack(outs, 0).view(L, T, Q, 1, C).permute(1, 0, 2, 3, 4)                     # (t, l, q, b, c)

    def _attend(self, q, k, v, scale, out=None):
        """(B, Lq, H, dh) x (B, Lk, H, dh) views -> (B, Lq, H*dh) bf16 (written into the view `out` when given): csrc/flash_attn.cu
        where it measured faster than cuDNN's SDPA (blocks.flash_attn_wins), else SDPA on the same strided views."""
        from .blocks import flash_attn_wins
        B, Lq, H, dh = q.shape
        if flash_attn_wins(B, H, Lq):
            return ops.flash_attn(q, k, v, scale, out=out)
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), scale=scale)   # (B, H, Lq, dh)
        o = o.transpose(1, 2).reshape(B, Lq, H * dh)
        if out is not None:
            out.copy_(o)
            return out
        return o

    def _refine_rows(self, instance_embeds, frame_embeds):
        """py:104-145 for batch 1 (bf16 inference) with the tokens kept as (t, q, c) rows throughout: the same cuBLAS GEMMs as the
        module path, but no (t, bq, c) <-> (q, bt, c) <-> (b, c, t, q) permutation copies between the blocks, attention on strided
        views of the packed QKV projections, the residual + LayerNorm kernel after every block, the convolutions over time as
        one row gather + one GEMM each (replicate padding = clamped frame index): 21 launches per layer."""
        f = self._stacked()
        L, C, H = self.num_layers, f["C"], self.num_heads
        dh = C // H
        _, _, T, Q = instance_embeds.shape
        scale = 1.0 / (dh ** 0.5)
        eps = self.decoder_norm.eps
        bf = torch.bfloat16
        dev = instance_embeds.device
        if "b16" not in f:                                            # bf16 copies of the biases for cuBLAS's epilogue
            f["b16"] = [{k: p[k][1].to(bf) for k in ("t_qkv", "t_o", "c5", "c3", "o_qkv", "o_o", "x_q", "x_o", "f1", "f2")} for p in f["layers"]]
            f["b_kv16"] = f["b_kv"].to(bf)
            f["taps"] = {}
        if (T, Q, str(dev)) not in f["taps"]:
            def rows(k):   # source row of (row (t, q), tap j): clamp(t + j - k // 2) * Q + q
                t = (torch.arange(T, device=dev)[:, None] + torch.arange(k, device=dev)[None] - k // 2).clamp_(0, T - 1)   # (T, k)
                return (t[:, None, :] * Q + torch.arange(Q, device=dev)[None, :, None]).reshape(T * Q, k)
            f["taps"][(T, Q, str(dev))] = (rows(5), rows(3))
        r5, r3 = f["taps"][(T, Q, str(dev))]

        def lin(i, name, x, relu=False):
            w, b = f["layers"][i][name][0], f["b16"][i][name]
            return torch._addmm_activation(b, x, w.t()) if relu else torch.addmm(b, x, w.t())
        ln = lambda y, res, g: ops.add_layernorm(y, res, g[0], g[1], eps, lp_dtype=bf)[:2]
        x32 = instance_embeds[0].permute(1, 2, 0).float().contiguous().view(T * Q, C)              # rows t*Q + q, fp32 stream
        x16 = x32.to(bf)
        mem = frame_embeds[0].permute(1, 2, 0).to(bf).contiguous().view(T * Q, C)
        kv = torch.addmm(f["b_kv16"], mem, f["w_kv"].t()).view(T, Q, L, 2, H, dh)                  # all layers' cross-attention K / V
        outs = []
        for i in range(L):
            p = f["layers"][i]
            # time self-attention: every query attends over its own T frames (py:105-113)
            qkv = lin(i, "t_qkv", x16).view(T, Q, 3, H, dh).permute(1, 0, 2, 3, 4)                 # (q, t, 3, H, dh) view
            o = torch.empty((T, Q, C), dtype=bf, device=dev)
            self._attend(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], scale, out=o.permute(1, 0, 2))  # written as (t, q, c) rows
            x32, x16 = ln(lin(i, "t_o", o.view(T * Q, C)), x32, p["ln_t"])
            # short-term convolution over time (py:116-119)
            h = lin(i, "c5", x16[r5].view(T * Q, 5 * C), relu=True)
            x32, x16 = ln(lin(i, "c3", h[r3].view(T * Q, 3 * C)), x32, p["ln_c"])
            # object self-attention within each frame (py:125-129)
            qkv = lin(i, "o_qkv", x16).view(T, Q, 3, H, dh)
            o = self._attend(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], scale)
            x32, x16 = ln(lin(i, "o_o", o.reshape(T * Q, C)), x32, p["ln_o"])
            # cross-attention to the same frame's segmenter queries (py:132-137)
            q = lin(i, "x_q", x16).view(T, Q, H, dh)
            o = self._attend(q, kv[:, :, i, 0], kv[:, :, i, 1], scale)
            x32, x16 = ln(lin(i, "x_o", o.reshape(T * Q, C)), x32, p["ln_x"])
            # FFN (py:140-142)
            x32, x16 = ln(lin(i, "f2", lin(i, "f1", x16, relu=True)), x32, p["ln_f"])
            outs.append(x32)
        return torch.stack(outs, 0).view(L, T, Q, 1, C).permute(1, 0, 2, 3, 4)                     # (t, l, q, b, c)

    use_row_layout = True     # bf16 inference, batch 1: _refine_rows (library GEMMs on (t, q, c) rows)

    def refine(self, instance_embeds, frame_embeds):
        """The 6 refinement layers (py:104-145).  (b, c, t, q) x2 -> stacked per-layer outputs (t, l, q, b, c), fp32."""
        n_batch, n_channel, n_frames, n_instance = instance_embeds.size()
        if (self.use_row_layout and not self.use_fused_kernels and not self.training and _fast_path(instance_embeds) and n_batch == 1
                and gemm_dtype() == torch.bfloat16 and n_channel // self.num_heads in (32, 64) and n_channel % 4 == 0):
            return self._refine_rows(instance_embeds, frame_embeds)
        if (self.use_fused_kernels and not self.training and _fast_path(instance_embeds) and n_batch == 1
                and gemm_dtype() == torch.bfloat16 and n_channel % 128 == 0 and n_channel <= 512
                and n_channel // self.num_heads in (32, 64)):
            return self._refine_fused(instance_embeds, frame_embeds)
        outputs = []
        output = instance_embeds.float()
        frame_embeds = frame_embeds.float().permute(3, 0, 2, 1).flatten(1, 2)        # (q, bt, c)
        for i in range(self.num_layers):
            output = output.permute(2, 0, 3, 1).flatten(1, 2)                        # (t, bq, c)
            output = self.transformer_time_self_attention_layers[i](output, tgt_mask=None, tgt_key_padding_mask=None, query_pos=None)
            if _fast_path(output) and not self.training:
                xt = output.permute(1, 0, 2).contiguous()                            # (bq, t, c): rows for the conv GEMMs and the residual
                y = self._short_conv_rows(i, xt.to(gemm_dtype()))
                output = add_norm(self.conv_norms[i], y, xt)                         # (bq, t, c)
                output = output.reshape(n_batch, n_instance, n_frames, n_channel).permute(1, 0, 2, 3).flatten(1, 2)   # (q, bt, c)
            else:
                output = output.permute(1, 2, 0)                                     # (bq, c, t)
                y = self._short_conv(i, output)
                output = add_norm(self.conv_norms[i], y.transpose(1, 2).contiguous(), output.transpose(1, 2).contiguous()).transpose(1, 2)
                output = output.reshape(n_batch, n_instance, n_channel, n_frames).permute(1, 0, 3, 2).flatten(1, 2)   # (q, bt, c)
            output = self.transformer_obj_self_attention_layers[i](output, tgt_mask=None, tgt_key_padding_mask=None, query_pos=None)
            output = self.transformer_cross_attention_layers[i](output, frame_embeds, memory_mask=None,
                                                                memory_key_padding_mask=None, pos=None, query_pos=None)
            output = self.transformer_ffn_layers[i](output)
            output = output.reshape(n_instance, n_batch, n_frames, n_channel).permute(1, 3, 2, 0)                 # (b, c, t, q)
            outputs.append(output)
        return torch.stack(outputs, dim=0).permute(3, 0, 4, 1, 2)                    # (l, b, c, t, q) -> (t, l, q, b, c)

    def predict_masks(self, outputs, mask_features):
        """Masks of the last layer for a slice of frames: outputs (t', l, q, b, c) of `refine`, mask_features
        (b, t', c, h, w) of the SAME frames -> (b, q, t', h, w).  Lets every rank of a frame-sharded clip finish only
        its own frames against its local mask features."""
        dec = self.decoder_norm(outputs[:, -1:]).permute(1, 3, 0, 2, 4)              # (1, b, t', q, c)
        return self._masks(self.mask_embed(dec).float(), mask_features)[-1]

    def forward(self, instance_embeds, frame_embeds, mask_features, with_masks=True):
        """instance_embeds, frame_embeds (b, c, t, q); mask_features (b, t, c, h, w)."""
        outputs = self.refine(instance_embeds, frame_embeds)
        outputs_class, outputs_masks = self.prediction(outputs, mask_features, with_masks=with_masks)
        outputs = self.decoder_norm(outputs)
        return {
            "pred_logits": outputs_class[-1].transpose(1, 2),                        # (b, t, q, c)
            "pred_masks": None if outputs_masks is None else outputs_masks[-1],       # (b, q, t, h, w)
            "aux_outputs": self._set_aux_loss(outputs_class, outputs_masks),
            "pred_embds": outputs[:, -1].permute(2, 3, 0, 1),                        # (b, c, t, q)
        }

    @torch.jit.unused
    def _set_aux_loss(self, outputs_class, outputs_seg_masks):
        if outputs_seg_masks is None:
            return [{"pred_logits": a.transpose(1, 2)} for a in outputs_class[:-1]]
        return [{"pred_logits": a.transpose(1, 2), "pred_masks": b} for a, b in zip(outputs_class[:-1], outputs_seg_masks[:-1])]

    mask_dtype = torch.float32   # dtype of the inference-path mask logits (torch.bfloat16 halves the write traffic)

    def _masks(self, mask_embed, mask_features):
        """einsum "lbtqc,btchw->lbqthw" (py:185-189,223)."""
        if _fast_path(mask_features):
            l, b, t, q, c = mask_embed.shape
            if b == 1 and q <= 256 and gemm_dtype() == torch.bfloat16:
                # one clip: the GEMM writes "q t h w" directly (no transposition of the (t, q, h, w) result)
                return torch.stack([ops.mask_logits_clip(mask_embed[li, 0], mask_features[0], self.mask_dtype)[None]
                                    for li in range(l)], 0) if l > 1 else \
                    ops.mask_logits_clip(mask_embed[0, 0], mask_features[0], self.mask_dtype)[None, None]
            feats = mask_features.flatten(0, 1)
            out = [ops.mask_logits(mask_embed[li].flatten(0, 1), feats, self.mask_dtype)
                   .reshape(b, t, q, *mask_features.shape[-2:]).permute(0, 2, 1, 3, 4) for li in range(l)]
            return torch.stack(out, 0)
        return torch.einsum("lbtqc,btchw->lbqthw", mask_embed.float(), mask_features.float())

    def windows_prediction(self, outputs, mask_features, windows=5, with_masks=True):
        """py:169-194, on-device: the window loop is kept (API) but every window's GEMM reads resident features."""
        iters = (outputs.size(0) + windows - 1) // windows
        outputs_classes, outputs_masks = [], []
        for i in range(iters):
            s, e = i * windows, (i + 1) * windows
            decoder_output = self.decoder_norm(outputs[s:e]).permute(1, 3, 0, 2, 4)    # (l, b, t, q, c)
            outputs_classes.append(decoder_output)
            if with_masks:
                mask_embed = self.mask_embed(decoder_output).float()
                outputs_masks.append(self._masks(mask_embed, mask_features[:, s:e].to(mask_embed.device)))
        outputs_classes = self.pred_class(torch.cat(outputs_classes, dim=2))
        return outputs_classes, (torch.cat(outputs_masks, dim=3) if with_masks else None)

    def pred_class(self, decoder_output):
        """py:196-210: softmax-over-time weighted mean of the queries, then the class head."""
        T = decoder_output.size(2)
        activation = linear(self.activation_proj, decoder_output).float().softmax(dim=2)
        class_output = (decoder_output * activation).sum(dim=2, keepdim=True).repeat(1, 1, T, 1, 1)
        return linear(self.class_embed, class_output).float().transpose(2, 3)

    def prediction(self, outputs, mask_features, with_masks=True):
        if self.training:
            decoder_output = self.decoder_norm(outputs).permute(1, 3, 0, 2, 4)
            outputs_class = self.pred_class(decoder_output)
            mask_embed = self.mask_embed(decoder_output)
            return outputs_class, torch.einsum("lbtqc,btchw->lbqthw", mask_embed.float(), mask_features.float())
        outputs = outputs[:, -1:]
        return self.windows_prediction(outputs, mask_features, windows=self.windows, with_masks=with_masks)
