"""Drop-in for ReferringTracker_noiser and Noiser (P/dvis_Plus/tracker.py:94-380, P/dvis_Plus/noiser.py).

Same constructor keywords, parameter names, recurrent state (`last_outputs`, `last_frame_embeds`,
`last_reference`, `_clear_memory`) and output dictionary.  Inference-path differences:
  * the 1x1 `mask_feature_proj` conv over all (T, C, H, W) mask features (py:199) is folded into the query side:
        einsum(me, W F + b) = einsum(W^T me, F) + me . b
    so the projected feature maps are never materialised (saves one read + one write of T*C*H*W per window);
  * the mask einsum (py:379) runs on the tcgen05 mask GEMM;
  * `with_masks=False` skips mask prediction altogether -- the offline meta-architecture deletes the tracker's
    masks right after the call (P/dvis_Plus/meta_architecture.py:1486).
"""
import random

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from .blocks import (MLP, FFNLayer, ReferringCrossAttentionLayer, SelfAttentionLayer, _cast_cached, _fast_path, add_norm,
                     linear)
from .precision import gemm_dtype
from .pixel_decoder import _c2_xavier_fill


class Noiser:
    """P/dvis_Plus/noiser.py: Hungarian matching of the current frame's queries to the previous frame's (the noise
    branches only fire in training)."""

    def __init__(self, noise_ratio=0.8, mode="wa"):
        assert mode in ["none", "rs", "wa", "cc"]
        self.mode = mode
        self.noise_ratio = noise_ratio

    def _rs_noise_forward(self, cur_embeds):
        indices = list(range(cur_embeds.shape[0]))
        np.random.shuffle(indices)
        return indices, cur_embeds[indices]

    def _wa_noise_forward(self, cur_embeds):
        indices = list(range(cur_embeds.shape[0]))
        np.random.shuffle(indices)
        noise_init = cur_embeds[indices]
        weight_ratio = torch.rand(cur_embeds.shape[0], 1, 1)
        noise_init = cur_embeds * weight_ratio.to(cur_embeds) + noise_init * (1.0 - weight_ratio.to(cur_embeds))
        ret = torch.arange(cur_embeds.shape[0], dtype=torch.int64).numpy()
        sel = (weight_ratio[:, 0, 0] < 0.5).to(torch.bool).numpy()
        ret[sel] = np.array(indices)[sel]
        return list(ret), noise_init

    def _cc_noise_forward(self, cur_embeds):
        C = cur_embeds.shape[-1]
        indices = torch.randint(0, C, (cur_embeds.shape[0],)).unsqueeze(-1).unsqueeze(-1)
        weight = torch.arange(C, dtype=torch.int64).unsqueeze(0).unsqueeze(0)
        weight = (weight < indices).to(torch.float32).to(cur_embeds)
        indices_, cur_embeds_ = self._rs_noise_forward(cur_embeds)
        ret_embeds = cur_embeds * weight + cur_embeds_ * (1 - weight)
        ret = torch.arange(cur_embeds.shape[0], dtype=torch.int64).numpy()
        sel = (indices[:, 0, 0] < C // 2).to(torch.bool).numpy()
        ret[sel] = np.array(indices_)[sel]
        return list(ret), ret_embeds

    def match_embds(self, ref_embds, cur_embds):
        """noiser.py:43-56: cosine cost, scipy.optimize.linear_sum_assignment on the host."""
        from scipy.optimize import linear_sum_assignment
        ref_embds, cur_embds = ref_embds.detach()[:, 0, :].float(), cur_embds.detach()[:, 0, :].float()
        ref_embds = ref_embds / (ref_embds.norm(dim=1)[:, None] + 1e-6)
        cur_embds = cur_embds / (cur_embds.norm(dim=1)[:, None] + 1e-6)
        C = 1 - torch.mm(cur_embds, ref_embds.transpose(0, 1))
        C = C.cpu()
        C = torch.where(torch.isnan(C), torch.full_like(C, 0), C)
        return linear_sum_assignment(C.transpose(0, 1))[1]

    def __call__(self, ref_embeds, cur_embeds, cur_embeds_no_norm=None, activate=False, cur_classes=None):
        if cur_embeds_no_norm is None:
            cur_embeds_no_norm = cur_embeds
        matched_indices = self.match_embds(ref_embeds, cur_embeds)
        if activate and random.random() < self.noise_ratio:
            if self.mode == "rs":
                return self._rs_noise_forward(cur_embeds_no_norm)
            if self.mode == "wa":
                return self._wa_noise_forward(cur_embeds_no_norm)
            if self.mode == "cc":
                return self._cc_noise_forward(cur_embeds_no_norm)
        return matched_indices, cur_embeds_no_norm[matched_indices]


class ReferringTracker_noiser(nn.Module):
    def __init__(self, hidden_channel=256, feedforward_channel=2048, num_head=8, decoder_layer_num=6, mask_dim=256,
                 class_num=25, noise_mode="hard", noise_ratio=0.5):
        super().__init__()
        self.num_heads = num_head
        self.num_layers = decoder_layer_num
        self.transformer_self_attention_layers = nn.ModuleList()
        self.transformer_cross_attention_layers = nn.ModuleList()
        self.transformer_ffn_layers = nn.ModuleList()
        for _ in range(self.num_layers):
            self.transformer_self_attention_layers.append(SelfAttentionLayer(hidden_channel, num_head, 0.0))
            self.transformer_cross_attention_layers.append(ReferringCrossAttentionLayer(hidden_channel, num_head, 0.0))
            self.transformer_ffn_layers.append(FFNLayer(hidden_channel, feedforward_channel, 0.0))
        self.use_memory = False
        self.decoder_norm = nn.LayerNorm(hidden_channel)
        self.class_embed = nn.Linear(2 * hidden_channel, class_num + 1)
        self.mask_embed = MLP(hidden_channel, hidden_channel, mask_dim, 3)
        self.ref_proj = MLP(hidden_channel, hidden_channel, hidden_channel, 3)
        for layer in self.ref_proj.layers:
            _c2_xavier_fill(layer)
        self.mask_feature_proj = nn.Conv2d(mask_dim, mask_dim, kernel_size=1, stride=1, padding=0)
        self.last_outputs = None
        self.last_frame_embeds = None
        self.last_reference = None
        # the reference passes noise_mode='hard' by default, which its Noiser asserts against; DVIS configs set it
        self.noiser = Noiser(noise_ratio=noise_ratio, mode=noise_mode if noise_mode != "hard" else "none")
        self.use_fast_path = True      # inference: batched matching + CUDA-graph frame steps (set False for the eager loop)
        self.use_cuda_graph = True
        self.match_on_host = False     # True: SciPy linear_sum_assignment on the host instead of the GPU LAP kernel
        # bf16 mode: the attention cores run on csrc/flash_attn.cu (dvis_flash_attn) instead of the library SDPA: 4.2 us vs
        # 6.5 us per call in the dependent chain at Q=200, 8 heads x 64 (profiles/r2_temporal_kernels.md)
        self.use_custom_attention = True
        # opt-in: every linear step on csrc/small_linear.cu as well (LayerNorm folded into the consumers' prologues: 36 launches
        # per frame, no library kernel in the chain).  Parity-green, but a mma.sync + cp.async GEMM step costs 4.9 us against
        # 3.3 us for cuBLAS's TMA / tcgen05 small-GEMM kernels, so the chain as a whole is slower (7.5 ms vs 4.3 ms per clip)
        self.use_fused_kernels = False
        self._fast = None

    def _clear_memory(self):
        self.last_outputs = None
        self.last_reference = None

    def _layer_stack(self, identity, tgt_fn, frame_key, memory):
        """6 x (referring cross-attn -> self-attn -> FFN); tgt_fn(j, ms_output) gives the query of layer j (py:236-329)."""
        ms_output = [identity]
        for j in range(self.num_layers):
            output = self.transformer_cross_attention_layers[j](ms_output[-1], tgt_fn(j, ms_output), frame_key, memory,
                                                                memory_mask=None, memory_key_padding_mask=None,
                                                                pos=None, query_pos=None)
            output = self.transformer_self_attention_layers[j](output, tgt_mask=None, tgt_key_padding_mask=None, query_pos=None)
            output = self.transformer_ffn_layers[j](output)
            ms_output.append(output)
        return ms_output

    def forward(self, frame_embeds, mask_features, resume=False, return_indices=False, frame_classes=None,
                frame_embeds_no_norm=None, with_masks=True):
        """frame_embeds (b, c, t, q); mask_features (b, t, c, h, w) (may be None when with_masks=False)."""
        if (not self.training) and _fast_path(frame_embeds) and frame_embeds.shape[0] == 1 and self.use_fast_path:
            return self._forward_fast(frame_embeds, mask_features, resume, return_indices, frame_embeds_no_norm, with_masks)
        frame_embeds = frame_embeds.permute(2, 3, 0, 1).float()                  # t, q, b, c
        if frame_embeds_no_norm is not None:
            frame_embeds_no_norm = frame_embeds_no_norm.permute(2, 3, 0, 1).float()
        n_frame = frame_embeds.shape[0]
        outputs, ret_indices, all_refs = [], [], []
        for i in range(n_frame):
            cur = frame_embeds[i]
            cur_nn = frame_embeds_no_norm[i] if frame_embeds_no_norm is not None else cur
            cur_cls = None if frame_classes is None else frame_classes[i]
            frame_key = cur_nn
            if i == 0 and resume is False:
                self._clear_memory()
                indices, noised_init = self.noiser(cur, cur, cur_embeds_no_norm=cur_nn, activate=False, cur_classes=cur_cls)
                self.last_frame_embeds = cur[indices]
                ret_indices.append(indices)
                ms_output = self._layer_stack(
                    noised_init, lambda j, ms: self.ref_proj(frame_key if j == 0 else ms[-1]).float(), frame_key, cur_nn)
                ms_output[0] = cur_nn[indices]
                self.last_reference = self.ref_proj(frame_key).float()
            else:
                reference = self.ref_proj(self.last_outputs[-1]).float()
                self.last_reference = reference
                indices, noised_init = self.noiser(self.last_frame_embeds, cur, cur_embeds_no_norm=cur_nn,
                                                   activate=self.training, cur_classes=cur_cls)
                self.last_frame_embeds = cur[indices]
                ret_indices.append(indices)
                ms_output = self._layer_stack(noised_init, lambda j, ms: reference, frame_key, cur_nn)
                ms_output[0] = cur_nn[indices]
            all_refs.append(self.last_reference)
            ms_output = torch.stack([m.float() for m in ms_output], dim=0)      # (1 + layers, q, b, c)
            self.last_outputs = ms_output
            outputs.append(ms_output[1:])
        outputs = torch.stack(outputs, dim=0)                                    # (t, l, q, b, c)
        all_refs = torch.stack(all_refs, dim=0)                                  # (t, q, b, c)
        if not self.training:
            outputs = outputs[:, -1:]
        outputs_class, outputs_masks = self.prediction(outputs, mask_features, all_refs, with_masks=with_masks)
        out = {
            "pred_logits": outputs_class[-1].transpose(1, 2),                    # (b, t, q, c)
            "pred_masks": None if outputs_masks is None else outputs_masks[-1],   # (b, q, t, h, w)
            "aux_outputs": self._set_aux_loss(outputs_class, outputs_masks),
            "pred_embds": outputs[:, -1].permute(2, 3, 0, 1),                    # (b, c, t, q)
            "pred_references": all_refs.permute(2, 3, 0, 1),
        }
        return (out, ret_indices) if return_indices else out

    @torch.jit.unused
    def _set_aux_loss(self, outputs_class, outputs_seg_masks):
        if outputs_seg_masks is None:
            return [{"pred_logits": a.transpose(1, 2)} for a in outputs_class[:-1]]
        return [{"pred_logits": a.transpose(1, 2), "pred_masks": b} for a, b in zip(outputs_class[:-1], outputs_seg_masks[:-1])]

    def prediction(self, outputs, mask_features, references, with_masks=True):
        """py:368-380.  outputs (t,l,q,b,c); mask_features (b,t,c,h,w) *un-projected*; references (t,q,b,c)."""
        decoder_output = self.decoder_norm(outputs.float()).permute(1, 3, 0, 2, 4)          # (l, b, t, q, c)
        references = references.unsqueeze(1).repeat(1, decoder_output.size(0), 1, 1, 1).permute(1, 3, 0, 2, 4)
        outputs_class = linear(self.class_embed, torch.cat([references, decoder_output], dim=-1)).float().transpose(2, 3)
        if not with_masks:
            return outputs_class, None
        mask_embed = self.mask_embed(decoder_output).float()                               # (l, b, t, q, c)
        W = self.mask_feature_proj.weight.flatten(1).float()                               # (c_out, c_in)
        bias = self.mask_feature_proj.bias.float()
        if _fast_path(mask_features):
            # fold the 1x1 conv into the embeddings: me' = me @ W  (c_in), offset = me . b
            l, b, t, q, c = mask_embed.shape
            me_f = torch.matmul(mask_embed, W)
            offs = torch.matmul(mask_embed, bias)                                          # (l, b, t, q)
            masks = []
            for li in range(l):
                feats = mask_features.flatten(0, 1)                                        # (b*t, c, h, w)
                m = ops.mask_logits(me_f[li].flatten(0, 1), feats, torch.float32)          # (b*t, q, h, w)
                m = m.reshape(b, t, q, *m.shape[-2:]) + offs[li][..., None, None]
                masks.append(m.permute(0, 2, 1, 3, 4))
            return outputs_class, torch.stack(masks, 0)
        shape = mask_features.shape
        mf = F.conv2d(mask_features.flatten(0, 1).float(), self.mask_feature_proj.weight, self.mask_feature_proj.bias).reshape(shape)
        return outputs_class, torch.einsum("lbtqc,btchw->lbqthw", mask_embed, mf)


    # ================================================================================================
    # Inference fast path.  Same arithmetic as the loop in forward() (py:210-340), reorganised around what
    # actually depends on what:
    #   * the Hungarian chain depends only on the SEGMENTER's frame embeddings, never on tracker outputs
    #     (noiser.py:43-56 gets last_frame_embeds = cur[indices], py:224,285): all T cost matrices come from one batched
    #     GEMM and one device->host copy; each frame's assignment is solved against the un-permuted previous frame and
    #     composed with the previous permutation (idx_t = sigma_t[idx_{t-1}]) -- T host syncs become 1;
    #   * keys / values of the referring cross-attention depend only on the frame's own queries: projected for all
    #     frames and all layers in one GEMM before the sequential part;
    #   * for frames after the first the cross-attention query of every layer is the same `reference`
    #     (py:278,293,313): the 6 layers' cross-attention cores run as one batched attention;
    #   * the remaining strictly sequential per-frame chain is captured once in a CUDA graph and replayed per frame.
    # ================================================================================================
    def _stacked(self, dt):
        L = self.num_layers
        ca, sa, ff = self.transformer_cross_attention_layers, self.transformer_self_attention_layers, self.transformer_ffn_layers
        params = [p for p in self.parameters()]
        key = (dt, tuple(p._version for p in params), params[0].data_ptr())
        if self._fast is None or self._fast["key"] != key:
            C = self.decoder_norm.normalized_shape[0]
            d = lambda t: t.detach().to(dt).contiguous()
            f = dict(key=key, C=C)
            f["wq"] = d(torch.cat([ca[j].multihead_attn.in_proj_weight[:C] for j in range(L)], 0))          # (L*C, C)
            f["bq"] = d(torch.cat([ca[j].multihead_attn.in_proj_bias[:C] for j in range(L)], 0))
            f["wkv"] = d(torch.cat([ca[j].multihead_attn.in_proj_weight[C:] for j in range(L)], 0))         # (L*2C, C)
            f["bkv"] = d(torch.cat([ca[j].multihead_attn.in_proj_bias[C:] for j in range(L)], 0))
            f["wo"] = d(torch.stack([ca[j].multihead_attn.out_proj.weight.t() for j in range(L)], 0))       # (L, C, C) = W^T
            f["bo"] = d(torch.stack([ca[j].multihead_attn.out_proj.bias for j in range(L)], 0))[:, None, :]
            if dt == torch.bfloat16:
                # operands of the fused temporal-stage kernels (ops.linear_small / ops.flash_attn): bf16 weights in
                # nn.Linear layout, fp32 biases and LayerNorm parameters
                w = lambda t: t.detach().to(torch.bfloat16).contiguous()
                v = lambda t: t.detach().float().contiguous()
                ln = lambda n: (v(n.weight), v(n.bias))
                f["k_ref"] = [(w(l.weight), v(l.bias)) for l in self.ref_proj.layers]
                f["k_wq"], f["k_bq"] = f["wq"], v(f["bq"])
                f["k_wo"] = w(torch.stack([ca[j].multihead_attn.out_proj.weight for j in range(L)], 0))       # (L, C_out, C_in)
                f["k_bo"] = v(torch.stack([ca[j].multihead_attn.out_proj.bias for j in range(L)], 0))
                f["k_layers"] = [dict(ln_ca=ln(ca[j].norm), ln_sa=ln(sa[j].norm), ln_ff=ln(ff[j].norm),
                                      w_qkv=w(sa[j].self_attn.in_proj_weight), b_qkv=v(sa[j].self_attn.in_proj_bias),
                                      w_o=w(sa[j].self_attn.out_proj.weight), b_o=v(sa[j].self_attn.out_proj.bias),
                                      w_1=w(ff[j].linear1.weight), b_1=v(ff[j].linear1.bias),
                                      w_2=w(ff[j].linear2.weight), b_2=v(ff[j].linear2.bias)) for j in range(L)]
            self._fast = f
            self._graphs = {}
        return self._fast

    def _match_all(self, cur, ref0):
        """cur (T, Q, C) normalised-query embeddings of the window; ref0 (Q, C) reference of the first frame (the frame
        itself at the start of a video, last_frame_embeds when resuming).  -> (T, Q) int64 DEVICE tensor of indices.
        All T assignment problems are solved concurrently on the GPU (dvis_lap_chain): no host sync at all."""
        unit = lambda z: z / (z.norm(dim=-1, keepdim=True) + 1e-6)
        n = unit(cur.float())
        prev = torch.cat([unit(ref0.float())[None], n[:-1]], 0)
        cost = 1 - torch.bmm(prev, n.transpose(1, 2))                                # (T, Q_ref, Q_cur) == C.T of noiser.py:49-54
        if self.match_on_host:                                                      # SciPy on the host (debug / cross-check)
            from scipy.optimize import linear_sum_assignment
            c = cost.cpu()
            c = torch.where(torch.isnan(c), torch.zeros_like(c), c).numpy()
            idx, out = None, []
            for t in range(c.shape[0]):
                sigma = linear_sum_assignment(c[t])[1]
                idx = sigma if idx is None else sigma[idx]
                out.append(idx)
            return torch.as_tensor(np.stack(out), device=cur.device, dtype=torch.long)
        return ops.lap_chain(cost)[1]

    def _frame_body(self, f, ref_src, identity, kv, first):
        """One frame of py:236-329 on (Q, C) tensors.  ref_src: last_outputs[-1] of the previous frame (or the frame key
        for the first frame); identity: cur_no_norm[indices]; kv: (Q, L, 2, H, dh) projected keys / values.
        Returns the stacked layer outputs (L, Q, C) fp32 and this frame's reference (Q, C) fp32 (py:276,279)."""
        L, C, H = self.num_layers, f["C"], self.num_heads
        dh = C // H
        Q = identity.shape[0]
        dt = f["wq"].dtype
        sa, ff = self.transformer_self_attention_layers, self.transformer_ffn_layers
        ca = self.transformer_cross_attention_layers
        scale = 1.0 / (dh ** 0.5)

        def ln(norm, x, res32):
            """LayerNorm(x + res) -> (fp32 stream, GEMM-dtype copy) in one kernel."""
            y32, ylp, _ = ops.add_layernorm(x.contiguous(), res32, norm.weight, norm.bias, norm.eps, lp_dtype=dt)
            return y32, ylp

        def lin(layer, x_lp, relu=False):
            w, b = _cast_cached(layer, dt) if layer.weight.dtype != dt else (layer.weight, layer.bias)
            if relu:
                return torch._addmm_activation(b, x_lp, w.t())          # bias + ReLU in the GEMM epilogue
            return torch.addmm(b, x_lp, w.t())

        own_attn = self.use_custom_attention and dt == torch.bfloat16 and dh in (32, 64)

        def attend(q, k, v):
            """q (B, Lq, H, dh), k / v (B, Lk, H, dh) views with packed heads -> (B, Lq, C)."""
            if own_attn:
                return ops.flash_attn(q, k, v, scale)
            o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), scale=scale)
            return o.transpose(1, 2).reshape(q.shape[0], q.shape[1], C)

        def ref_mlp(x_lp):
            n = self.ref_proj.num_layers
            for i, layer in enumerate(self.ref_proj.layers):
                x_lp = lin(layer, x_lp, relu=i < n - 1)
            return x_lp

        outs = []
        x32 = identity
        reference = ref_mlp(ref_src.to(dt))                                                           # (Q, C)
        kvl = kv.permute(1, 0, 2, 3, 4)                                                               # (L, Q, 2, H, dh) view
        if not first:
            q_all = F.linear(reference, f["wq"], f["bq"]).view(Q, L, H, dh).permute(1, 0, 2, 3)      # (L, Q, H, dh) view
            o_all = torch.baddbmm(f["bo"], attend(q_all, kvl[:, :, 0], kvl[:, :, 1]), f["wo"])        # (L, Q, C)
        x_lp = None
        for j in range(L):
            if first:
                tgt = reference if j == 0 else ref_mlp(x_lp)
                q = F.linear(tgt, f["wq"][j * C:(j + 1) * C], f["bq"][j * C:(j + 1) * C]).view(1, Q, H, dh)
                o = torch.addmm(f["bo"][j], attend(q, kvl[j:j + 1, :, 0], kvl[j:j + 1, :, 1])[0], f["wo"][j])
            else:
                o = o_all[j]
            x32, x_lp = ln(ca[j].norm, o, x32)
            m = sa[j].self_attn
            w, b = m._weights(dt)
            qkv = torch.addmm(b, x_lp, w.t()).view(1, Q, 3, H, dh)
            o = attend(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2])[0]                                   # (Q, C)
            x32, x_lp = ln(sa[j].norm, lin(m.out_proj, o), x32)
            x32, x_lp = ln(ff[j].norm, lin(ff[j].linear2, lin(ff[j].linear1, x_lp, relu=True)), x32)
            outs.append(x32)
        return torch.stack(outs, 0), reference.float()

    def _fused_ok(self, f):
        """The hand-written temporal-stage kernels cover bf16 mode, post-norm layers, hidden sizes 256 / 384 / 512 and
        head dims 32 / 64 (every DVIS++ / DVIS-DAQ config).  fp32 mode keeps the library GEMMs (CUDA-core fp32)."""
        C = f["C"]
        return self.use_fused_kernels and "k_layers" in f and C % 128 == 0 and C <= 512 and (C // self.num_heads) in (32, 64)

    def _frame_body_fused(self, f, prev, prev_is_pre, identity, kv, first):
        """One frame of py:236-329 on libdvis_b200 kernels only: 37 launches (3 + 1 + 1 + 1 + 1 + 6 x 5) for a frame after the
        first, 67 for the first frame of a video (its referring query is re-derived per layer, py:244-251).  Every linear step
        is a plain bf16 GEMM (csrc/small_linear.cu); the post-norm blocks' LayerNorms run in the PRODUCERS' epilogues
        (dvis_linear_small_ln: the last CTA of a 32-row block normalises it), so no LayerNorm is a dependent step of its own
        except the one that has no producer GEMM (layer 0's cross-attention residual).
        prev (Q, C) fp32: the previous frame's NORMALISED last-layer output (or, for the first frame, the frame key);
        identity (Q, C) fp32; kv (Q, L, 2, H, dh) bf16.  `prev_is_pre` is kept for the caller's bookkeeping and must be False.
        -> (out (Q, C) fp32: this frame's normalised last-layer output, None, reference (Q, C) fp32, inner: the L-1 inner
        layer outputs, fp32)."""
        assert not prev_is_pre
        L, C, H = self.num_layers, f["C"], self.num_heads
        dh = C // H
        Q = identity.shape[0]
        scale = 1.0 / (dh ** 0.5)
        lay = f["k_layers"]
        eps = self.transformer_ffn_layers[0].norm.eps
        (w1, b1), (w2, b2), (w3, b3) = f["k_ref"]
        kvl = kv.permute(1, 0, 2, 3, 4)                                                              # (L, Q, 2, H, dh) view

        def ref_mlp(x16):
            """ref_proj(x) -> (fp32, bf16)"""
            _, h, _, _ = ops.linear_small(w1, b1, x=x16, relu=True)
            _, h, _, _ = ops.linear_small(w2, b2, x=h, relu=True)
            r32, r16, _, _ = ops.linear_small(w3, b3, x=h, out_f32=True)
            return r32, r16

        x16 = prev.to(torch.bfloat16)
        if not first:
            # reference = ref_proj(last_outputs[-1]) (py:278); the 6 layers' referring cross-attention shares it (py:293,313)
            ref32, ref16 = ref_mlp(x16)
            _, q_all, _, _ = ops.linear_small(f["k_wq"], f["k_bq"], x=ref16)                          # (Q, L*C)
            o = ops.flash_attn(q_all.view(Q, L, H, dh).permute(1, 0, 2, 3), kvl[:, :, 0], kvl[:, :, 1], scale)   # (L, Q, C)
            _, o_all, _, _ = ops.linear_small(f["k_wo"], f["k_bo"], x=o)                              # (L, Q, C) bf16
        inner = []
        x32 = None                                                                                   # output of the previous layer
        for j in range(L):
            p = lay[j]
            if first:
                # tgt = ref_proj(frame key) for layer 0, ref_proj(previous layer's output) afterwards (py:244-251)
                r32, r16 = ref_mlp(x16)
                if j == 0:
                    ref32 = r32
                _, q, _, _ = ops.linear_small(f["k_wq"][j * C:(j + 1) * C], f["k_bq"][j * C:(j + 1) * C], x=r16)
                o = ops.flash_attn(q.view(1, Q, H, dh), kvl[j:j + 1, :, 0], kvl[j:j + 1, :, 1], scale)[0]
                _, o_j, _, _ = ops.linear_small(f["k_wo"][j], f["k_bo"][j], x=o)
                # x1 = LN_ca(x + o_j) (tracker.py:47-48), x = identity or the previous layer's output
                x1_32, x1_16, _ = ops.add_layernorm(o_j, identity if j == 0 else x32, *p["ln_ca"], eps, lp_dtype=torch.bfloat16)
            elif j == 0:
                x1_32, x1_16, _ = ops.add_layernorm(o_all[0], identity, *p["ln_ca"], eps, lp_dtype=torch.bfloat16)
            # self-attention block: QKV, attention, out-proj + residual + LN_sa (epilogue)
            qkv = ops.linear_small(p["w_qkv"], p["b_qkv"], x=x1_16)[1].view(1, Q, 3, H, dh)
            o = ops.flash_attn(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], scale)[0]
            x2_32, x2_16, _, _ = ops.linear_small_ln(p["w_o"], p["b_o"], o, x1_32, p["ln_sa"], eps=eps)
            # FFN block: FFN1 + ReLU, FFN2 (split-K) + residual + LN_ffn (epilogue) [+ the next layer's LN_ca(x + o_all[j+1])]
            hid = ops.linear_small(p["w_1"], p["b_1"], x=x2_16, relu=True)[1]
            nxt = (not first) and j + 1 < L
            x32, x16, x1_32, x1_16 = ops.linear_small_ln(p["w_2"], p["b_2"], hid, x2_32, p["ln_ff"], eps=eps,
                                                          src1=o_all[j + 1] if nxt else None, ln2=lay[j + 1]["ln_ca"] if nxt else None)
            if j + 1 < L:
                inner.append(x32)
        return x32, None, ref32, inner

    def _graph_step(self, f, first, Q, dev, prev_is_pre=False):
        """Capture one frame step (first / later frame; fused or library body) once per configuration; returns
        (graph, static inputs (prev, identity, kv), static outputs).  The static buffers are shared by every replay of a
        configuration: replays must stay on ONE stream (the tracker's callers do)."""
        fused = self._fused_ok(f)
        key = (first, prev_is_pre, fused, self.use_custom_attention, Q, str(dev), gemm_dtype())
        g = self._graphs.get(key)
        if g is None:
            C, L, H = f["C"], self.num_layers, self.num_heads
            s_ref = torch.zeros(Q, C, device=dev)
            s_id = torch.zeros(Q, C, device=dev)
            s_kv = torch.zeros(Q, L, 2, H, C // H, device=dev, dtype=gemm_dtype())
            if fused:
                body = lambda: self._frame_body_fused(f, s_ref, prev_is_pre, s_id, s_kv, first)
            else:
                body = lambda: self._frame_body(f, s_ref, s_id, s_kv, first)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    body()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                s_out = body()
            g = (graph, s_ref, s_id, s_kv, s_out)
            self._graphs[key] = g
        return g

    def _run_frame(self, f, prev, prev_is_pre, identity, kv, first, fused):
        """One frame step, through its CUDA graph unless an outer capture is running; returns fresh tensors."""
        if self.use_cuda_graph and not torch.cuda.is_current_stream_capturing():
            graph, s_ref, s_id, s_kv, s_out = self._graph_step(f, first, identity.shape[0], identity.device, prev_is_pre)
            s_ref.copy_(prev)
            s_id.copy_(identity)
            s_kv.copy_(kv)
            graph.replay()
            clone = lambda o: None if o is None else [x.clone() for x in o] if isinstance(o, list) else o.clone()
            return tuple(clone(o) for o in s_out)
        if fused:
            return self._frame_body_fused(f, prev, prev_is_pre, identity, kv, first)
        return self._frame_body(f, prev, identity, kv, first)

    def _forward_fast(self, frame_embeds, mask_features, resume, return_indices, frame_embeds_no_norm, with_masks):
        dt = gemm_dtype()
        f = self._stacked(dt)
        L, C, H = self.num_layers, f["C"], self.num_heads
        cur = frame_embeds[0].permute(1, 2, 0).float().contiguous()                       # (T, Q, C)
        cur_nn = cur if frame_embeds_no_norm is None else frame_embeds_no_norm[0].permute(1, 2, 0).float().contiguous()
        T, Q, _ = cur.shape
        dev = cur.device
        start_of_video = not resume
        if start_of_video:
            self._clear_memory()
        ref0 = cur[0] if start_of_video else self.last_frame_embeds[:, 0, :]
        idx_dev = self._match_all(cur, ref0)                                               # (T, Q) on the device
        init = torch.gather(cur_nn, 1, idx_dev[..., None].expand(-1, -1, C))               # cur_nn[t][idx_t]
        self.last_frame_embeds = torch.gather(cur[-1], 0, idx_dev[-1][:, None].expand(-1, C))[:, None, :]
        refs = []
        prev_last = None if start_of_video else self.last_outputs[-1][:, 0, :]
        if self._fused_ok(f):
            # keys / values of all frames and layers: one launch (T*Q rows x L*2C columns)
            kv = ops.linear_small(f["wkv"], f["bkv"].float(), x=cur_nn.to(dt).view(T * Q, C))[1].view(T, Q, L, 2, H, C // H)
            lasts, prev = [], prev_last                                                     # normalised last-layer outputs
            for t in range(T):
                first = prev is None
                prev, _, reference, inner = self._run_frame(f, cur_nn[t] if first else prev, False, init[t], kv[t], first, True)
                lasts.append(prev)
                refs.append(reference)
            last_stack = torch.stack([init[T - 1]] + inner + [lasts[-1]], 0)
            outputs = torch.stack(lasts, 0)[:, None, :, None, :]                            # (t, 1, q, b, c)  eval: last layer
        else:
            # keys / values of all frames and layers: one GEMM
            kv = F.linear(cur_nn.to(dt), f["wkv"], f["bkv"]).view(T, Q, L, 2, H, C // H)       # per frame: (Q, L, 2, H, dh)
            outs = []
            for t in range(T):
                first = prev_last is None
                layer_out, reference = self._run_frame(f, cur_nn[t] if first else prev_last, False, init[t], kv[t], first, False)
                refs.append(reference)
                prev_last = layer_out[-1]
                outs.append(layer_out)
            last_stack = torch.cat([init[T - 1][None], outs[-1]], 0)                        # only the window's last frame is kept
            outputs = torch.stack([o[-1:] for o in outs], 0)[:, :, :, None, :]              # (t, 1, q, b, c)  eval: last layer
        self.last_outputs = last_stack[:, :, None, :]                                       # (1+L, q, b, c)
        self.last_reference = refs[-1][:, None, :]
        all_refs = torch.stack(refs, 0)[:, :, None, :]                                      # (t, q, b, c)
        outputs_class, outputs_masks = self.prediction(outputs, mask_features, all_refs, with_masks=with_masks)
        out = {
            "pred_logits": outputs_class[-1].transpose(1, 2),
            "pred_masks": None if outputs_masks is None else outputs_masks[-1],
            "aux_outputs": self._set_aux_loss(outputs_class, outputs_masks),
            "pred_embds": outputs[:, -1].permute(2, 3, 0, 1),
            "pred_references": all_refs.permute(2, 3, 0, 1),
        }
        return (out, [i for i in idx_dev.cpu().numpy()]) if return_indices else out
