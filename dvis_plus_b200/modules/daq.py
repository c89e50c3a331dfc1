"""DVIS-DAQ drop-ins: SlotCrossAttentionLayer (D/dvis_daq/slot_attention.py:89-173),
VideoInstanceCutter's inference path (D/dvis_daq/track_module.py:102-260,606-791) and the DAQ TemporalRefiner
(D/dvis_daq/refiner.py:6-246).  Same constructor keywords, parameter names and recurrent state as the reference.

The tracker works on a RUN-TIME number of queries (track queries + `num_new_ins` new-instance queries, track_module.py:640);
every kernel behind it takes Q at run time.  Per frame:
  * the mask einsum "lbqc,bchw->lbqhw" (:767) runs on the tcgen05 mask GEMM with the 1x1 `mask_feature_proj` (:609-611)
    folded into the query side, and only for the last layer (the reference computes all L+1 layers and uses [-1]);
  * the mask-pooled positional embeddings (`get_mask_pos_embed`, :771-791) are one (Q x HW) @ (HW x C) GEMM on the
    thresholded logits instead of a broadcast (b, q, c, h, w) product in 50-query chunks;
  * the per-instance bookkeeping (:700-747) stays on the host, as in the reference, with one device->host copy of the
    validity flags per frame.
Training (`forward` with targets / matcher, :319-604) is orchestration outside the hot path and is not mirrored.
"""
import random

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from .blocks import (MLP, CrossAttentionLayer, FFNLayer, MultiheadAttention, SelfAttentionLayer, _fast_path, _with_pos,
                     _xavier_reset, linear)
from .refiner import TemporalRefiner as _PlusTemporalRefiner


class SlotAttention(nn.Module):
    """One cross-attention step whose softmax runs over the SLOTS (slot_attention.py:6-66)."""

    def __init__(self, in_features, num_iterations, num_slots, slot_size, mlp_hidden_size, eps=1e-6):
        super().__init__()
        self.in_features, self.num_iterations, self.num_slots = in_features, num_iterations, num_slots
        self.slot_size, self.mlp_hidden_size, self.eps = slot_size, mlp_hidden_size, eps
        self.attn_scale = slot_size ** -0.5
        self.norm_inputs = nn.LayerNorm(in_features)
        self.project_q = nn.Sequential(nn.LayerNorm(slot_size), nn.Linear(slot_size, slot_size, bias=False))
        self.project_k = nn.Linear(in_features, slot_size, bias=False)

    def forward(self, inputs, inputs_k, slots):
        """inputs, inputs_k (B, N, C); slots (B, M, C) -> updates (M, B, C)."""
        k = linear(self.project_k, self.norm_inputs(inputs_k.float())).float()
        q = linear(self.project_q[1], self.project_q[0](slots.float())).float()
        attn = F.softmax(self.attn_scale * torch.einsum("bnc,bmc->bnm", k, q), dim=-1) + self.eps
        attn = attn / attn.sum(dim=1, keepdim=True)
        return torch.einsum("bnm,bnc->bmc", attn, inputs.float()).transpose(0, 1)


class SlotCrossAttentionLayer(nn.Module):
    def __init__(self, d_model, nhead, dropout=0.0, activation="relu", normalize_before=False):
        super().__init__()
        self.multihead_attn = MultiheadAttention(d_model, nhead, dropout=dropout)
        self.norm = nn.LayerNorm(d_model)
        self.normalize_before = normalize_before
        self.slot_attn = SlotAttention(in_features=d_model, num_iterations=1, num_slots=0, slot_size=d_model,
                                       mlp_hidden_size=d_model, eps=1e-6)
        _xavier_reset(self)

    def forward(self, tgt, memory, memory_mask=None, memory_key_padding_mask=None, pos=None, query_pos=None, slot_query=None):
        assert memory_key_padding_mask is None and not self.normalize_before
        if slot_query is None:
            slot_query = tgt
        tgt2 = self.multihead_attn(_with_pos(tgt, query_pos), _with_pos(memory, pos), memory, attn_mask=memory_mask).float()
        tgt3 = self.slot_attn(tgt2.transpose(0, 1), (tgt + tgt2).transpose(0, 1), slot_query.transpose(0, 1))
        return self.norm(tgt.float() + tgt3)


class VideoInstanceSequence:
    """Per-instance track record (track_module.py:16-100), inference-time fields."""

    def __init__(self, start_time, matched_gt_id=-1, maximum_chache=10):
        self.sT, self.eT, self.maximum_chache = start_time, -1, maximum_chache
        self.dead, self.gt_id, self.invalid_frames = False, matched_gt_id, 0
        self.embeds, self.pred_logits, self.pred_masks, self.appearance = [], [], [], []
        self.pos_embeds, self.similarity_guided_pos_embed, self.similarity_guided_pos_embed_list = [], None, []

    def update_pos(self, pos_embed):
        """Similarity-guided fusion of the positional embedding (track_module.py:71-100)."""
        self.pos_embeds.append(pos_embed)
        if not self.similarity_guided_pos_embed_list:
            self.similarity_guided_pos_embed = pos_embed
        else:
            sidx = max(0, len(self.pos_embeds) - self.maximum_chache)
            hist = torch.stack(self.pos_embeds[sidx:-1], dim=0)
            sim = torch.sum(torch.einsum("bc,c->b", F.normalize(hist, dim=-1), F.normalize(pos_embed.squeeze(), dim=-1))) / hist.shape[0]
            beta = max(0, sim)
            self.similarity_guided_pos_embed = (1 - beta) * self.similarity_guided_pos_embed + beta * pos_embed
        self.similarity_guided_pos_embed_list.append(self.similarity_guided_pos_embed)


class VideoInstanceCutter(nn.Module):
    def __init__(self, hidden_dim=256, feedforward_dim=2048, num_head=8, decoder_layer_num=6, mask_dim=256, num_classes=25,
                 num_new_ins=100, training_select_threshold=0.1, inference_select_threshold=0.1, kick_out_frame_num=8,
                 mask_nms_thr=0.6, match_score_thr=0.3, num_slots=5, keep_threshold=0.01, task="vis", ovis_infer=True):
        super().__init__()
        self.num_heads, self.hidden_dim, self.num_layers = num_head, hidden_dim, decoder_layer_num
        self.num_classes, self.task, self.ovis_infer = num_classes, task, ovis_infer
        self.transformer_self_attention_layers = nn.ModuleList()
        self.transformer_cross_attention_layers = nn.ModuleList()
        self.transformer_ffn_layers = nn.ModuleList()
        self.slot_cross_attention_layers = nn.ModuleList()
        self.slot_ffn_layers = nn.ModuleList()
        for _ in range(self.num_layers):
            self.transformer_self_attention_layers.append(SelfAttentionLayer(hidden_dim, num_head, 0.0))
            self.transformer_cross_attention_layers.append(CrossAttentionLayer(hidden_dim, num_head, 0.0))
            self.transformer_ffn_layers.append(FFNLayer(hidden_dim, feedforward_dim, 0.0))
            self.slot_cross_attention_layers.append(SlotCrossAttentionLayer(hidden_dim, num_head, 0.0))
            self.slot_ffn_layers.append(FFNLayer(hidden_dim, feedforward_dim, 0.0))
        self.decoder_norm = nn.LayerNorm(hidden_dim)
        self.class_embed = nn.Linear(hidden_dim, num_classes + 1)
        self.pos_embed = MLP(mask_dim, hidden_dim, hidden_dim, 3)
        self.mask_embed = MLP(hidden_dim, hidden_dim, mask_dim, 3)
        self.mask_feature_proj = nn.Conv2d(mask_dim, mask_dim, kernel_size=1, stride=1, padding=0)
        self.new_ins_embeds = nn.Embedding(1, hidden_dim)
        self.bg_slots = nn.Embedding(num_slots, hidden_dim)
        self.num_new_ins, self.num_slots = num_new_ins, num_slots
        self.training_select_thr, self.inference_select_thr = training_select_threshold, inference_select_threshold
        self.kick_out_frame_num, self.mask_nms_thr = kick_out_frame_num, mask_nms_thr
        self.match_score_thr, self.keep_threshold = match_score_thr, keep_threshold
        self.memory_seq_ids = []
        # track x query assignment: SciPy on the host like the reference (default this round), or the GPU Hungarian kernel
        # (ops.lap_rect, no host copy of the cost matrix) with match_on_host = False
        self.match_on_host = True
        self._clear_memory()

    def _clear_memory(self):
        self.video_ins_hub = dict()
        self.last_seq_ids = None
        self.track_queries = None
        self.track_embeds = None

    def forward(self, *args, **kwargs):
        raise NotImplementedError("VideoInstanceCutter.forward is the training path (targets / matcher / losses, "
                                  "D/dvis_daq/track_module.py:319-604): orchestration outside the inference hot path")

    def readout(self, read_type="last"):
        """track_module.py:251-283."""
        assert read_type in ("last", "last_pos")
        dev = self.new_ins_embeds.weight.device
        out = []
        for seq_id in self.last_seq_ids:
            seq = self.video_ins_hub[seq_id]
            if read_type == "last":
                idx = -1
                while seq.embeds[idx] is None:
                    idx -= 1
                out.append(seq.embeds[idx])
            else:
                out.append(seq.similarity_guided_pos_embed)
        if out:
            return torch.stack(out, dim=0).unsqueeze(1)
        return torch.empty((0, 1, self.hidden_dim), dtype=torch.float32, device=dev)

    def match_with_embeds(self, trc_queries_feat, seg_queries_feat):
        """track_module.py:749-759: nearest segmenter query per track query, Hungarian-assigned where possible.  On the
        device the (tracks x queries) assignment runs in the GPU Hungarian kernel (ops.lap_rect): no `C.cpu()` sync."""
        t, s = trc_queries_feat.detach()[:, 0, :].float(), seg_queries_feat.detach()[:, 0, :].float()
        t = t / (t.norm(dim=1)[:, None] + 1e-6)
        s = s / (s.norm(dim=1)[:, None] + 1e-6)
        C = 1 - torch.mm(t, s.transpose(0, 1))
        least = torch.min(C, dim=1)[1]
        if _fast_path(C) and not self.match_on_host and 0 < C.shape[0] <= 1024 and 0 < C.shape[1] <= 1024:
            assigned = ops.lap_rect(C)
            return torch.where(assigned >= 0, assigned, least)
        from scipy.optimize import linear_sum_assignment
        rows, cols = linear_sum_assignment(C.cpu())
        least[torch.as_tensor(rows, device=least.device)] = torch.as_tensor(cols, dtype=torch.int64, device=least.device)
        return least

    def prediction(self, outputs, mask_features):
        """track_module.py:761-769 on un-projected `mask_features` (b, c, h, w): outputs (l, q, b, c) ->
        class (l, b, q, K+1), masks (l, b, q, h, w)."""
        dec = self.decoder_norm(outputs.float().transpose(1, 2))
        cls = linear(self.class_embed, dec).float()
        me = self.mask_embed(dec).float()                                            # (l, b, q, c)
        W = self.mask_feature_proj.weight.flatten(1).float()
        bias = self.mask_feature_proj.bias.float()
        if _fast_path(mask_features):
            me_f, offs = torch.matmul(me, W), torch.matmul(me, bias)                 # fold the 1x1 conv into the queries
            masks = torch.stack([ops.mask_logits(me_f[l], mask_features, torch.float32) + offs[l][..., None, None]
                                 for l in range(me.shape[0])], 0)
        else:
            mf = F.conv2d(mask_features.float(), self.mask_feature_proj.weight, self.mask_feature_proj.bias)
            masks = torch.einsum("lbqc,bchw->lbqhw", me, mf)
        return cls, masks

    def get_mask_pos_embed(self, mask, mask_features):
        """track_module.py:771-791: mask (b, q, h, w) logits, mask_features (b, c, h, w) -> (pos (q,b,c), obj (q,b,c))."""
        seg = mask.to(mask_features.device).sigmoid() > 0.5                           # (b, q, h, w)
        b, q = seg.shape[:2]
        if _fast_path(mask_features):
            feat = mask_features.permute(0, 2, 3, 1).reshape(b, -1, mask_features.shape[1])      # (b, hw, c) view if channels_last
            pooled = torch.bmm(seg.flatten(2).to(feat.dtype), feat).float()           # exact products, fp32 accumulation
        else:
            pooled = torch.einsum("bqp,bcp->bqc", seg.flatten(2).float(), mask_features.flatten(2).float())
        pooled = pooled / (seg.flatten(2).sum(-1, keepdim=True).float() + 1e-8)
        return self.pos_embed(pooled).float().transpose(0, 1), pooled.transpose(0, 1)

    def _stack(self, queries, memory, query_pos=None, pos=None):
        outs = [queries]
        x = queries
        for j in range(self.num_layers):
            x = self.transformer_cross_attention_layers[j](x, memory, query_pos=query_pos, pos=pos)
            x = self.transformer_self_attention_layers[j](x)
            x = self.transformer_ffn_layers[j](x)
            outs.append(x)
        return outs

    @torch.no_grad()
    def inference(self, frame_embeds, mask_features, frames_info, start_frame_id, resume=False, to_store="cpu"):
        """track_module.py:606-747.  frame_embeds (1, c, t, fq); mask_features (1, t, c, h, w) (un-projected);
        frames_info: {"seg_query_feat": nn.Embedding, "valid": [[bool (fq,)]...], "pred_masks": [[(fq, h, w)]...]}."""
        fe = frame_embeds.permute(2, 3, 0, 1).float()                                # t, q, b, c
        T, fQ, B, _ = fe.shape
        assert B == 1
        seg_query_feat = frames_info["seg_query_feat"].weight.unsqueeze(1).repeat(1, B, 1).float()
        new_ins = self.new_ins_embeds.weight.unsqueeze(1).repeat(self.num_new_ins, B, 1)
        bg = self.bg_slots.weight.unsqueeze(1).repeat(1, B, 1)
        for i in range(T):
            cur = fe[i]
            mf_i = mask_features[:, i]
            valid_fq = frames_info["valid"][i][0]
            first = i == 0 and resume is False
            slot_last = None
            if first:
                self._clear_memory()
                ms = self._stack(cur, cur)
            else:
                fq_pos, _ = self.get_mask_pos_embed(frames_info["pred_masks"][i][0][None], mf_i)
                queries = torch.cat([self.track_queries, new_ins])
                queries_pos = torch.cat([self.track_embeds, fq_pos])
                ms = self._stack(queries, cur, query_pos=queries_pos, pos=fq_pos)
                anchors = torch.cat([self.track_queries, bg])
                slots = seg_query_feat[self.match_with_embeds(anchors, seg_query_feat)]
                slots_query = torch.cat([self.track_embeds, bg], dim=0)
                for j in range(self.num_layers):
                    slots = self.slot_cross_attention_layers[j](slots, cur, query_pos=anchors, slot_query=slots_query)
                    slots = self.slot_ffn_layers[j](slots)
                slot_last = slots
            last = ms[-1].float()                                                    # (q', b, c)
            cls, masks = self.prediction(last[None], mf_i)                           # only the last layer is consumed
            cls, masks = cls[-1, 0], masks[-1]                                       # (q', K+1), (b, q', h, w)
            track_embeds, _ = self.get_mask_pos_embed(masks, mf_i)                   # (q', b, c)
            if first:
                valid = valid_fq
            else:
                num_tq = self.track_queries.shape[0]
                if self.ovis_infer:
                    slot_cls = linear(self.class_embed, self.decoder_norm(slot_last.float().transpose(0, 1))).float()[0]
                    trc = cls[:num_tq].softmax(-1)[:, :-1].max(dim=1)[0]
                    fg = slot_cls[:num_tq].softmax(-1)[:, :-1].max(dim=1)[0]
                    det = cls[-self.num_new_ins:].softmax(-1)[:, :-1].max(dim=1)[0]
                    valid = torch.cat([(trc > self.inference_select_thr) & (fg > self.keep_threshold),
                                       det > self.inference_select_thr], dim=0)
                else:
                    valid = cls.softmax(-1)[:, :-1].max(dim=1)[0] > self.inference_select_thr
            cur_seq_ids = []
            for k, ok in enumerate(valid.cpu().tolist()):                            # one host copy per frame
                if self.last_seq_ids is not None and k < len(self.last_seq_ids):
                    seq_id = self.last_seq_ids[k]
                else:
                    seq_id = random.randint(0, 100000)
                    while seq_id in self.video_ins_hub or seq_id in self.memory_seq_ids:
                        seq_id = random.randint(0, 100000)
                if ok:
                    if seq_id not in self.video_ins_hub:
                        self.video_ins_hub[seq_id] = VideoInstanceSequence(start_frame_id + i, seq_id)
                        self.memory_seq_ids.append(seq_id)
                    seq = self.video_ins_hub[seq_id]
                    seq.invalid_frames = 0
                    seq.appearance.append(True)
                    seq.update_pos(track_embeds[k, 0, :])
                elif self.last_seq_ids is not None and seq_id in self.last_seq_ids:
                    seq = self.video_ins_hub[seq_id]
                    seq.invalid_frames += 1
                    if seq.invalid_frames >= self.kick_out_frame_num:
                        seq.dead = True
                        continue
                    seq.appearance.append(False)
                else:
                    continue
                seq.embeds.append(last[k, 0, :])
                seq.pred_logits.append(cls[k, :])
                m = masks[0, k]
                seq.pred_masks.append(m.to(to_store).to(torch.float32) if to_store == "cpu" else m)
                cur_seq_ids.append(seq_id)
            self.last_seq_ids = cur_seq_ids
            self.track_queries = self.readout("last")
            self.track_embeds = self.readout("last_pos")


class TemporalRefiner(_PlusTemporalRefiner):
    """DAQ refiner (D/dvis_daq/refiner.py): the DVIS++ refiner with an optional short-conv branch, un-normalised
    `pred_embds`, and two extra (unused at inference) arguments."""

    def __init__(self, hidden_channel=256, feedforward_channel=2048, num_head=8, decoder_layer_num=6, mask_dim=256,
                 class_num=25, windows=5, use_local_attn=False):
        super().__init__(hidden_channel, feedforward_channel, num_head, decoder_layer_num, mask_dim, class_num, windows)
        self.use_local_attn = use_local_attn
        if not use_local_attn:                      # the reference only creates the conv branch when it is used (:43-55)
            self.conv_short_aggregate_layers = nn.ModuleList()
            self.conv_norms = nn.ModuleList()
        self.padding_embed = nn.Identity()

    def _short_conv(self, i, x):
        return super()._short_conv(i, x)

    def refine(self, instance_embeds, frame_embeds):
        if self.use_local_attn:
            return super().refine(instance_embeds, frame_embeds)
        n_batch, n_channel, n_frames, n_instance = instance_embeds.size()
        outputs, output = [], instance_embeds.float()
        frame_embeds = frame_embeds.float().permute(3, 0, 2, 1).flatten(1, 2)
        for i in range(self.num_layers):
            output = output.permute(2, 0, 3, 1).flatten(1, 2)                                   # (t, bq, c)
            output = self.transformer_time_self_attention_layers[i](output)
            output = output.reshape(n_frames, n_batch, n_instance, n_channel).permute(2, 1, 0, 3).flatten(1, 2)   # (q, bt, c)
            output = self.transformer_obj_self_attention_layers[i](output)
            output = self.transformer_cross_attention_layers[i](output, frame_embeds)
            output = self.transformer_ffn_layers[i](output)
            output = output.reshape(n_instance, n_batch, n_frames, n_channel).permute(1, 3, 2, 0)
            outputs.append(output)
        return torch.stack(outputs, dim=0).permute(3, 0, 4, 1, 2)

    def forward(self, instance_embeds, padding_mask, frame_embeds, mask_features, matched_gt_ids=None, with_masks=True):
        outputs = self.refine(instance_embeds, frame_embeds)
        outputs_class, outputs_masks = self.prediction(outputs, mask_features, with_masks=with_masks)
        return {
            "pred_logits": outputs_class[-1].transpose(1, 2),
            "pred_masks": None if outputs_masks is None else outputs_masks[-1],
            "aux_outputs": self._set_aux_loss(outputs_class, outputs_masks),
            "pred_embds": outputs[:, -1].permute(2, 3, 0, 1),            # not normalised in DAQ (refiner.py:147,155)
        }
