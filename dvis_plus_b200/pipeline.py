"""Offline DVIS++ clip inference over the hot path, frame-sharded across the GPUs of one NVSwitch box.

Mirrors what DVIS_Plus_offline.run_window_inference does between the backbone and post-processing
(P/dvis_Plus/meta_architecture.py:1446-1500): sem_seg_head (pixel decoder + predictor) per frame, tracker over the
frame queries, temporal refiner, final masks.  The reference processes windows sequentially on one GPU; here the T
frames of a clip are split into contiguous blocks of T/G frames, one block per rank:

    rank r:  pixel decoder + predictor on its frames   (all per-frame work, mask features stay local)
    all ranks: ONE all_gather of the packed per-frame query block [pred_embds | pred_embds_without_norm | pred_logits]
    all ranks: tracker + refiner layers on the gathered (T, Q, C) queries   (replicated; sequential in t / tiny)
    rank r:  final mask GEMM for ITS frames against ITS mask features

Nothing but the query block (0.84 MB / frame at Q=200) ever crosses NVLink.
"""
import torch
import torch.distributed as dist


class OfflineClipRunner:
    def __init__(self, pixel_decoder, predictor, tracker, refiner, group=None):
        self.pixel_decoder, self.predictor, self.tracker, self.refiner = pixel_decoder, predictor, tracker, refiner
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0

    # -- the exchange step -----------------------------------------------------------------------------
    @staticmethod
    def pack_queries(seg_out):
        """(1, C, t, Q) x2 + (1, t, Q, K) -> one contiguous (t, Q, 2C+K) fp32 block."""
        e = seg_out["pred_embds"][0].permute(1, 2, 0)                    # (t, Q, C)
        n = seg_out["pred_embds_without_norm"][0].permute(1, 2, 0)
        return torch.cat([e, n, seg_out["pred_logits"][0]], dim=-1).float().contiguous()

    @staticmethod
    def unpack_queries(block, C):
        e = block[..., :C].permute(2, 0, 1)[None]                         # (1, C, T, Q)
        n = block[..., C:2 * C].permute(2, 0, 1)[None]
        return e, n, block[..., 2 * C:][None]                             # logits (1, T, Q, K)

    def gather_queries(self, block):
        if self.world == 1:
            return block
        out = torch.empty((self.world * block.shape[0],) + tuple(block.shape[1:]), dtype=block.dtype, device=block.device)
        dist.all_gather_into_tensor(out, block, group=self.group)        # contiguous blocks: temporal order preserved
        return out

    # -- one clip ----------------------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(self, features):
        """features: backbone maps of THIS rank's frames, dict name -> (t_local, C_i, H_i, W_i).
        Returns pred_logits (1, T, Q, K+1) and pred_embds (1, C, T, Q) for the whole clip (identical on all ranks) and
        pred_masks (1, Q, t_local, H/4, W/4) for this rank's frames."""
        mask_features, _, multi_scale = self.pixel_decoder.forward_features(features)
        seg = self.predictor(multi_scale, mask_features)
        return self.temporal_stage(seg, mask_features)

    @torch.no_grad()
    def segment_stage(self, features):
        """Per-frame stage on this rank's frames: -> (packed query block (t_local, Q, 2C+K+1) fp32, mask_features)."""
        mask_features, _, multi_scale = self.pixel_decoder.forward_features(features)
        seg = self.predictor(multi_scale, mask_features)
        return self.pack_queries(seg), mask_features

    @torch.no_grad()
    def temporal_from_block(self, block, mask_features, C):
        """Tracker + refiner + final masks from the GATHERED query block (T, Q, 2C+K+1) and local mask features."""
        t_local = mask_features.shape[0]
        frame_embds, frame_embds_no_norm, _ = self.unpack_queries(block, C)
        track = self.tracker(frame_embds, None, resume=False, frame_embeds_no_norm=frame_embds_no_norm, with_masks=False)
        outputs = self.refiner.refine(track["pred_embds"], frame_embds_no_norm)
        dec = self.refiner.decoder_norm(outputs[:, -1:]).permute(1, 3, 0, 2, 4)
        logits = self.refiner.pred_class(dec)[-1].transpose(1, 2)
        t0 = self.rank * t_local
        masks = self.refiner.predict_masks(outputs[t0:t0 + t_local], mask_features[None])
        return {"pred_logits": logits, "pred_masks": masks, "pred_embds": dec[0].permute(0, 3, 1, 2),
                "online_pred_logits": track["pred_logits"]}

    # -- the temporal stage in two halves (RoundRobinClipRunner runs the first half on ONE rank per clip) -------------------
    @torch.no_grad()
    def temporal_payload(self, block, C):
        """Tracker + refiner on the gathered query block -> one contiguous fp32 (T, Q, Cm + 2(K+1) + C) tensor
        [mask embeddings | refined class logits | online (tracker) class logits | refined query embeddings]: everything a
        rank needs to finish the clip's masks for its own frames."""
        frame_embds, frame_embds_no_norm, _ = self.unpack_queries(block, C)
        track = self.tracker(frame_embds, None, resume=False, frame_embeds_no_norm=frame_embds_no_norm, with_masks=False)
        outputs = self.refiner.refine(track["pred_embds"], frame_embds_no_norm)           # (T, l, q, 1, c)
        dec = self.refiner.decoder_norm(outputs[:, -1:]).permute(1, 3, 0, 2, 4)           # (1, 1, T, q, c)
        logits = self.refiner.pred_class(dec)[-1].transpose(1, 2)                         # (1, T, q, K+1)
        emb = self.refiner.mask_embed(dec).float()                                        # (1, 1, T, q, Cm)
        return torch.cat([emb[0, 0], logits[0].float(), track["pred_logits"][0].float(), dec[0, 0].float()], dim=-1).contiguous()

    @torch.no_grad()
    def outputs_from_payload(self, payload, mask_features, C):
        """The second half: this rank's masks from the payload's mask embeddings and its local mask features; the clip-level
        tensors are sliced out of the payload.  Same dictionary as temporal_from_block."""
        t_local = mask_features.shape[0]
        K1 = self.refiner.class_embed.out_features
        Cm = payload.shape[-1] - 2 * K1 - C
        t0 = self.rank * t_local
        emb = payload[t0:t0 + t_local, :, :Cm]
        masks = self.refiner._masks(emb[None, None], mask_features[None])[-1]            # (1, q, t_local, h, w)
        return {"pred_logits": payload[None, :, :, Cm:Cm + K1], "pred_masks": masks,
                "pred_embds": payload[..., Cm + 2 * K1:].permute(2, 0, 1)[None],
                "online_pred_logits": payload[None, :, :, Cm + K1:Cm + 2 * K1]}

    @torch.no_grad()
    def vis_from_block(self, block, mask_features, C, post, img_size, output_size, first_resize_size=None, packed=False):
        """Tracker + refiner + the video-instance post-processing of DVIS_Plus_offline.forward's eval branch
        (P/dvis_Plus/meta_architecture.py:1377-1396 -> post_processing py:758-772 -> inference_video_vis py:818-868), fused
        around the final mask GEMM: the `max_num` instances are selected from the time-averaged class logits FIRST, so the
        GEMM produces max_num (not Q) masks per frame, and the resize chain + threshold runs straight from those stride-4
        logits (ops.vis_masks).  The reference computes all Q masks for all frames, moves them to the host in fp32, and
        up-samples the selected ones to the padded image size twice.

        block: the gathered (T, Q, 2C+K+1) query block; mask_features: THIS rank's (t_local, C_m, h, w) features;
        post: modules.postprocess.VideoPostProcessor (task "vis").  Nothing here synchronises with the host.
        -> dict of DEVICE tensors: pred_scores (n,), pred_labels (n,), pred_ids (n,) -- identical on all ranks -- and
        pred_masks (n, t_local, H_out, W_out) bool for this rank's frames (`packed`: one bit per pixel, uint8
        (n, t_local, H_out, ceil(W_out / 8)), see ops.unpack_masks)."""
        payload = self.vis_payload(block, C, post)
        return self.vis_from_payload(payload, mask_features, post.max_num, img_size, output_size, first_resize_size, packed)

    @torch.no_grad()
    def vis_payload(self, block, C, post):
        """First half of vis_from_block (tracker + refiner + instance selection): one flat fp32 tensor
        [mask embeddings of the n selected instances for every frame (T, n, Cm) | scores (n) | labels (n) | query ids (n)] --
        T*n*Cm + 3n floats (164 KB at T=16, n=10): what RoundRobinClipRunner broadcasts from the clip's owner rank."""
        frame_embds, frame_embds_no_norm, _ = self.unpack_queries(block, C)
        track = self.tracker(frame_embds, None, resume=False, frame_embeds_no_norm=frame_embds_no_norm, with_masks=False)
        outputs = self.refiner.refine(track["pred_embds"], frame_embds_no_norm)           # (T, l, q, 1, c)
        dec = self.refiner.decoder_norm(outputs[:, -1:]).permute(1, 3, 0, 2, 4)           # (1, 1, T, q, c)
        logits = self.refiner.pred_class(dec)[-1].transpose(1, 2)                         # (1, T, q, K+1)
        mean_logits = logits[0].float().mean(0)                                           # post_processing, py:763-767
        aux_logits = track["pred_logits"][0].float().mean(0)
        scores, labels, query = post.select_vis(mean_logits, aux_logits)
        emb = self.refiner.mask_embed(dec[0, 0].index_select(1, query)).float()           # (T, n, Cm)
        return torch.cat([emb.flatten(), scores.float(), labels.float(), query.float()])  # ints < 2^24: exact in fp32

    @torch.no_grad()
    def vis_from_payload(self, payload, mask_features, n, img_size, output_size, first_resize_size=None, packed=False):
        """Second half: this rank's final masks from the payload and its local mask features."""
        from . import ops
        t_local = mask_features.shape[0]
        T = (payload.numel() - 3 * n) // (n * self.refiner.mask_embed.layers[-1].out_features)
        emb = payload[:payload.numel() - 3 * n].view(T, n, -1)
        tail = payload[payload.numel() - 3 * n:].view(3, n)
        t0 = self.rank * t_local
        low = self.refiner._masks(emb[t0:t0 + t_local][None, None], mask_features[None])[0, 0]   # (n, t_local, h, w), frame-major
        h, w = low.shape[-2:]
        first = first_resize_size if first_resize_size is not None else (4 * h, 4 * w)    # stride-4 mask features
        masks = ops.vis_masks(low, None, first, img_size, output_size, packed=packed)
        return {"pred_scores": tail[0], "pred_labels": tail[1].long(), "pred_ids": tail[2].long(), "pred_masks": masks}

    @torch.no_grad()
    def temporal_stage(self, seg, mask_features):
        """Exchange + tracker + refiner + final masks, given this rank's segmenter outputs (`seg`: the predictor's
        dict for t_local frames) and its local mask features (t_local, C, H, W)."""
        C = seg["pred_embds"].shape[1]
        t_local = mask_features.shape[0]
        block = self.gather_queries(self.pack_queries(seg))
        frame_embds, frame_embds_no_norm, _ = self.unpack_queries(block, C)
        # the offline model discards the tracker's masks (meta_architecture.py:1486): skip them
        track = self.tracker(frame_embds, None, resume=False, frame_embeds_no_norm=frame_embds_no_norm, with_masks=False)
        outputs = self.refiner.refine(track["pred_embds"], frame_embds_no_norm)           # (T, l, q, 1, c)
        last = outputs[:, -1:]
        dec = self.refiner.decoder_norm(last).permute(1, 3, 0, 2, 4)                      # (1, 1, T, q, c)
        logits = self.refiner.pred_class(dec)[-1].transpose(1, 2)                         # (1, T, q, K+1)
        t0 = self.rank * t_local
        masks = self.refiner.predict_masks(outputs[t0:t0 + t_local], mask_features[None])
        return {"pred_logits": logits, "pred_masks": masks, "pred_embds": dec[0].permute(0, 3, 1, 2),
                "online_pred_logits": track["pred_logits"]}


class OnlineClipRunner:
    """DVIS_Plus_online.run_window_inference (P/dvis_Plus/meta_architecture.py:774-816) between the backbone and the
    post-processing: windows of `window_size` frames go through the segmenter head (pixel decoder + predictor) and the
    referring tracker, which carries its state from window to window (`resume=True` from the second window on, or from the
    first when `keep` continues a previous call, py:793-797).  The reference moves every window's logits / masks / embeddings
    to the host in fp32 (py:800-802) to save GPU memory; here they stay on the device for the fused post-processing
    (modules.postprocess.VideoPostProcessor).  BASELINE config 3 (T=5 clip, Q=200, tracker cross-attention)."""

    def __init__(self, pixel_decoder, predictor, tracker, window_size=30):
        self.pixel_decoder, self.predictor, self.tracker = pixel_decoder, predictor, tracker
        self.window_size = window_size

    @torch.no_grad()
    def __call__(self, features, keep=False):
        """features: backbone maps of the clip, dict name -> (T, C_i, H_i, W_i).
        -> pred_logits (1, T, Q, K+1), pred_masks (1, Q, T, H/4, W/4), pred_embds (1, C, T, Q)."""
        T = next(iter(features.values())).shape[0]
        logits, masks, embds = [], [], []
        for i, start in enumerate(range(0, T, self.window_size)):
            window = {k: v[start:start + self.window_size] for k, v in features.items()}
            mask_features, _, multi_scale = self.pixel_decoder.forward_features(window)
            out = self.predictor(multi_scale, mask_features)
            track = self.tracker(out["pred_embds"], mask_features.unsqueeze(0), resume=(i != 0 or keep),
                                 frame_embeds_no_norm=out["pred_embds_without_norm"])
            logits.append(track["pred_logits"].float())
            masks.append(track["pred_masks"])
            embds.append(track["pred_embds"].float())
        return {"pred_logits": torch.cat(logits, dim=1), "pred_masks": torch.cat(masks, dim=2), "pred_embds": torch.cat(embds, dim=2)}


class DAQOnlineRunner:
    """DVIS_DAQ_online.run_window_inference (D/dvis_daq/meta_architecture.py:488-597) between the backbone and the
    post-processing -- BASELINE config 5's clip loop.  Windows of 5 frames (hard-coded in the reference, py:491) go through
    the segmenter head; its confident queries (class score above `aux_inference_select_thr`, py:515-516) are offered to the
    VideoInstanceCutter as candidates for new instances; the cutter keeps one VideoInstanceSequence per object in its hub.
    After the last window every sequence that lived for at least `noise_frame_num` frames (or is still alive at the clip
    end, py:540-544) becomes one row of the outputs: time-averaged class logits, per-frame masks (-1e4 where the object
    does not exist), per-frame logits, a padding mask.  Dead sequences leave the hub (py:576-577).

    `segment(window_features)` defaults to pixel decoder + predictor; mask logits stay wherever `to_store` says
    (the reference parks them on the host; "cpu" reproduces that, None keeps them on the device)."""

    window_size = 5

    def __init__(self, pixel_decoder, predictor, cutter, num_classes, aux_inference_select_thr, noise_frame_num=2, segment=None,
                 to_store=None):
        self.pixel_decoder, self.predictor, self.cutter = pixel_decoder, predictor, cutter
        self.num_classes, self.select_thr, self.noise_frame_num = num_classes, aux_inference_select_thr, noise_frame_num
        self.segment = segment if segment is not None else self._segment
        self.to_store = to_store

    def _segment(self, window):
        mask_features, _, multi_scale = self.pixel_decoder.forward_features(window)
        return self.predictor(multi_scale, mask_features)

    @torch.no_grad()
    def __call__(self, features, keep=False, long_video_start_fidx=-1):
        """features: dict name -> (T, C_i, H_i, W_i) backbone maps of the clip (anything `segment` can slice by frame).
        -> the reference's dict: pred_logits (1, n, K+1), pred_masks (1, n, T, h, w), pred_ids (1, n), shape,
        padding_masks (1, n, T), full_logits (1, n, T, K+1); empty lists when no instance survives."""
        video_start = max(long_video_start_fidx, 0)
        num_frames = next(iter(features.values())).shape[0]
        H = W = dev = None
        for i, start in enumerate(range(0, num_frames, self.window_size)):
            out = self.segment({k: v[start:start + self.window_size] for k, v in features.items()})
            frame_embds = out["pred_embds"]                                              # (1, c, t, q)
            mask_features = out["mask_features"].unsqueeze(0)
            logits = out["pred_logits"][0].float()                                       # (t, q, K+1)
            masks = out["pred_masks"][0].transpose(0, 1)                                 # (t, q, h, w)
            H, W, dev = mask_features.shape[-2], mask_features.shape[-1], frame_embds.device
            valid = logits.softmax(dim=-1)[..., :-1].max(dim=-1)[0] > self.select_thr    # (t, q)
            frame_info = {"pred_logits": [[l] for l in logits], "pred_masks": [[m] for m in masks], "valid": [[v] for v in valid],
                          "seg_query_feat": self.predictor.query_feat, "seg_query_embed": self.predictor.query_embed}
            self.cutter.inference(frame_embds, mask_features, frame_info, video_start + start, resume=(i != 0 or keep),
                                  to_store=self.to_store if self.to_store is not None else dev)
        store = self.to_store if self.to_store is not None else dev
        rows, dead = [], []
        for seq_id, seq in self.cutter.video_ins_hub.items():
            n = len(seq.pred_masks)
            if n < self.noise_frame_num and seq.sT + n < video_start + num_frames:       # too short and already gone: noise
                continue
            first = max(video_start - seq.sT, 0)                                         # frames before this clip are skipped
            if first >= n:
                continue
            full_masks = torch.full((num_frames, H, W), -1e4, dtype=torch.float32, device=store)
            full_logits = torch.full((num_frames, self.num_classes + 1), -1e4, dtype=torch.float32, device=dev)
            full_logits[:, -1] = 1.0
            padding = torch.ones(num_frames, dtype=torch.bool)
            t0 = seq.sT + first - video_start
            full_masks[t0:t0 + n - first] = torch.stack(seq.pred_masks[first:]).to(full_masks)
            seq_logits = torch.stack(seq.pred_logits[first:]).to(full_logits)
            full_logits[t0:t0 + n - first] = seq_logits
            padding[t0:t0 + n - first] = False
            rows.append((seq_logits.mean(0), full_masks, full_logits, padding, seq_id))
            if seq.dead:
                dead.append(seq_id)
        for seq_id in dead:                                                              # long videos: free finished objects
            self.cutter.video_ins_hub.pop(seq_id)
        if rows:
            outputs = {"pred_logits": torch.stack([r[0] for r in rows])[None], "pred_masks": torch.stack([r[1] for r in rows])[None],
                       "pred_ids": torch.as_tensor([r[4] for r in rows], dtype=torch.int64)[None],
                       "padding_masks": torch.stack([r[3] for r in rows])[None], "full_logits": torch.stack([r[2] for r in rows])[None]}
        else:
            outputs = {"pred_logits": [], "pred_masks": [], "pred_ids": [], "padding_masks": [], "full_logits": []}
        outputs["shape"] = (H, W)
        return outputs


class DAQOfflineRunner:
    """DVIS_DAQ_offline.run_window_inference (D/dvis_daq/meta_architecture.py:1332-1365, with common_inference py:1169-1330
    and minvis_post_processing py:1400-1437) between the backbone and the post-processing:

      1. the segmenter head over all windows of the clip, results concatenated (py:1139-1167);
      2. the VideoInstanceCutter over the same windows, fed with the confident segmenter queries (py:1186-1221);
      3. every instance sequence that survived (>= `noise_frame_num` frames, or still alive at the clip end) becomes a row:
         time-averaged logits, per-frame masks (-1e4 outside its life span), its per-frame track queries padded with its
         position embedding at both ends, a padding mask, its id (py:1229-1270);
      4. the `offline_topk_ins` best rows by class score (py:1285-1296); if fewer than the cutter's `num_new_ins` remain, the
         free slots are filled with MinVIS-linked segmenter queries ranked by score (py:1298-1311);
      5. the DAQ TemporalRefiner over (instances x frames) gives the final logits and masks (py:1353-1356).

    `segment(window_features)` defaults to pixel decoder + predictor.  Mask logits of the online stage live on `to_store`
    ("cpu" parks them on the host like the reference, None keeps them on the device)."""

    def __init__(self, pixel_decoder, predictor, cutter, refiner, num_classes, aux_inference_select_thr, noise_frame_num=2,
                 offline_topk_ins=20, window_size=30, segment=None, to_store=None):
        self.pixel_decoder, self.predictor, self.cutter, self.refiner = pixel_decoder, predictor, cutter, refiner
        self.num_classes, self.select_thr, self.noise_frame_num = num_classes, aux_inference_select_thr, noise_frame_num
        self.offline_topk_ins, self.window_size = offline_topk_ins, window_size
        self.segment = segment if segment is not None else self._segment
        self.to_store = to_store

    def _segment(self, window):
        mask_features, _, multi_scale = self.pixel_decoder.forward_features(window)
        return self.predictor(multi_scale, mask_features)

    @torch.no_grad()
    def __call__(self, features, keep=False, long_video_start_fidx=-1):
        """-> {"pred_logits": (1, n, K+1), "pred_masks": (1, n, T, h, w), "pred_ids": (1, n), "shape": (h, w)}
        (empty lists when exactly one instance slot is left, like the reference's `instance_embeds.shape[-1] == 1` case)."""
        from .modules.postprocess import minvis_link_indices
        video_start = max(long_video_start_fidx, 0)
        num_frames = next(iter(features.values())).shape[0]
        windows = [(s, min(s + self.window_size, num_frames)) for s in range(0, num_frames, self.window_size)]
        outs = [self.segment({k: v[s:e] for k, v in features.items()}) for s, e in windows]
        frame_embeds = torch.cat([o["pred_embds"] for o in outs], dim=2)                       # (1, c, T, q)
        mask_features = torch.cat([o["mask_features"] for o in outs], dim=0).unsqueeze(0)      # (1, T, cm, h, w)
        seg_logits = torch.cat([o["pred_logits"] for o in outs], dim=1).float()                # (1, T, q, K+1)
        seg_masks = torch.cat([o["pred_masks"] for o in outs], dim=2)                          # (1, q, T, h, w)
        dev = frame_embeds.device
        store = self.to_store if self.to_store is not None else dev
        C, T = frame_embeds.shape[1], frame_embeds.shape[2]
        H, W = mask_features.shape[-2:]
        logits_t, masks_t = seg_logits[0], seg_masks[0].transpose(0, 1)                        # (T, q, K+1), (T, q, h, w)
        valid = logits_t.softmax(dim=-1)[..., :-1].max(dim=-1)[0] > self.select_thr
        for i, (s, e) in enumerate(windows):
            info = {"valid": [[v] for v in valid[s:e]], "pred_logits": [[l] for l in logits_t[s:e]],
                    "pred_masks": [[m] for m in masks_t[s:e]], "seg_query_feat": self.predictor.query_feat,
                    "seg_query_embed": self.predictor.query_embed}
            self.cutter.inference(frame_embeds[:, :, s:e], mask_features[:, s:e], info, video_start + s,
                                  resume=(i != 0 or keep), to_store=store)
        rows, dead = [], []
        for seq_id, seq in self.cutter.video_ins_hub.items():
            n = len(seq.pred_masks)
            if n < self.noise_frame_num and seq.sT + n < video_start + num_frames:
                continue
            first = max(video_start - seq.sT, 0)
            if first >= n:
                continue
            t0 = seq.sT + first - video_start
            full_masks = torch.full((num_frames, H, W), -1e4, dtype=torch.float32, device=store)
            full_masks[t0:t0 + n - first] = torch.stack(seq.pred_masks[first:]).to(full_masks)
            mean_logits = torch.stack(seq.pred_logits[first:]).float().mean(0)
            front = seq.sT - video_start                                                        # py:1249 (long videos)
            tail = num_frames - len(seq.embeds) - front
            pad = self.refiner.padding_embed(seq.similarity_guided_pos_embed)
            queries = torch.cat([pad[None].repeat(front, 1), torch.stack(seq.embeds), pad[None].repeat(tail, 1)])
            padding = torch.tensor([True] * front + [False] * len(seq.embeds) + [True] * tail, device=dev)
            rows.append((mean_logits, full_masks, queries, padding, seq_id))
            if seq.dead:
                dead.append(seq_id)
        if rows:
            online_logits = torch.stack([r[0] for r in rows])[None]                             # (1, n, K+1)
            online_masks = torch.stack([r[1] for r in rows])[None]                              # (1, n, T, h, w)
            trc_queries = torch.stack([r[2] for r in rows])[None]                               # (1, n, T, c)
            padding_masks = torch.stack([r[3] for r in rows])[None]                             # (1, n, T)
            seq_ids = torch.as_tensor([r[4] for r in rows], dtype=torch.int64, device=dev)
        else:
            online_logits = seg_logits.new_zeros((1, 0, seg_logits.shape[-1]))
            online_masks = torch.zeros((1, 0, T, H, W), dtype=torch.float32, device=store)
            trc_queries = seg_logits.new_zeros((1, 0, T, C))
            padding_masks = torch.zeros((1, 0, T), dtype=torch.bool, device=dev)
            seq_ids = torch.zeros(0, dtype=torch.int64, device=dev)
        scores = online_logits[0].softmax(dim=-1)[:, :-1].max(dim=-1)[0]
        top = torch.arange(scores.shape[0], device=dev) if self.offline_topk_ins > scores.shape[0] else \
            scores.topk(self.offline_topk_ins, sorted=False)[1]
        online_logits, trc_queries, padding_masks = online_logits[:, top], trc_queries[:, top], padding_masks[:, top]
        online_masks = online_masks[:, top.to(online_masks.device)]
        seq_id_list = seq_ids[top].tolist()
        num_left = self.cutter.num_new_ins - online_logits.shape[1]
        if num_left > 0:                                                                        # fill with naively linked queries
            idx = minvis_link_indices(frame_embeds[0].permute(1, 2, 0).float())                 # (T, q)
            t_ar = torch.arange(T, device=idx.device)[:, None]
            link_logits = logits_t[t_ar, idx].sum(0) / T                                        # (q, K+1)
            link_masks = masks_t[t_ar.to(masks_t.device), idx.to(masks_t.device)].transpose(0, 1)   # (q, T, h, w)
            link_embds = frame_embeds[0].permute(1, 2, 0).float()[t_ar, idx].transpose(0, 1)    # (q, T, c)
            best = link_logits.softmax(dim=-1)[:, :-1].max(dim=-1)[0].topk(num_left, sorted=False)[1]
            online_logits = torch.cat([online_logits, link_logits[best][None]], dim=1)
            online_masks = torch.cat([online_masks, link_masks[best.to(link_masks.device)][None].to(online_masks)], dim=1)
            trc_queries = torch.cat([trc_queries, link_embds[best][None]], dim=1)
            padding_masks = torch.cat([padding_masks, torch.zeros((1, num_left, num_frames), dtype=torch.bool, device=dev)], dim=1)
            seq_id_list += [(10000 + video_start) * 10000 + ii * 1000 for ii in range(1, num_left + 1)]
        for seq_id in dead:                                                                     # long videos: free finished objects
            self.cutter.video_ins_hub.pop(seq_id)
        instance_embeds = trc_queries.permute(0, 3, 2, 1)                                       # (1, c, T, n)
        if instance_embeds.shape[-1] == 1:                                                      # py:1348-1351
            return {"pred_logits": [], "pred_masks": [], "pred_ids": [], "shape": (H, W), "online_out": None}
        out = self.refiner(instance_embeds, padding_masks, frame_embeds, mask_features, None)
        return {"pred_logits": out["pred_logits"][:, 0], "pred_masks": out["pred_masks"],
                "pred_ids": torch.as_tensor(seq_id_list, dtype=torch.int64)[None], "shape": (H, W),
                "online_out": {"pred_logits": online_logits, "pred_masks": online_masks}}


class GraphedClipRunner:
    """The clip pipeline as CUDA graphs, software-pipelined across clips.

    Stage A (pixel decoder + predictor on this rank's frames) and stage B (tracker + refiner + final masks) are each
    captured once into a CUDA graph; the NCCL all-gather of the query block runs between them.  Nothing in either stage
    synchronises with the host (the Hungarian matching runs on the GPU), so a step is 2 graph launches + 1 collective
    instead of ~2000 kernel launches -- which is what keeps the host off the critical path when the per-rank work shrinks
    with the number of GPUs.  With `depth` = 2 two instances alternate and stage B of clip i (latency-bound, tiny kernels)
    overlaps stage A of clip i+1 (bandwidth / tensor bound) on a second stream.
    """

    def __init__(self, runner: OfflineClipRunner, example_features, depth=2, vis=None, d2h_stream=False, eager_b=False,
                 stream_a=None, stream_b=None):
        """d2h_stream: copy the results to the host on a stream of their own instead of the temporal stage's stream, so a
        clip's device->host copy no longer delays the NEXT clip's temporal stage (opt-in until timed on a B200).
        vis: None -> stage B ends with all Q mask logits (temporal_from_block); or a dict(post=VideoPostProcessor,
        img_size=, output_size=, first_resize_size=None, packed=False) -> stage B is vis_from_block: instances selected before
        the final mask GEMM, fused resize / threshold, outputs = final (optionally bit-packed) masks + scores / labels / ids."""
        self.r = runner
        self.depth = depth
        self.vis = vis
        # eager_b: stage B is not one big captured graph but is issued from the host on its stream (the tracker then replays
        # its own small per-frame graphs): measured with tests/perf/stage_overlap_probe.py
        self.eager_b = eager_b
        self.slots = []
        # stream_a / stream_b: optional externally created streams (e.g. streams of two green contexts = disjoint SM partitions,
        # dvis_plus_b200.partition.sm_partition_streams)
        self.stream_a = stream_a if stream_a is not None else torch.cuda.Stream()
        self.stream_b = stream_b if stream_b is not None else torch.cuda.Stream(priority=-1)   # latency-bound stage: its tiny kernels go first
        self.stream_c = torch.cuda.Stream()                      # host -> device copies of the next clip's inputs
        self.stream_d = torch.cuda.Stream() if d2h_stream else None   # device -> host copies of the results
        self.n = 0
        self.captured_launches = 0
        dev = next(iter(example_features.values())).device
        with torch.no_grad():
            for _ in range(2):                                   # populate caches / autotune outside capture
                blk, mf = runner.segment_stage(example_features)
                self._stage_b(runner.gather_queries(blk), mf, self._C(blk))
        torch.cuda.synchronize(dev)
        from . import _lib
        for _ in range(depth):
            n0 = _lib.launch_count
            slot = {"in": {k: v.clone() for k, v in example_features.items()}}
            ga = torch.cuda.CUDAGraph()
            with torch.cuda.graph(ga, stream=self.stream_a), torch.no_grad():
                slot["block"], slot["mf"] = runner.segment_stage(slot["in"])
            slot["ga"] = ga
            slot["gathered"] = torch.empty((runner.world * slot["block"].shape[0],) + tuple(slot["block"].shape[1:]),
                                           dtype=slot["block"].dtype, device=dev)
            gb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gb, stream=self.stream_b), torch.no_grad():
                slot["out"] = self._stage_b(slot["gathered"], slot["mf"], self._C(slot["block"]))
            slot["gb"] = gb
            slot["ev_a"] = torch.cuda.Event()
            slot["ev_b"] = torch.cuda.Event()
            slot["ev_c"] = torch.cuda.Event()
            slot["ev_o"] = torch.cuda.Event()                         # stage B done (results ready to be copied out)
            self.captured_launches = _lib.launch_count - n0       # libdvis_b200 kernels inside one clip's two graphs
            self.slots.append(slot)
        torch.cuda.synchronize(dev)

    def _C(self, block):
        K1 = self.r.refiner.class_embed.out_features
        return (block.shape[-1] - K1) // 2

    def _stage_b(self, gathered, mask_features, C):
        if self.vis is None:
            return self.r.temporal_from_block(gathered, mask_features, C)
        return self.r.vis_from_block(gathered, mask_features, C, **self.vis)

    def submit(self, features=None, d2h=None):
        """Enqueue one clip.  `features`: optional host (pinned) tensors copied into the slot's input buffers on the
        copy stream first; `d2h`: optional dict name -> pinned host tensor that receives slot["out"][name] after stage B.
        Returns the slot; its `out` is valid once slot["ev_b"] has completed."""
        slot = self.slots[self.n % self.depth]
        self.n += 1
        cur = torch.cuda.current_stream()
        if features is not None:
            self.stream_c.wait_stream(cur)
            self.stream_c.wait_event(slot["ev_a"])                # the slot's previous stage A has consumed its inputs
            with torch.cuda.stream(self.stream_c):
                for k, v in features.items():
                    slot["in"][k].copy_(v, non_blocking=True)
                slot["ev_c"].record(self.stream_c)
            self.stream_a.wait_event(slot["ev_c"])
        self.stream_a.wait_stream(cur)
        self.stream_a.wait_event(slot["ev_b"])                    # this slot's previous clip has finished with its buffers
        with torch.cuda.stream(self.stream_a):
            slot["ga"].replay()
            slot["ev_a"].record(self.stream_a)
        self.stream_b.wait_event(slot["ev_a"])
        with torch.cuda.stream(self.stream_b):
            if self.r.world > 1:
                dist.all_gather_into_tensor(slot["gathered"], slot["block"], group=self.r.group)
            else:
                slot["gathered"].copy_(slot["block"])
            if self.eager_b:
                with torch.no_grad():
                    slot["out"] = self._stage_b(slot["gathered"], slot["mf"], self._C(slot["block"]))
                if self.stream_d is not None:
                    for v in slot["out"].values():
                        v.record_stream(self.stream_d)
            else:
                slot["gb"].replay()
            if d2h is not None and self.stream_d is None:
                for k, v in d2h.items():
                    v.copy_(slot["out"][k], non_blocking=True)
            if d2h is None or self.stream_d is None:
                slot["ev_b"].record(self.stream_b)
            else:
                slot["ev_o"].record(self.stream_b)
        if d2h is not None and self.stream_d is not None:
            self.stream_d.wait_event(slot["ev_o"])
            with torch.cuda.stream(self.stream_d):                    # one copy stream: copies of successive clips stay ordered
                for k, v in d2h.items():
                    v.copy_(slot["out"][k], non_blocking=True)
                slot["ev_b"].record(self.stream_d)                    # the slot is free once its results have left
        return slot

    def wait_all(self):
        cur = torch.cuda.current_stream()
        cur.wait_stream(self.stream_a)
        cur.wait_stream(self.stream_b)
        cur.wait_stream(self.stream_c)
        if self.stream_d is not None:
            cur.wait_stream(self.stream_d)


class RoundRobinClipRunner:
    """Streams of clips over the G GPUs of one box with the temporal stage OWNED round-robin instead of replicated.

    OfflineClipRunner / GraphedClipRunner run tracker + refiner on every rank (identical results, no extra exchange): right
    for the latency of one clip, but for a stream of clips it is the Amdahl term of the strong scaling -- at 8 GPUs the
    replicated ~8 ms dwarf the 1.5 ms a rank spends on its two frames.  Here clip n's temporal stage runs only on rank
    n mod G; its result (mask embeddings + class logits, 3-10 MB) is broadcast and every rank finishes the masks of its
    own frames.  Per clip and rank that is 1/G of a temporal stage, overlapped with the other ranks' per-frame work:

        stream A   seg(n)  gather(n)  seg(n+1)  gather(n+1) ...                  (all ranks, every clip)
        stream T   temporal(n) for the clips this rank owns (n mod G == rank)
        stream B   broadcast(n)  masks(n)  broadcast(n+1) ...                    (2nd communicator: broadcasts never
                                                                                   queue behind the all-gathers)

    The temporal stage has a stream of its own: broadcast(n) completes only when clip n's owner has finished temporal(n), so
    on one in-order stream the owner of clip n+1 could not start temporal(n+1) before temporal(n) had finished on ANOTHER
    rank -- the temporal stages of the G ranks would run one after the other and nothing would be gained.

    Every rank issues the same collectives in the same order on each communicator.  `depth` slots (default G + 2) bound
    the clips in flight.  With CUDA tensors the three stages are captured into CUDA graphs per slot (like
    GraphedClipRunner); on CPU tensors (gloo, tests) everything runs eagerly and synchronously.  Results are identical to
    the replicated runners' (tests/test_pipeline_dist.py, world size 2)."""

    def __init__(self, runner: OfflineClipRunner, example_features, depth=None, graphs=True, vis=None):
        """vis: as in GraphedClipRunner -- the owner selects the instances, the payload shrinks to the selected instances'
        mask embeddings (T*n*Cm + 3n floats), every rank finishes fused-post-processed masks of its own frames."""
        self.r = runner
        self.vis = vis
        self.world, self.rank = runner.world, runner.rank
        self.depth = depth or self.world + 2
        self.n = 0
        self.cuda = next(iter(example_features.values())).is_cuda
        self.graphs = graphs and self.cuda
        # a second communicator for the broadcasts (every rank must create it, in the same order)
        # (built from the RUNNER's group, which may be a subgroup of the job -- e.g. one box of a multi-node run; new_group
        # is collective over the default group, so every process of the job must construct its runner)
        self.group_bc = None
        if self.world > 1:
            ranks = dist.get_process_group_ranks(runner.group) if runner.group is not None else list(range(dist.get_world_size()))
            self.group_bc = dist.new_group(ranks=ranks)
            self._global_rank_of = list(ranks)
        self.captured_launches = 0
        self.slots = []
        if self.cuda:
            self.stream_a, self.stream_b, self.stream_c = torch.cuda.Stream(), torch.cuda.Stream(priority=-1), torch.cuda.Stream()
            self.stream_t = torch.cuda.Stream(priority=-1)           # the owner's temporal stage (see the class docstring)
            with torch.no_grad():
                for _ in range(2):                                   # populate caches / autotune outside capture
                    blk, mf = runner.segment_stage(example_features)
                    self._finish(self._payload(runner.gather_queries(blk), self._C(blk)), mf, self._C(blk))
            torch.cuda.synchronize()
        for _ in range(self.depth):
            self.slots.append(self._make_slot(example_features))
        if self.cuda:
            torch.cuda.synchronize()

    def _C(self, block):
        return (block.shape[-1] - self.r.refiner.class_embed.out_features) // 2

    def _payload(self, gathered, C):
        if self.vis is None:
            return self.r.temporal_payload(gathered, C)
        return self.r.vis_payload(gathered, C, self.vis["post"])

    def _finish(self, payload, mask_features, C):
        if self.vis is None:
            return self.r.outputs_from_payload(payload, mask_features, C)
        v = self.vis
        return self.r.vis_from_payload(payload, mask_features, v["post"].max_num, v["img_size"], v["output_size"],
                                       v.get("first_resize_size"), v.get("packed", False))

    def _payload_shape(self, gathered):
        K1 = self.r.refiner.class_embed.out_features
        Cm = self.r.refiner.mask_embed.layers[-1].out_features
        if self.vis is not None:
            n = self.vis["post"].max_num
            return (gathered.shape[0] * n * Cm + 3 * n,)
        return (gathered.shape[0], gathered.shape[1], Cm + 2 * K1 + self._C(gathered))

    def _make_slot(self, example_features):
        r = self.r
        slot = {"in": {k: v.clone() for k, v in example_features.items()}}
        if not self.graphs:
            if self.cuda:                                             # eager on the device: static exchange buffers
                with torch.no_grad():
                    blk, _ = r.segment_stage(slot["in"])
                slot["gathered"] = blk.new_zeros((self.world * blk.shape[0],) + tuple(blk.shape[1:]))
                slot["payload"] = blk.new_zeros(self._payload_shape(slot["gathered"]))
                slot.update(ev_a=torch.cuda.Event(), ev_b=torch.cuda.Event(), ev_c=torch.cuda.Event(), ev_t=torch.cuda.Event())
            return slot
        from . import _lib
        n0 = _lib.launch_count
        ga = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ga, stream=self.stream_a), torch.no_grad():
            slot["block"], slot["mf"] = r.segment_stage(slot["in"])
        C = self._C(slot["block"])
        slot["gathered"] = torch.zeros((self.world * slot["block"].shape[0],) + tuple(slot["block"].shape[1:]),
                                       dtype=slot["block"].dtype, device=slot["block"].device)
        gt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gt, stream=self.stream_t), torch.no_grad():   # captured where it replays: own cuBLAS workspace
            slot["payload"] = self._payload(slot["gathered"], C)
        gm = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gm, stream=self.stream_b), torch.no_grad():
            slot["out"] = self._finish(slot["payload"], slot["mf"], C)
        slot.update(ga=ga, gt=gt, gm=gm, ev_a=torch.cuda.Event(), ev_b=torch.cuda.Event(), ev_c=torch.cuda.Event(),
                    ev_t=torch.cuda.Event())
        self.captured_launches = _lib.launch_count - n0      # one clip's kernels if this rank owned every temporal stage
        return slot

    def _exchange(self, slot, owner):
        if self.world > 1:
            dist.broadcast(slot["payload"], src=self._global_rank_of[owner], group=self.group_bc)   # src is a GLOBAL rank

    @torch.no_grad()
    def submit(self, features=None, d2h=None):
        """Enqueue one clip (same contract as GraphedClipRunner.submit).  Returns the slot; slot["out"] is valid once
        slot["ev_b"] has completed (immediately on CPU)."""
        r = self.r
        slot = self.slots[self.n % self.depth]
        owner = self.n % self.world
        self.n += 1
        if not self.cuda:                                             # eager, synchronous (gloo / tests)
            feats = features if features is not None else slot["in"]
            slot["block"], slot["mf"] = r.segment_stage(feats)
            C = self._C(slot["block"])
            slot["gathered"] = r.gather_queries(slot["block"])
            slot["payload"] = self._payload(slot["gathered"], C) if owner == self.rank else \
                slot["gathered"].new_empty(self._payload_shape(slot["gathered"]))
            self._exchange(slot, owner)
            slot["out"] = self._finish(slot["payload"], slot["mf"], C)
            return slot
        cur = torch.cuda.current_stream()
        if features is not None:
            self.stream_c.wait_stream(cur)
            self.stream_c.wait_event(slot["ev_a"])                    # the slot's previous stage A has consumed its inputs
            with torch.cuda.stream(self.stream_c):
                for k, v in features.items():
                    slot["in"][k].copy_(v, non_blocking=True)
                slot["ev_c"].record(self.stream_c)
            self.stream_a.wait_event(slot["ev_c"])
        self.stream_a.wait_stream(cur)
        self.stream_a.wait_event(slot["ev_b"])                        # the slot's previous clip is completely finished
        with torch.cuda.stream(self.stream_a):
            if self.graphs:
                slot["ga"].replay()
            else:
                slot["block"], slot["mf"] = r.segment_stage(slot["in"])
            if self.world > 1:
                dist.all_gather_into_tensor(slot["gathered"], slot["block"], group=r.group)
            else:
                slot["gathered"].copy_(slot["block"])
            slot["ev_a"].record(self.stream_a)
        C = self._C(slot["block"])
        if owner == self.rank:
            self.stream_t.wait_event(slot["ev_a"])
            with torch.cuda.stream(self.stream_t):
                if self.graphs:
                    slot["gt"].replay()
                else:
                    slot["payload"].copy_(self._payload(slot["gathered"], C))
                slot["ev_t"].record(self.stream_t)
            self.stream_b.wait_event(slot["ev_t"])
        self.stream_b.wait_event(slot["ev_a"])
        with torch.cuda.stream(self.stream_b):
            self._exchange(slot, owner)
            if self.graphs:
                slot["gm"].replay()
            else:
                slot["out"] = self._finish(slot["payload"], slot["mf"], C)
            if d2h is not None:
                for k, v in d2h.items():
                    v.copy_(slot["out"][k], non_blocking=True)
            slot["ev_b"].record(self.stream_b)
        return slot

    def wait_all(self):
        if not self.cuda:
            return
        cur = torch.cuda.current_stream()
        cur.wait_stream(self.stream_a)
        cur.wait_stream(self.stream_b)
        cur.wait_stream(self.stream_c)
        cur.wait_stream(self.stream_t)
