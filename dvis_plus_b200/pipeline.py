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
