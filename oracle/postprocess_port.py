"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Torch-CPU restatement of the reference's video post-processing.

Follows DVIS_Plus_online.post_processing / inference_video_vis / inference_video_vps / inference_video_vss
(P/dvis_Plus/meta_architecture.py:758-772, 818-868, 870-956, 958-979; P = /root/reference/DVIS_Plus), i.e. SURVEY.md
section 8f rank 2.  Plain torch ops on CPU fp32 -- the same F.softmax / topk / F.interpolate calls the reference makes --
written as free functions of explicit arguments instead of methods of the meta-architecture.  `dvis_plus_b200/` never
imports this file.

Parity pin: tests/test_oracle_postprocess.py compares every function with golden outputs produced by calling the
UNMODIFIED reference methods in the build container (tests/golden/make_golden_postprocess.py -> postprocess_*.pt).
Beyond the reference's outputs the functions can also return the resized mask logits / probabilities, which the
parity tests of the CUDA kernels use to tell decision-boundary pixels (|logit| ~ 0) from real mismatches.
"""
import torch
import torch.nn.functional as F


def post_processing(pred_logits, num_masks, aux_logits=None):
    """py:758-772: average the class logits over time, ids = arange(Q).
    pred_logits (1, T, Q, K+1) -> (1, Q, K+1); aux_logits (1, T, Q, K+1) -> (Q, K+1)."""
    out_logits = torch.mean(pred_logits[0], dim=0).unsqueeze(0)
    ids = [torch.arange(0, num_masks)]
    if aux_logits is not None:
        return out_logits, ids, torch.mean(aux_logits[0], dim=0)
    return out_logits, ids


def resize_chain(masks, img_size, output_height, output_width, first_resize_size, sigmoid=False):
    """py:838-844 (vis) / 889-895 (vps) / 968-974 (vss): bilinear resize to the padded input size, crop the padding,
    [sigmoid], bilinear resize to the output size.  masks (n, T, h, w) float."""
    m = F.interpolate(masks, size=first_resize_size, mode="bilinear", align_corners=False)
    m = m[:, :, : img_size[0], : img_size[1]]
    if sigmoid:
        m = m.sigmoid()
    return F.interpolate(m, size=(output_height, output_width), mode="bilinear", align_corners=False)


def vis_scores(pred_cls, aux_pred_cls=None):
    """py:823-827: softmax over classes without the no-object column, max with the online (aux) scores."""
    scores = F.softmax(pred_cls, dim=-1)[:, :-1]
    if aux_pred_cls is not None:
        scores = torch.maximum(scores, F.softmax(aux_pred_cls, dim=-1)[:, :-1].to(scores))
    return scores


def inference_video_vis(pred_cls, pred_masks, img_size, output_height, output_width, first_resize_size, pred_id,
                        num_classes, max_num, aux_pred_cls=None, return_logits=False):
    """py:818-868.  pred_cls (Q, K+1), pred_masks (Q, T, h, w), pred_id (Q,)."""
    if len(pred_cls) > 0:
        scores = vis_scores(pred_cls, aux_pred_cls)
        labels = torch.arange(num_classes).unsqueeze(0).repeat(scores.shape[0], 1).flatten(0, 1)
        scores_per_image, topk_indices = scores.flatten(0, 1).topk(max_num, sorted=False)
        labels_per_image = labels[topk_indices]
        topk_indices = topk_indices // num_classes
        resized = resize_chain(pred_masks[topk_indices].float(), img_size, output_height, output_width, first_resize_size)
        masks = resized > 0.0
        out = dict(image_size=(output_height, output_width), pred_scores=scores_per_image.tolist(),
                   pred_labels=labels_per_image.tolist(), pred_masks=[m for m in masks],
                   pred_ids=pred_id[topk_indices].tolist(), task="vis")
        if return_logits:
            out["resized_logits"] = resized
            out["query_indices"] = topk_indices
        return out
    return dict(image_size=(output_height, output_width), pred_scores=[], pred_labels=[], pred_masks=[], pred_ids=[],
                task="vis")


def inference_video_vss(pred_cls, pred_masks, img_size, output_height, output_width, first_resize_size,
                        aux_pred_cls=None, return_scores=False):
    """py:958-979: class-probability weighted sum of the mask probabilities, arg-max over classes."""
    mask_cls = F.softmax(pred_cls, dim=-1)[..., :-1]
    if aux_pred_cls is not None:
        mask_cls = torch.maximum(mask_cls, F.softmax(aux_pred_cls, dim=-1)[..., :-1].to(mask_cls))
    cur_masks = resize_chain(pred_masks.float(), img_size, output_height, output_width, first_resize_size, sigmoid=True)
    semseg = torch.einsum("qc,qthw->cthw", mask_cls, cur_masks)
    sem_score, sem_mask = semseg.max(0)
    out = dict(image_size=(output_height, output_width), pred_masks=sem_mask, task="vss")
    if return_scores:
        out["semseg"] = semseg
    return out


def inference_video_vps(pred_cls, pred_masks, img_size, output_height, output_width, first_resize_size, pred_id,
                        num_classes, num_thing_classes, object_mask_threshold, overlap_threshold, aux_pred_cls=None,
                        return_probs=False):
    """py:870-956: keep confident non-background queries, per-pixel arg-max of score * mask probability, drop segments
    whose arg-max area is a small fraction of their own mask, merge stuff segments of one class."""
    pred_cls = F.softmax(pred_cls, dim=-1)
    if aux_pred_cls is not None:
        aux = F.softmax(aux_pred_cls, dim=-1)[:, :-1]
        pred_cls = pred_cls.clone()
        pred_cls[:, :-1] = torch.maximum(pred_cls[:, :-1], aux.to(pred_cls))
    scores, labels = pred_cls.max(-1)
    keep = labels.ne(num_classes) & (scores > object_mask_threshold)
    cur_scores, cur_classes, cur_ids = scores[keep], labels[keep], pred_id[keep]
    cur_masks = resize_chain(pred_masks[keep].float(), img_size, output_height, output_width, first_resize_size, sigmoid=True)
    cur_prob_masks = cur_scores.view(-1, 1, 1, 1) * cur_masks
    h, w = cur_masks.shape[-2:]
    panoptic_seg = torch.zeros((cur_masks.size(1), h, w), dtype=torch.int32)
    segments_infos, out_ids = [], []
    current_segment_id = 0
    out = dict(image_size=(output_height, output_width), pred_masks=panoptic_seg, segments_infos=segments_infos,
               pred_ids=out_ids, task="vps")
    if cur_masks.shape[0] == 0:
        return out
    cur_mask_ids = cur_prob_masks.argmax(0)
    stuff_memory_list = {}
    for k in range(cur_classes.shape[0]):
        pred_class = cur_classes[k].item()
        isthing = pred_class < num_thing_classes
        mask_area = (cur_mask_ids == k).sum().item()
        original_area = (cur_masks[k] >= 0.5).sum().item()
        mask = (cur_mask_ids == k) & (cur_masks[k] >= 0.5)
        if mask_area > 0 and original_area > 0 and mask.sum().item() > 0:
            if mask_area / original_area < overlap_threshold:
                continue
            if not isthing:
                if int(pred_class) in stuff_memory_list.keys():
                    panoptic_seg[mask] = stuff_memory_list[int(pred_class)]
                    continue
                else:
                    stuff_memory_list[int(pred_class)] = current_segment_id + 1
            current_segment_id += 1
            panoptic_seg[mask] = current_segment_id
            segments_infos.append({"id": current_segment_id, "isthing": bool(isthing), "category_id": int(pred_class)})
            out_ids.append(cur_ids[k])
    if return_probs:
        out["cur_masks"] = cur_masks
        out["cur_prob_masks"] = cur_prob_masks
        out["keep"] = keep
    return out
