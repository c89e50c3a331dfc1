"""Video post-processing of the DVIS++ meta-architectures on the GPU.

`VideoPostProcessor` carries the handful of attributes the reference methods read from the meta-architecture
(`num_queries`, `sem_seg_head.num_classes`, `max_num`, `object_mask_threshold`, `overlap_threshold`,
`metadata.thing_dataset_id_to_contiguous_id`; P/dvis_Plus/meta_architecture.py:484-499) and exposes the same methods with
the same signatures and result dictionaries:

    post_processing            DVIS_Plus_online.post_processing        py:758-772   (mean class logits + query ids)
    post_processing_minvis     MinVIS.post_processing                  py:255-301   (frame-by-frame Hungarian re-ordering)
    inference_video            MinVIS.inference_video                  py:361-401
    inference_video_vis        DVIS_Plus_online.inference_video_vis    py:818-868
    inference_video_vps        DVIS_Plus_online.inference_video_vps    py:870-956
    inference_video_vss        DVIS_Plus_online.inference_video_vss    py:958-979

The reference up-samples every kept query's mask logits to the padded input size in fp32, crops, resizes again and only
then thresholds / arg-maxes, all through materialised (n, T, H, W) fp32 tensors.  Here the chain is evaluated per output
pixel by the kernels of csrc/postproc.cu (ops.vis_masks / vps_argmax / vss_argmax); only the final bool / int32 / int64
result is ever written.  There is no CPU path: CPU tensors raise.
"""
import torch

from .. import ops


def minvis_link_indices(embds):
    """MinVIS.match_from_embds chained over the frames (py:255-264, 284-290): embds (t, q, c) -> (t, q) int64 indices such
    that frame i's queries re-ordered by indices[i] line up with frame i-1's re-ordered queries (frame 0: identity).
    The optimum of a row-permuted assignment problem is the permuted optimum, so indices_t = sigma_t[indices_{t-1}] with
    sigma_t solved against the UN-permuted previous frame: on the device all T-1 problems run as one batched launch of the
    GPU Hungarian kernel (ops.lap_chain), no host sync; CPU tensors (training-time / test use) go through SciPy."""
    T, Q = embds.shape[:2]
    ident = torch.arange(Q, device=embds.device)[None]
    if T == 1:
        return ident
    n = embds / embds.norm(dim=2, keepdim=True)                                  # py:256-257 (no epsilon)
    cost = 1 - torch.bmm(n[:-1], n[1:].transpose(1, 2))                          # (t-1, q_prev, q_cur) == C.T of py:259-262
    if cost.is_cuda:
        return torch.cat([ident, ops.lap_chain(cost)[1]], 0)
    from scipy.optimize import linear_sum_assignment
    idx, rows = ident[0], [ident[0]]
    for t in range(T - 1):
        sigma = torch.as_tensor(linear_sum_assignment(cost[t].numpy())[1], dtype=torch.int64)
        idx = sigma[idx]
        rows.append(idx)
    return torch.stack(rows, 0)


class VideoPostProcessor:
    def __init__(self, num_classes, num_queries=None, max_num=20, object_mask_threshold=0.8, overlap_threshold=0.8,
                 num_thing_classes=0, task="vis"):
        self.num_classes = num_classes
        self.num_queries = num_queries            # informational; like DVIS-DAQ the methods use the actual query count
        self.max_num = max_num
        self.object_mask_threshold = object_mask_threshold
        self.overlap_threshold = overlap_threshold
        self.num_thing_classes = num_thing_classes   # len(metadata.thing_dataset_id_to_contiguous_id), py:921
        assert task in ("vis", "vps", "vss"), "Only support vis, vss and vps !"   # py:490
        self.task = task
        self.inference_video_task = {"vis": self.inference_video_vis, "vss": self.inference_video_vss,
                                     "vps": self.inference_video_vps}[task]       # py:494-499

    # ------------------------------------------------------------------------------------------------------------
    def post_processing(self, outputs, aux_logits=None):
        """py:758-772: average the class logits over the frames and append the query ids."""
        pred_logits = outputs["pred_logits"][0]                                  # (t, q, c)
        outputs["pred_logits"] = torch.mean(pred_logits, dim=0).unsqueeze(0)
        outputs["ids"] = [torch.arange(0, outputs["pred_masks"].size(1))]
        if aux_logits is not None:
            return outputs, torch.mean(aux_logits[0], dim=0)
        return outputs

    def post_processing_minvis(self, outputs):
        """py:255-301: align every frame's queries to the previous (already aligned) frame by Hungarian matching on the
        cosine distance of the query embeddings, average the logits, stack the re-ordered masks.

        The reference solves T-1 dependent problems with one device->host copy + SciPy call each (py:261-262).  The
        optimum of a row-permuted assignment problem is the permuted optimum, so indices_t = sigma_t[indices_{t-1}] with
        sigma_t solved against the UN-permuted previous frame: all problems are independent and run as one batched
        launch of the GPU Hungarian kernel (ops.lap_chain), no host sync."""
        pred_logits, pred_masks, pred_embds = outputs["pred_logits"][0], outputs["pred_masks"][0], outputs["pred_embds"][0]
        T = pred_logits.shape[0]
        idx = minvis_link_indices(pred_embds.permute(1, 2, 0).float())
        t_ar = torch.arange(T, device=idx.device)[:, None]
        out_logits = pred_logits[t_ar, idx].sum(0) / T                           # sum(out_logits) / len(out_logits)
        out_masks = pred_masks.permute(1, 0, 2, 3)[t_ar, idx].permute(1, 0, 2, 3)  # (q, t, h, w)
        outputs["pred_logits"] = out_logits.unsqueeze(0)
        outputs["pred_masks"] = out_masks.unsqueeze(0)
        return outputs

    # ------------------------------------------------------------------------------------------------------------
    def select_vis(self, pred_cls, aux_pred_cls=None):
        """The instance selection of inference_video_vis (py:823-835) alone: -> (scores, labels, query indices), each
        (max_num,) on the device, score descending.  Lets a caller pick the instances BEFORE the final mask GEMM so that
        only max_num (not Q) masks per frame are ever computed (pipeline.OfflineClipRunner.vis_inference)."""
        return ops.vis_topk(pred_cls, self.max_num, aux_pred_cls)

    # bool masks cross PCIe as 1 bit per pixel and are unpacked on the host (8x less device->host); False copies bytes
    packed_transfer = True

    def inference_video_vis(self, pred_cls, pred_masks, img_size, output_height, output_width, first_resize_size, pred_id,
                            aux_pred_cls=None, masks_on_device=False):
        """py:818-868.  pred_cls (Q, K+1); pred_masks (Q, T, h, w) f32|bf16 device view (any query / frame strides);
        pred_id (Q,).  `pred_masks` of the result: list of (T, H_out, W_out) bool CPU tensors like the reference
        (`masks_on_device=True` keeps them on the GPU as one (n, T, H_out, W_out) tensor's rows)."""
        if len(pred_cls) > 0:
            scores, labels, query = self.select_vis(pred_cls, aux_pred_cls)
            out_size = (output_height, output_width)
            if masks_on_device or not self.packed_transfer:
                masks = ops.vis_masks(pred_masks, query, first_resize_size, img_size, out_size)
                masks = masks if masks_on_device else masks.cpu()
            else:
                masks = ops.unpack_masks(ops.vis_masks(pred_masks, query, first_resize_size, img_size, out_size, packed=True).cpu(),
                                         output_width)
            pred_ids = torch.as_tensor(pred_id, device=query.device)[query]
            out_scores = scores.tolist()
            out_labels = labels.tolist()
            out_ids = pred_ids.tolist()
            out_masks = [m for m in masks]
        else:
            out_scores, out_labels, out_masks, out_ids = [], [], [], []
        return {"image_size": (output_height, output_width), "pred_scores": out_scores, "pred_labels": out_labels,
                "pred_masks": out_masks, "pred_ids": out_ids, "task": "vis"}

    def inference_video(self, pred_cls, pred_masks, img_size, output_height, output_width, first_resize_size, masks_on_device=False):
        """MinVIS.inference_video (py:361-401): the ten best (query, class) pairs, no ids, no online scores."""
        saved, self.max_num = self.max_num, 10                                   # py:369 hard-codes topk(10)
        try:
            out = self.inference_video_vis(pred_cls, pred_masks, img_size, output_height, output_width, first_resize_size,
                                           torch.arange(len(pred_cls)), masks_on_device=masks_on_device)
        finally:
            self.max_num = saved
        return {k: out[k] for k in ("image_size", "pred_scores", "pred_labels", "pred_masks")}

    def inference_video_vss(self, pred_cls, pred_masks, img_size, output_height, output_width, first_resize_size,
                            pred_id=None, aux_pred_cls=None, masks_on_device=False):
        """py:958-979."""
        mask_cls = ops.class_scores(pred_cls, aux_pred_cls)[:, :-1]
        sem_mask = ops.vss_argmax(pred_masks, mask_cls, first_resize_size, img_size, (output_height, output_width))
        return {"image_size": (output_height, output_width), "pred_masks": sem_mask if masks_on_device else sem_mask.cpu(),
                "task": "vss"}

    def inference_video_vps(self, pred_cls, pred_masks, img_size, output_height, output_width, first_resize_size, pred_id,
                            aux_pred_cls=None, masks_on_device=False):
        """py:870-956.  The per-pixel arg-max and the three areas every kept query needs come from ONE kernel; the
        sequential segment bookkeeping (overlap filter, stuff merging) then runs on 3*n integers on the host, and a second
        kernel paints the panoptic map.  Two small device<->host exchanges instead of ~4 synchronising reductions over
        full-resolution tensors per kept query."""
        scores_all = ops.class_scores(pred_cls, aux_pred_cls)
        scores, labels = scores_all.max(-1)
        keep = labels.ne(self.num_classes) & (scores > self.object_mask_threshold)
        keep_idx = keep.nonzero().flatten()                                      # host sync (the reference's boolean indexing syncs too)
        T = pred_masks.shape[1]
        dev = scores_all.device
        segments_infos, out_ids = [], []
        if keep_idx.numel() == 0:
            panoptic = torch.zeros((T, output_height, output_width), dtype=torch.int32, device=dev)
            return {"image_size": (output_height, output_width), "pred_masks": panoptic if masks_on_device else panoptic.cpu(),
                    "segments_infos": segments_infos, "pred_ids": out_ids, "task": "vps"}
        cur_scores = scores[keep_idx].contiguous()
        win, areas = ops.vps_argmax(pred_masks, keep_idx, cur_scores, first_resize_size, img_size, (output_height, output_width))
        cur_classes = labels[keep_idx].tolist()
        cur_ids = torch.as_tensor(pred_id).cpu()[keep_idx.cpu()]
        mask_area, original_area, inter_area = areas.cpu().tolist()
        seg_of_k = [0] * len(cur_classes)
        current_segment_id = 0
        stuff_memory_list = {}
        for k, pred_class in enumerate(cur_classes):
            isthing = pred_class < self.num_thing_classes
            if mask_area[k] > 0 and original_area[k] > 0 and inter_area[k] > 0:
                if mask_area[k] / original_area[k] < self.overlap_threshold:
                    continue
                if not isthing:
                    if int(pred_class) in stuff_memory_list.keys():
                        seg_of_k[k] = stuff_memory_list[int(pred_class)]
                        continue
                    else:
                        stuff_memory_list[int(pred_class)] = current_segment_id + 1
                current_segment_id += 1
                seg_of_k[k] = current_segment_id
                segments_infos.append({"id": current_segment_id, "isthing": bool(isthing), "category_id": int(pred_class)})
                out_ids.append(cur_ids[k])
        panoptic = ops.vps_paint(win, torch.tensor(seg_of_k, dtype=torch.int32, device=dev))
        return {"image_size": (output_height, output_width), "pred_masks": panoptic if masks_on_device else panoptic.cpu(),
                "segments_infos": segments_infos, "pred_ids": out_ids, "task": "vps"}
