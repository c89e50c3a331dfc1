"""The per-pixel arithmetic of the post-processing kernels (dvis_plus_b200/csrc/resize_core.cuh), compiled for the HOST
by tests/hostcore/, against the oracle and the reference's golden vectors.  This checks on CPU what the CUDA kernels of
csrc/postproc.cu compute per pixel (the -m gpu tests in test_postprocess_gpu.py check the kernels themselves)."""
import os
import sys

import pytest
import torch

from oracle import postprocess_port as pp
from postproc_util import assert_labels_match, assert_masks_match

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostcore"))
import hostcore_binding as hc  # noqa: E402

GEOMS = [  # (h, w), first resize, image size, output size
    ((12, 20), (48, 80), (45, 78), (45, 78)),      # identity second resize: strip walker
    ((12, 20), (48, 80), (45, 78), (67, 117)),     # up-scaling second resize
    ((12, 20), (48, 80), (45, 78), (30, 52)),      # down-scaling second resize
    ((12, 20), (48, 80), (48, 80), (48, 80)),      # no padding, width a multiple of 8 (vector-store path on the GPU)
    ((7, 9), (28, 36), (25, 33), (25, 33)),        # odd sizes, strip tail
    ((23, 40), (92, 160), (90, 160), (180, 320)),  # exact 2x second resize
    ((5, 6), (20, 24), (20, 24), (3, 2)),          # output smaller than the logits
    ((12, 20), (48, 80), (45, 78), (200, 301)),    # strong up-scale: intermediate rows reused over many output rows
    ((30, 30), (120, 120), (118, 119), (17, 13)),  # strong down-scale: source rows jump, nothing is reused
    ((46, 80), (184, 320), (180, 320), (100, 177)),  # several row bands per plane, odd width
]


@pytest.mark.parametrize("geom", GEOMS)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_vis_masks_vs_oracle(geom, dtype):
    (h, w), first, img, out = geom
    g = torch.Generator().manual_seed(h * 100 + w)
    masks = (torch.randn(6, 3, h, w, generator=g) * 3).to(dtype)
    sel = torch.tensor([4, 0, 4, 5], dtype=torch.int64)
    ours = hc.vis_masks(masks, sel, first, img, out)
    ref = pp.resize_chain(masks[sel].float(), img, out[0], out[1], first)
    assert_masks_match(ours, ref > 0, ref, tol=2e-5, max_boundary_frac=1e-3)
    # no selection + a strided (frame-major) layout: masks stored as (T, Q, h, w), viewed as (Q, T, h, w)
    fm = masks.transpose(0, 1).contiguous().transpose(0, 1)
    assert not fm.is_contiguous()
    ours2 = hc.vis_masks(fm, None, first, img, out)
    ref2 = pp.resize_chain(masks.float(), img, out[0], out[1], first)
    assert_masks_match(ours2, ref2 > 0, ref2, tol=2e-5, max_boundary_frac=1e-3)


def test_vis_masks_vs_reference_golden(golden):
    g = golden("postprocess_vis.pt")
    for name, c in g["cases"].items():
        Ho, Wo = c["output_size"]
        o = pp.inference_video_vis(g["pred_cls"], g["pred_masks"], g["img_size"], Ho, Wo, g["first_resize_size"], g["pred_id"],
                                   g["num_classes"], c["max_num"], aux_pred_cls=g["aux_cls"] if c["use_aux"] else None,
                                   return_logits=True)
        ours = hc.vis_masks(g["pred_masks"], o["query_indices"], g["first_resize_size"], g["img_size"], (Ho, Wo))
        assert_masks_match(ours, torch.stack(o["pred_masks"]), o["resized_logits"], tol=2e-5)


@pytest.mark.parametrize("geom", GEOMS[:5])
def test_vps_argmax_vs_oracle(geom):
    (h, w), first, img, out = geom
    g = torch.Generator().manual_seed(7 + h)
    masks = torch.randn(9, 2, h, w, generator=g) * 3
    keep_idx = torch.tensor([1, 3, 4, 8], dtype=torch.int64)
    keep_score = torch.tensor([0.9, 0.5, 0.7, 0.95])
    win, areas = hc.vps_argmax(masks, keep_idx, keep_score, first, img, out)
    cur = pp.resize_chain(masks[keep_idx], img, out[0], out[1], first, sigmoid=True)
    prob = keep_score.view(-1, 1, 1, 1) * cur
    ref_ids = prob.argmax(0)
    ids = torch.where(win >= 0, win, ~win).long()
    assert_labels_match(ids, ref_ids, prob, tol=1e-5)
    same = ids == ref_ids
    solid_ref = cur.gather(0, ref_ids[None])[0] >= 0.5
    margin = (cur.gather(0, ref_ids[None])[0] - 0.5).abs()
    assert ((win >= 0) == solid_ref)[same & (margin > 1e-5)].all()
    n = keep_idx.numel()
    ref_areas = torch.stack([torch.stack([(ref_ids == k).sum() for k in range(n)]),
                             torch.stack([(cur[k] >= 0.5).sum() for k in range(n)]),
                             torch.stack([((ref_ids == k) & (cur[k] >= 0.5)).sum() for k in range(n)])])
    assert (areas - ref_areas).abs().max().item() <= 3, (areas, ref_areas)
    assert areas[0].sum().item() == ref_ids.numel()


@pytest.mark.parametrize("geom", GEOMS[:3])
@pytest.mark.parametrize("K", [5, 19])
def test_vss_argmax_vs_oracle(geom, K):
    (h, w), first, img, out = geom
    g = torch.Generator().manual_seed(3 + K)
    masks = torch.randn(10, 2, h, w, generator=g) * 3
    cls = torch.randn(10, K + 1, generator=g) * 2
    mask_cls = cls.softmax(-1)[:, :-1]                       # row stride K + 1
    ours = hc.vss_argmax(masks, mask_cls, first, img, out)
    ref = pp.inference_video_vss(cls, masks, img, out[0], out[1], first, return_scores=True)
    assert_labels_match(ours, ref["pred_masks"], ref["semseg"], tol=1e-5)
