"""The post-processing oracle (oracle/postprocess_port.py) against golden outputs of the UNMODIFIED reference methods
(tests/golden/postprocess_*.pt, made by tests/golden/make_golden_postprocess.py).  CPU only."""
import torch

from oracle import postprocess_port as pp
from postproc_util import assert_labels_match, assert_masks_match, sort_instances


def test_vis_matches_reference(golden):
    g = golden("postprocess_vis.pt")
    for name, c in g["cases"].items():
        Ho, Wo = c["output_size"]
        out = pp.inference_video_vis(g["pred_cls"], g["pred_masks"], g["img_size"], Ho, Wo, g["first_resize_size"], g["pred_id"],
                                     g["num_classes"], c["max_num"], aux_pred_cls=g["aux_cls"] if c["use_aux"] else None,
                                     return_logits=True)
        s, l, i, m = sort_instances(out["pred_scores"], out["pred_labels"], out["pred_ids"], torch.stack(out["pred_masks"]))
        rs, rl_, ri, rm = sort_instances(c["pred_scores"], c["pred_labels"], c["pred_ids"], c["pred_masks"])
        torch.testing.assert_close(s, rs, rtol=1e-6, atol=1e-7)
        assert torch.equal(l, rl_) and torch.equal(i, ri), name
        _, _, _, lg = sort_instances(out["pred_scores"], out["pred_labels"], out["pred_ids"], out["resized_logits"])
        assert_masks_match(m, rm, lg, tol=1e-5)
        assert out["image_size"] == (Ho, Wo) and out["task"] == "vis"


def test_vis_empty():
    out = pp.inference_video_vis(torch.zeros(0, 6), torch.zeros(0, 3, 4, 4), (14, 15), 14, 15, (16, 16), torch.zeros(0), 5, 10)
    assert out["pred_masks"] == [] and out["pred_scores"] == [] and out["pred_labels"] == [] and out["pred_ids"] == []


def test_vss_matches_reference(golden):
    g = golden("postprocess_vss.pt")
    for name, c in g["cases"].items():
        Ho, Wo = c["output_size"]
        out = pp.inference_video_vss(g["pred_cls"], g["pred_masks"], g["img_size"], Ho, Wo, g["first_resize_size"],
                                     aux_pred_cls=g["aux_cls"] if c["use_aux"] else None, return_scores=True)
        assert_labels_match(out["pred_masks"], c["pred_masks"], out["semseg"], tol=1e-6)


def test_vps_matches_reference(golden):
    g = golden("postprocess_vps.pt")
    for name, c in g["cases"].items():
        Ho, Wo = c["output_size"]
        out = pp.inference_video_vps(g["pred_cls"], g["pred_masks"], g["img_size"], Ho, Wo, g["first_resize_size"], g["pred_id"],
                                     g["num_classes"], g["num_thing_classes"], c["object_mask_threshold"], c["overlap_threshold"],
                                     aux_pred_cls=g["aux_cls"] if c["use_aux"] else None)
        assert out["segments_infos"] == c["segments_infos"], name
        assert [int(i) for i in out["pred_ids"]] == c["pred_ids"], name
        assert out["pred_masks"].dtype == torch.int32
        assert (out["pred_masks"] != c["pred_masks"]).float().mean().item() < 1e-3, name


def test_post_processing_matches_reference(golden):
    g = golden("postprocess_logits.pt")
    logits, ids, aux = pp.post_processing(g["pred_logits"], g["pred_masks"].size(1), aux_logits=g["aux_logits"])
    torch.testing.assert_close(logits, g["dvis_logits"], rtol=0, atol=1e-6)
    torch.testing.assert_close(aux, g["dvis_aux"], rtol=0, atol=1e-6)
    assert torch.equal(ids[0], g["dvis_ids"])
