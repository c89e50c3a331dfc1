"""Host logic of dvis_plus_b200.modules.postprocess.VideoPostProcessor against the reference's golden vectors, on CPU.

The C-ABI entry points the module calls cannot run without a GPU, so this test swaps them for TEST DOUBLES: the host
compilation of the very same per-pixel code (tests/hostcore), torch for softmax / top-k and SciPy for the assignment.
What is checked here is everything around the kernels -- selection, id bookkeeping, the vps segment filter and stuff
merging, result dictionaries, the MinVIS index algebra.  The kernels themselves are checked by test_postprocess_gpu.py."""
import os
import sys

import pytest
import torch

from dvis_plus_b200 import ops
from dvis_plus_b200.modules.postprocess import VideoPostProcessor
from oracle import postprocess_port as pp
from postproc_util import assert_labels_match, assert_masks_match, sort_instances

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostcore"))
import hostcore_binding as hc  # noqa: E402


@pytest.fixture
def doubles(monkeypatch):
    def class_scores(pred_cls, aux=None):
        s = pred_cls.float().softmax(-1)
        if aux is not None:
            s = s.clone()
            s[:, :-1] = torch.maximum(s[:, :-1], aux.float().softmax(-1)[:, :-1])
        return s

    def vis_topk(pred_cls, max_num, aux=None):
        s = class_scores(pred_cls, aux)[:, :-1]
        K = s.shape[1]
        v, i = s.flatten().topk(max_num, sorted=True)
        return v, i % K, i // K

    def lap_chain(cost):
        from scipy.optimize import linear_sum_assignment
        idx, sig, out = None, [], []
        for t in range(cost.shape[0]):
            s = torch.as_tensor(linear_sum_assignment(cost[t].numpy())[1])
            idx = s if idx is None else s[idx]
            sig.append(s)
            out.append(idx)
        return torch.stack(sig), torch.stack(out)

    monkeypatch.setattr(ops, "class_scores", class_scores)
    monkeypatch.setattr(ops, "vis_topk", vis_topk)
    def vis_masks(m, sel, first, img, out, packed=False):
        masks = hc.vis_masks(m, sel, first, img, out)
        if not packed:
            return masks
        import numpy as np
        return torch.from_numpy(np.packbits(masks.numpy(), axis=-1, bitorder="little"))

    monkeypatch.setattr(ops, "vis_masks", vis_masks)
    monkeypatch.setattr(ops, "vps_argmax", lambda m, ki, ks, first, img, out: hc.vps_argmax(m, ki, ks, first, img, out))
    monkeypatch.setattr(ops, "vps_paint", lambda win, seg: torch.where(win >= 0, seg[win.clamp(min=0).long()], torch.zeros_like(win)))
    monkeypatch.setattr(ops, "vss_argmax", lambda m, mc, first, img, out: hc.vss_argmax(m, mc, first, img, out))
    monkeypatch.setattr(ops, "lap_chain", lap_chain)


def test_vis_dict_matches_reference(golden, doubles):
    g = golden("postprocess_vis.pt")
    for name, c in g["cases"].items():
        Ho, Wo = c["output_size"]
        post = VideoPostProcessor(g["num_classes"], num_queries=12, max_num=c["max_num"])
        out = post.inference_video_task(g["pred_cls"], g["pred_masks"], g["img_size"], Ho, Wo, g["first_resize_size"], g["pred_id"],
                                        aux_pred_cls=g["aux_cls"] if c["use_aux"] else None)
        assert set(out) == {"image_size", "pred_scores", "pred_labels", "pred_masks", "pred_ids", "task"}
        assert out["image_size"] == (Ho, Wo) and out["task"] == "vis" and len(out["pred_masks"]) == c["max_num"]
        assert all(m.dtype == torch.bool and m.shape == (3, Ho, Wo) for m in out["pred_masks"])
        s, l, i, m = sort_instances(out["pred_scores"], out["pred_labels"], out["pred_ids"], torch.stack(out["pred_masks"]))
        rs, rl, ri, rm = sort_instances(c["pred_scores"], c["pred_labels"], c["pred_ids"], c["pred_masks"])
        torch.testing.assert_close(s, rs, rtol=1e-6, atol=1e-7)
        assert torch.equal(l, rl) and torch.equal(i, ri), name
        o = pp.inference_video_vis(g["pred_cls"], g["pred_masks"], g["img_size"], Ho, Wo, g["first_resize_size"], g["pred_id"],
                                   g["num_classes"], c["max_num"], aux_pred_cls=g["aux_cls"] if c["use_aux"] else None, return_logits=True)
        _, _, _, lg = sort_instances(o["pred_scores"], o["pred_labels"], o["pred_ids"], o["resized_logits"])
        assert_masks_match(m, rm, lg, tol=2e-5)
    empty = VideoPostProcessor(5).inference_video_vis(g["pred_cls"][:0], g["pred_masks"][:0], g["img_size"], 45, 78,
                                                      g["first_resize_size"], g["pred_id"][:0])
    assert empty["pred_masks"] == [] and empty["pred_scores"] == [] and empty["pred_ids"] == [] and empty["task"] == "vis"


def test_vps_dict_matches_reference(golden, doubles):
    g = golden("postprocess_vps.pt")
    for name, c in g["cases"].items():
        Ho, Wo = c["output_size"]
        post = VideoPostProcessor(g["num_classes"], object_mask_threshold=c["object_mask_threshold"],
                                  overlap_threshold=c["overlap_threshold"], num_thing_classes=g["num_thing_classes"], task="vps")
        out = post.inference_video_task(g["pred_cls"], g["pred_masks"], g["img_size"], Ho, Wo, g["first_resize_size"], g["pred_id"],
                                        aux_pred_cls=g["aux_cls"] if c["use_aux"] else None)
        assert set(out) == {"image_size", "pred_masks", "segments_infos", "pred_ids", "task"} and out["task"] == "vps"
        assert out["segments_infos"] == c["segments_infos"], name
        assert [int(i) for i in out["pred_ids"]] == c["pred_ids"], name
        assert out["pred_masks"].dtype == torch.int32 and out["pred_masks"].shape == c["pred_masks"].shape
        assert (out["pred_masks"] != c["pred_masks"]).float().mean().item() < 1e-3, name


def test_vss_dict_matches_reference(golden, doubles):
    g = golden("postprocess_vss.pt")
    for name, c in g["cases"].items():
        Ho, Wo = c["output_size"]
        post = VideoPostProcessor(g["num_classes"], task="vss")
        aux = g["aux_cls"] if c["use_aux"] else None
        out = post.inference_video_task(g["pred_cls"], g["pred_masks"], g["img_size"], Ho, Wo, g["first_resize_size"], None, aux_pred_cls=aux)
        ref = pp.inference_video_vss(g["pred_cls"], g["pred_masks"], g["img_size"], Ho, Wo, g["first_resize_size"], aux_pred_cls=aux,
                                     return_scores=True)
        assert out["pred_masks"].dtype == torch.int64 and out["task"] == "vss"
        assert_labels_match(out["pred_masks"], c["pred_masks"], ref["semseg"], tol=1e-5)


def test_post_processing_matches_reference(golden, doubles):
    g = golden("postprocess_logits.pt")
    post = VideoPostProcessor(5)
    outs, aux = post.post_processing(dict(pred_logits=g["pred_logits"].clone(), pred_masks=g["pred_masks"]), aux_logits=g["aux_logits"])
    torch.testing.assert_close(outs["pred_logits"], g["dvis_logits"], rtol=0, atol=1e-6)
    torch.testing.assert_close(aux, g["dvis_aux"], rtol=0, atol=1e-6)
    assert torch.equal(outs["ids"][0], g["dvis_ids"])
    mv = post.post_processing_minvis(dict(pred_logits=g["pred_logits"].clone(), pred_masks=g["pred_masks"].clone(),
                                          pred_embds=g["pred_embds"].clone()))
    torch.testing.assert_close(mv["pred_logits"], g["minvis_logits"], rtol=1e-6, atol=1e-6)
    assert torch.equal(mv["pred_masks"], g["minvis_masks"])


def test_cpu_tensors_raise():
    post = VideoPostProcessor(5)
    with pytest.raises(RuntimeError, match="no CPU path"):
        post.inference_video_vis(torch.randn(4, 6), torch.randn(4, 2, 8, 8), (30, 30), 30, 30, (32, 32), torch.arange(4))


def test_pipeline_vis_from_block_equals_postprocessing_all_masks(doubles):
    """OfflineClipRunner.vis_from_block selects the instances BEFORE the final mask GEMM; the result must equal running the
    reference order (all Q masks -> post_processing -> inference_video_vis) on the same tracker / refiner outputs."""
    from dvis_plus_b200 import modules as M
    from dvis_plus_b200.pipeline import OfflineClipRunner
    T, Q, C, K, H, W = 4, 10, 64, 5, 8, 12
    torch.manual_seed(0)
    trk = M.ReferringTracker_noiser(hidden_channel=C, feedforward_channel=128, num_head=8, decoder_layer_num=2, mask_dim=32,
                                    class_num=K, noise_mode="none").eval()
    rfn = M.TemporalRefiner(hidden_channel=C, feedforward_channel=128, num_head=8, decoder_layer_num=2, mask_dim=32, class_num=K,
                            windows=2).eval()
    g = torch.Generator().manual_seed(1)
    seg = dict(pred_embds=torch.randn(1, C, T, Q, generator=g), pred_embds_without_norm=torch.randn(1, C, T, Q, generator=g),
               pred_logits=torch.randn(1, T, Q, K + 1, generator=g))
    mf = torch.randn(T, 32, H, W, generator=g)
    runner = OfflineClipRunner(None, None, trk, rfn)
    post = VideoPostProcessor(K, num_queries=Q, max_num=6)
    img, out_size = (30, 45), (41, 60)
    block = runner.pack_queries(seg)
    fused = runner.vis_from_block(block, mf, C, post, img, out_size)
    full = runner.temporal_from_block(block, mf, C)
    outs, aux = post.post_processing(dict(pred_logits=full["pred_logits"], pred_masks=full["pred_masks"]),
                                     aux_logits=full["online_pred_logits"])
    ref = post.inference_video_vis(outs["pred_logits"][0], outs["pred_masks"][0], img, *out_size, (4 * H, 4 * W), outs["ids"][0],
                                   aux_pred_cls=aux)
    torch.testing.assert_close(fused["pred_scores"], torch.tensor(ref["pred_scores"]), rtol=1e-6, atol=1e-7)
    assert fused["pred_labels"].tolist() == ref["pred_labels"] and fused["pred_ids"].tolist() == ref["pred_ids"]
    assert fused["pred_masks"].shape == (6, T, *out_size)
    assert (fused["pred_masks"] != torch.stack(ref["pred_masks"])).float().mean().item() < 1e-4


def test_vis_packed_transfer_gives_the_same_result(golden, doubles):
    g = golden("postprocess_vis.pt")
    c = g["cases"]["up_aux"]
    Ho, Wo = c["output_size"]
    outs = []
    for packed in (True, False):
        post = VideoPostProcessor(g["num_classes"], num_queries=12, max_num=c["max_num"])
        post.packed_transfer = packed
        outs.append(post.inference_video_vis(g["pred_cls"], g["pred_masks"], g["img_size"], Ho, Wo, g["first_resize_size"], g["pred_id"],
                                             aux_pred_cls=g["aux_cls"]))
    assert outs[0]["pred_scores"] == outs[1]["pred_scores"] and outs[0]["pred_ids"] == outs[1]["pred_ids"]
    assert all(m.dtype == torch.bool and m.shape == (3, Ho, Wo) for m in outs[0]["pred_masks"])
    assert torch.equal(torch.stack(outs[0]["pred_masks"]), torch.stack(outs[1]["pred_masks"]))
