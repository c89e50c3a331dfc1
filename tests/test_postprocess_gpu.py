"""-m gpu parity tests of the post-processing kernels (csrc/postproc.cu) through the C ABI: against the oracle
(oracle/postprocess_port.py, torch CPU) on identical seeded inputs, against golden outputs of the unmodified reference
methods, and -- at the benchmark's full size -- through size-independent properties.

Tolerance: thresholded / arg-maxed outputs must be IDENTICAL except at pixels where the oracle's own margin is below
1e-4 (fp32 interpolation rounds differently on every implementation; the reference's CPU and CUDA kernels disagree there
too); the number of such pixels is bounded as well.  See tests/postproc_util.py."""
import pytest
import torch

from dvis_plus_b200 import ops
from dvis_plus_b200.modules.postprocess import VideoPostProcessor
from oracle import postprocess_port as pp
from postproc_util import assert_labels_match, assert_masks_match, sort_instances

pytestmark = pytest.mark.gpu
DEV = "cuda"

GEOMS = [  # (h, w), first resize, image size, output size
    ((12, 20), (48, 80), (45, 78), (45, 78)),      # identity second resize: strip kernel, byte-store tail
    ((12, 20), (48, 80), (45, 78), (67, 117)),     # up-scaling second resize
    ((12, 20), (48, 80), (45, 78), (30, 52)),      # down-scaling second resize, width % 4 == 0
    ((12, 20), (48, 80), (48, 80), (48, 80)),      # strip kernel, 8-byte vector stores
    ((7, 9), (28, 36), (25, 33), (25, 33)),        # odd sizes
    ((23, 40), (92, 160), (90, 160), (180, 320)),  # exact 2x second resize
    ((5, 6), (20, 24), (20, 24), (3, 2)),          # output smaller than the logits
    ((12, 20), (48, 80), (45, 78), (200, 301)),    # strong up-scale: intermediate rows reused over many output rows
    ((30, 30), (120, 120), (118, 119), (17, 13)),  # strong down-scale: source rows jump, nothing is reused
    ((46, 80), (184, 320), (180, 320), (100, 177)),  # several row bands per plane, odd width
    ((46, 80), (184, 320), (180, 320), (180, 320)),  # several row bands and strips per plane
]


def test_class_scores_and_topk():
    g = torch.Generator().manual_seed(0)
    for Q, K, max_num, use_aux in ((200, 25, 10, True), (200, 25, 20, False), (100, 40, 100, True), (7, 3, 21, False), (300, 124, 50, True)):
        cls = torch.randn(Q, K + 1, generator=g) * 3
        aux = torch.randn(Q, K + 1, generator=g) * 3 if use_aux else None
        ref = pp.vis_scores(cls, aux)
        full = cls.softmax(-1)
        sc = ops.class_scores(cls.to(DEV), None if aux is None else aux.to(DEV)).cpu()
        torch.testing.assert_close(sc[:, :-1], ref, rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(sc[:, -1], full[:, -1], rtol=1e-5, atol=1e-7)
        s, l, q = (t.cpu() for t in ops.vis_topk(cls.to(DEV), max_num, None if aux is None else aux.to(DEV)))
        rs, ri = ref.flatten().topk(max_num, sorted=True)
        torch.testing.assert_close(s, rs, rtol=1e-5, atol=1e-7)
        assert (s[:-1] >= s[1:]).all()                                  # score descending
        flat = q * K + l
        assert flat.unique().numel() == max_num                         # no entry selected twice
        torch.testing.assert_close(ref.flatten()[flat], rs, rtol=1e-5, atol=1e-7)   # same multiset of scores
        gap = (rs[:-1] - rs[1:]).min().item() if max_num > 1 else 1.0
        if gap > 1e-5:
            assert torch.equal(flat, ri)
    with pytest.raises(RuntimeError, match="out of range"):              # torch.topk raises too (py:831)
        ops.vis_topk(torch.randn(3, 4, device=DEV), 10)


@pytest.mark.parametrize("geom", GEOMS)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_vis_masks_vs_oracle(geom, dtype):
    (h, w), first, img, out = geom
    g = torch.Generator().manual_seed(h * 100 + w)
    masks = (torch.randn(6, 3, h, w, generator=g) * 3).to(dtype)
    sel = torch.tensor([4, 0, 4, 5], dtype=torch.int64)
    ours = ops.vis_masks(masks.to(DEV), sel.to(DEV), first, img, out)
    assert ours.dtype == torch.bool and ours.shape == (4, 3, *out)
    ref = pp.resize_chain(masks[sel].float(), img, out[0], out[1], first)
    assert_masks_match(ours, ref > 0, ref)
    # all queries, frame-major storage (T, Q, h, w) viewed as (Q, T, h, w): the layout the mask GEMM produces
    fm = masks.to(DEV).transpose(0, 1).contiguous().transpose(0, 1)
    ours2 = ops.vis_masks(fm, None, first, img, out)
    ref2 = pp.resize_chain(masks.float(), img, out[0], out[1], first)
    assert_masks_match(ours2, ref2 > 0, ref2)


def test_vis_module_vs_reference_golden(golden):
    g = golden("postprocess_vis.pt")
    for name, c in g["cases"].items():
        Ho, Wo = c["output_size"]
        post = VideoPostProcessor(g["num_classes"], num_queries=12, max_num=c["max_num"])
        aux = g["aux_cls"].to(DEV) if c["use_aux"] else None
        out = post.inference_video_task(g["pred_cls"].to(DEV), g["pred_masks"].to(DEV), g["img_size"], Ho, Wo,
                                        g["first_resize_size"], g["pred_id"], aux_pred_cls=aux)
        assert all(m.device.type == "cpu" and m.dtype == torch.bool for m in out["pred_masks"])
        s, l, i, m = sort_instances(out["pred_scores"], out["pred_labels"], out["pred_ids"], torch.stack(out["pred_masks"]))
        rs, rl, ri, rm = sort_instances(c["pred_scores"], c["pred_labels"], c["pred_ids"], c["pred_masks"])
        torch.testing.assert_close(s, rs, rtol=1e-5, atol=1e-7)
        assert torch.equal(l, rl) and torch.equal(i, ri), name
        o = pp.inference_video_vis(g["pred_cls"], g["pred_masks"], g["img_size"], Ho, Wo, g["first_resize_size"], g["pred_id"],
                                   g["num_classes"], c["max_num"], aux_pred_cls=g["aux_cls"] if c["use_aux"] else None, return_logits=True)
        _, _, _, lg = sort_instances(o["pred_scores"], o["pred_labels"], o["pred_ids"], o["resized_logits"])
        assert_masks_match(m, rm, lg)
    empty = VideoPostProcessor(5).inference_video_vis(g["pred_cls"][:0].to(DEV), g["pred_masks"][:0].to(DEV), g["img_size"], 45, 78,
                                                      g["first_resize_size"], g["pred_id"][:0])
    assert empty["pred_masks"] == [] and empty["pred_scores"] == []


def _blob_logits(Q, T, h, w, gen):
    coarse = torch.randn(Q, T, max(h // 8, 2), max(w // 8, 2), generator=gen) * 6.0
    fine = torch.nn.functional.interpolate(coarse, size=(h, w), mode="bicubic", align_corners=False)
    return fine + 0.3 * torch.randn(Q, T, h, w, generator=gen) - 1.0


@pytest.mark.parametrize("case", ["720p_identity", "480p_to_720p"])
def test_vis_masks_full_size(case):
    """BASELINE metric size (720p, T=16, 10 of the queries kept) and a real two-resize geometry (480p inference of a 720p
    video).  The oracle (F.interpolate on the host) checks a few frames; size-independent properties cover all of them."""
    gen = torch.Generator().manual_seed(5)
    if case == "720p_identity":
        Q, T, (h, w), first, img, out = 24, 16, (184, 320), (736, 1280), (720, 1280), (720, 1280)
    else:
        Q, T, (h, w), first, img, out = 24, 8, (120, 216), (480, 864), (480, 854), (720, 1280)
    masks = _blob_logits(Q, T, h, w, gen)
    sel = torch.randperm(Q, generator=gen)[:10]
    d, dsel = masks.to(DEV), sel.to(DEV)
    ours = ops.vis_masks(d, dsel, first, img, out)
    assert ours.shape == (10, T, *out) and 0.02 < ours.float().mean().item() < 0.98     # a non-trivial pattern
    frames = [0, T // 2, T - 1]                                            # oracle on 3 frames x 4 instances
    ref = pp.resize_chain(masks[sel[:4]][:, frames], img, out[0], out[1], first)
    assert_masks_match(ours[:4][:, frames], ref > 0, ref, max_boundary_frac=1e-4)
    # bf16 logits (what the bf16 mask GEMM emits): same decision as the oracle on the bf16-rounded values
    ours_bf = ops.vis_masks(d.bfloat16(), dsel, first, img, out)
    ref_bf = pp.resize_chain(masks[sel[:4]][:, frames].bfloat16().float(), img, out[0], out[1], first)
    assert_masks_match(ours_bf[:4][:, frames], ref_bf > 0, ref_bf, max_boundary_frac=1e-4)
    # properties over the full size: negation flips every pixel whose value is not exactly 0; a positive shift only adds
    # pixels; constant logits give constant masks; selection commutes with the kernel; frames / instances are independent
    neg = ops.vis_masks(-d, dsel, first, img, out)
    assert (neg & ours).sum().item() == 0 and (~(neg | ours)).float().mean().item() < 1e-4
    shifted = ops.vis_masks(d + 0.5, dsel, first, img, out)
    assert (ours & ~shifted).sum().item() == 0 and shifted.sum() > ours.sum()
    assert ops.vis_masks(torch.full_like(d[:2], 0.25), None, first, img, out).all()
    assert not ops.vis_masks(torch.full_like(d[:2], -0.25), None, first, img, out).any()
    assert torch.equal(ops.vis_masks(d, None, first, img, out)[dsel], ours)
    assert torch.equal(ops.vis_masks(d[:, 3:5], dsel[2:7], first, img, out), ours[2:7, 3:5])


@pytest.mark.parametrize("geom", GEOMS[:5])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_vps_argmax_vs_oracle(geom, dtype):
    (h, w), first, img, out = geom
    g = torch.Generator().manual_seed(7 + h)
    masks = (torch.randn(9, 2, h, w, generator=g) * 3).to(dtype)
    keep_idx = torch.tensor([1, 3, 4, 8], dtype=torch.int64)
    keep_score = torch.tensor([0.9, 0.5, 0.7, 0.95])
    win, areas = ops.vps_argmax(masks.to(DEV), keep_idx.to(DEV), keep_score.to(DEV), first, img, out)
    win, areas = win.cpu(), areas.cpu()
    cur = pp.resize_chain(masks[keep_idx].float(), img, out[0], out[1], first, sigmoid=True)
    prob = keep_score.view(-1, 1, 1, 1) * cur
    ref_ids = prob.argmax(0)
    ids = torch.where(win >= 0, win, ~win).long()
    assert_labels_match(ids, ref_ids, prob, tol=1e-4)
    same = ids == ref_ids
    margin = (cur.gather(0, ref_ids[None])[0] - 0.5).abs()
    assert ((win >= 0) == (cur.gather(0, ref_ids[None])[0] >= 0.5))[same & (margin > 1e-4)].all()
    n = keep_idx.numel()
    ref_areas = torch.stack([torch.stack([(ref_ids == k).sum() for k in range(n)]),
                             torch.stack([(cur[k] >= 0.5).sum() for k in range(n)]),
                             torch.stack([((ref_ids == k) & (cur[k] >= 0.5)).sum() for k in range(n)])])
    assert (areas - ref_areas).abs().max().item() <= 3, (areas, ref_areas)
    assert areas[0].sum().item() == ref_ids.numel()
    # painting: segment ids through the winner map
    seg = torch.tensor([5, 0, 7, 7], dtype=torch.int32)
    pan = ops.vps_paint(win.to(DEV), seg.to(DEV)).cpu()
    assert torch.equal(pan, torch.where(win >= 0, seg[win.clamp(min=0).long()], torch.zeros_like(win)))


def test_vps_module_vs_reference_golden(golden):
    g = golden("postprocess_vps.pt")
    for name, c in g["cases"].items():
        Ho, Wo = c["output_size"]
        post = VideoPostProcessor(g["num_classes"], object_mask_threshold=c["object_mask_threshold"],
                                  overlap_threshold=c["overlap_threshold"], num_thing_classes=g["num_thing_classes"], task="vps")
        out = post.inference_video_task(g["pred_cls"].to(DEV), g["pred_masks"].to(DEV), g["img_size"], Ho, Wo,
                                        g["first_resize_size"], g["pred_id"], aux_pred_cls=g["aux_cls"].to(DEV) if c["use_aux"] else None)
        assert out["segments_infos"] == c["segments_infos"], name
        assert [int(i) for i in out["pred_ids"]] == c["pred_ids"], name
        assert out["pred_masks"].dtype == torch.int32 and out["pred_masks"].device.type == "cpu"
        assert (out["pred_masks"] != c["pred_masks"]).float().mean().item() < 1e-3, name


@pytest.mark.parametrize("geom", GEOMS[:4])
@pytest.mark.parametrize("K", [5, 19, 124])
def test_vss_argmax_vs_oracle(geom, K):
    (h, w), first, img, out = geom
    g = torch.Generator().manual_seed(3 + K)
    Q = 10 if K < 100 else 100
    masks = torch.randn(Q, 2, h, w, generator=g) * 3
    cls = torch.randn(Q, K + 1, generator=g) * 2
    scores = ops.class_scores(cls.to(DEV))
    ours = ops.vss_argmax(masks.to(DEV), scores[:, :-1], first, img, out)
    ref = pp.inference_video_vss(cls, masks, img, out[0], out[1], first, return_scores=True)
    assert ours.dtype == torch.int64
    assert_labels_match(ours, ref["pred_masks"], ref["semseg"], tol=1e-4)
    ours_bf = ops.vss_argmax(masks.to(DEV).bfloat16(), scores[:, :-1], first, img, out)
    ref_bf = pp.inference_video_vss(cls, masks.bfloat16().float(), img, out[0], out[1], first, return_scores=True)
    assert_labels_match(ours_bf, ref_bf["pred_masks"], ref_bf["semseg"], tol=1e-4)


def test_vss_module_vs_reference_golden(golden):
    g = golden("postprocess_vss.pt")
    for name, c in g["cases"].items():
        Ho, Wo = c["output_size"]
        aux = g["aux_cls"] if c["use_aux"] else None
        out = VideoPostProcessor(g["num_classes"], task="vss").inference_video_task(
            g["pred_cls"].to(DEV), g["pred_masks"].to(DEV), g["img_size"], Ho, Wo, g["first_resize_size"], None,
            aux_pred_cls=None if aux is None else aux.to(DEV))
        ref = pp.inference_video_vss(g["pred_cls"], g["pred_masks"], g["img_size"], Ho, Wo, g["first_resize_size"], aux_pred_cls=aux,
                                     return_scores=True)
        assert_labels_match(out["pred_masks"], c["pred_masks"], ref["semseg"], tol=1e-4)


def test_post_processing_vs_reference_golden(golden):
    g = golden("postprocess_logits.pt")
    post = VideoPostProcessor(5)
    outs, aux = post.post_processing(dict(pred_logits=g["pred_logits"].to(DEV), pred_masks=g["pred_masks"].to(DEV)),
                                     aux_logits=g["aux_logits"].to(DEV))
    torch.testing.assert_close(outs["pred_logits"].cpu(), g["dvis_logits"], rtol=0, atol=1e-6)
    torch.testing.assert_close(aux.cpu(), g["dvis_aux"], rtol=0, atol=1e-6)
    mv = post.post_processing_minvis(dict(pred_logits=g["pred_logits"].to(DEV), pred_masks=g["pred_masks"].to(DEV),
                                          pred_embds=g["pred_embds"].to(DEV)))
    torch.testing.assert_close(mv["pred_logits"].cpu(), g["minvis_logits"], rtol=1e-5, atol=1e-6)
    assert torch.equal(mv["pred_masks"].cpu(), g["minvis_masks"])        # GPU Hungarian == SciPy on the reference's input


@torch.no_grad()
@pytest.mark.parametrize("out_size", [(60, 90), (90, 135)])
def test_pipeline_vis_from_block_equals_postprocessing_all_masks(out_size):
    """Selecting the instances BEFORE the final mask GEMM (OfflineClipRunner.vis_from_block) gives the same result as the
    reference order: all Q masks -> post_processing -> inference_video_vis."""
    from dvis_plus_b200 import _lib, modules as M
    from dvis_plus_b200.modules.precision import precision
    from dvis_plus_b200.pipeline import OfflineClipRunner
    T, Q, C, K, H, W = 4, 24, 64, 5, 16, 24
    torch.manual_seed(0)
    trk = M.ReferringTracker_noiser(hidden_channel=C, feedforward_channel=128, num_head=8, decoder_layer_num=2, mask_dim=64,
                                    class_num=K, noise_mode="none").eval().to(DEV)
    rfn = M.TemporalRefiner(hidden_channel=C, feedforward_channel=128, num_head=8, decoder_layer_num=2, mask_dim=64, class_num=K,
                            windows=2).eval().to(DEV)
    g = torch.Generator().manual_seed(1)
    seg = dict(pred_embds=torch.randn(1, C, T, Q, generator=g).to(DEV), pred_embds_without_norm=torch.randn(1, C, T, Q, generator=g).to(DEV),
               pred_logits=torch.randn(1, T, Q, K + 1, generator=g).to(DEV))
    mf = torch.randn(T, 64, H, W, generator=g).to(DEV).to(torch.bfloat16, memory_format=torch.channels_last)
    runner = OfflineClipRunner(None, None, trk, rfn)
    post = VideoPostProcessor(K, num_queries=Q, max_num=10)
    img = (60, 90)
    n0 = _lib.launch_count
    with precision("bf16"):
        block = runner.pack_queries(seg)
        fused = runner.vis_from_block(block, mf, C, post, img, out_size)
        full = runner.temporal_from_block(block, mf, C)
    assert _lib.launch_count - n0 >= 4, "libdvis_b200 kernels did not run"
    outs, aux = post.post_processing(dict(pred_logits=full["pred_logits"], pred_masks=full["pred_masks"]),
                                     aux_logits=full["online_pred_logits"])
    ref = post.inference_video_vis(outs["pred_logits"][0], outs["pred_masks"][0], img, *out_size, (4 * H, 4 * W), outs["ids"][0],
                                   aux_pred_cls=aux)
    torch.testing.assert_close(fused["pred_scores"].cpu(), torch.tensor(ref["pred_scores"]), rtol=1e-5, atol=1e-7)
    assert fused["pred_labels"].tolist() == ref["pred_labels"] and fused["pred_ids"].tolist() == ref["pred_ids"]
    assert fused["pred_masks"].shape == (10, T, *out_size) and fused["pred_masks"].dtype == torch.bool
    assert (fused["pred_masks"].cpu() != torch.stack(ref["pred_masks"])).float().mean().item() < 1e-3
    # and the masks agree with the oracle's resize chain applied to the pipeline's own stride-4 logits
    low = outs["pred_masks"][0][torch.tensor(ref["pred_ids"], device=DEV)].float().cpu()
    chain = pp.resize_chain(low, img, out_size[0], out_size[1], (4 * H, 4 * W))
    assert_masks_match(fused["pred_masks"], chain > 0, chain, tol=1e-3, max_boundary_frac=5e-3)


@pytest.mark.parametrize("geom", [GEOMS[0], GEOMS[1], GEOMS[3], GEOMS[5], GEOMS[8], GEOMS[10]])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_vis_masks_packed(geom, dtype):
    """One bit per pixel (dvis_vis_masks_packed): exactly the byte kernel's masks, packed little-endian with zero padding."""
    import numpy as np
    (h, w), first, img, out = geom
    g = torch.Generator().manual_seed(h * 10 + w)
    masks = (torch.randn(5, 3, h, w, generator=g) * 3).to(dtype).to(DEV)
    sel = torch.tensor([3, 1, 4], dtype=torch.int64, device=DEV)
    ref = ops.vis_masks(masks, sel, first, img, out).cpu()
    packed = ops.vis_masks(masks, sel, first, img, out, packed=True).cpu()
    assert packed.dtype == torch.uint8 and packed.shape == (3, 3, out[0], (out[1] + 7) // 8)
    assert np.array_equal(packed.numpy(), np.packbits(ref.numpy(), axis=-1, bitorder="little"))
    assert torch.equal(ops.unpack_masks(packed, out[1]), ref)


def test_vis_module_packed_transfer_equals_plain(golden):
    g = golden("postprocess_vis.pt")
    c = g["cases"]["up_aux"]
    Ho, Wo = c["output_size"]
    outs = []
    for packed in (True, False):
        post = VideoPostProcessor(g["num_classes"], num_queries=12, max_num=c["max_num"])
        post.packed_transfer = packed
        outs.append(post.inference_video_vis(g["pred_cls"].to(DEV), g["pred_masks"].to(DEV), g["img_size"], Ho, Wo,
                                             g["first_resize_size"], g["pred_id"], aux_pred_cls=g["aux_cls"].to(DEV)))
    assert outs[0]["pred_scores"] == outs[1]["pred_scores"] and outs[0]["pred_ids"] == outs[1]["pred_ids"]
    assert torch.equal(torch.stack(outs[0]["pred_masks"]), torch.stack(outs[1]["pred_masks"]))


@pytest.mark.parametrize("geom", [GEOMS[0], GEOMS[1], GEOMS[3], GEOMS[5], GEOMS[8], GEOMS[10]])
def test_vis_masks_tiled_variant_is_bit_identical(geom, monkeypatch):
    """DVIS_VIS_MASKS_TILED=1 (source window of the CTA's tile staged in shared memory): same bytes / bits as the default."""
    (h, w), first, img, out = geom
    g = torch.Generator().manual_seed(h * 7 + w)
    masks = (torch.randn(5, 3, h, w, generator=g) * 3).to(DEV)
    sel = torch.tensor([3, 0, 4], dtype=torch.int64, device=DEV)
    ref, ref_p = ops.vis_masks(masks, sel, first, img, out), ops.vis_masks(masks, sel, first, img, out, packed=True)
    ref_b = ops.vis_masks(masks.bfloat16(), sel, first, img, out)
    monkeypatch.setenv("DVIS_VIS_MASKS_TILED", "1")
    assert torch.equal(ops.vis_masks(masks, sel, first, img, out), ref)
    assert torch.equal(ops.vis_masks(masks, sel, first, img, out, packed=True), ref_p)
    assert torch.equal(ops.vis_masks(masks.bfloat16(), sel, first, img, out), ref_b)
