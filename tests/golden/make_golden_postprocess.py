"""Golden vectors for the video post-processing, produced by calling the UNMODIFIED
reference methods DVIS_Plus_online.inference_video_vis / _vps / _vss / post_processing and MinVIS.post_processing
(P/dvis_Plus/meta_architecture.py:255-301, 758-772, 818-979) on CPU with a stand-in `self` that carries only the
attributes those methods read.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_postprocess.py
"""
import os
import sys
import types
import warnings

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import reference_loader as rl  # noqa: E402

warnings.filterwarnings("ignore")


def blob_logits(Q, T, h, w, gen):
    """Smooth, object-like mask logits: low-frequency field + a little pixel noise (decision boundaries with measure > 0)."""
    coarse = torch.randn(Q, T, max(h // 4, 2), max(w // 4, 2), generator=gen) * 6.0
    fine = F.interpolate(coarse, size=(h, w), mode="bicubic", align_corners=False)
    return fine + 0.3 * torch.randn(Q, T, h, w, generator=gen) - 1.0


def fake_self(num_classes, num_queries, max_num=10, thing=3, object_mask_threshold=0.3, overlap_threshold=0.8):
    return types.SimpleNamespace(
        sem_seg_head=types.SimpleNamespace(num_classes=num_classes), device="cpu", num_queries=num_queries,
        max_num=max_num, object_mask_threshold=object_mask_threshold, overlap_threshold=overlap_threshold,
        metadata=types.SimpleNamespace(thing_dataset_id_to_contiguous_id={i: i for i in range(thing)}))


def save(name, obj):
    path = os.path.join(HERE, name)
    torch.save(obj, path)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


@torch.no_grad()
def main():
    M = rl.load_meta_architecture()
    online = M.DVIS_Plus_online
    gen = torch.Generator().manual_seed(11)
    Q, K, T, h, w = 12, 5, 3, 12, 20
    first = (4 * h, 4 * w)                                   # padded network input (48, 80)
    img = (45, 78)                                           # un-padded size after augmentation
    pred_cls = torch.randn(Q, K + 1, generator=gen) * 2.0
    aux_cls = torch.randn(Q, K + 1, generator=gen) * 2.0
    masks = blob_logits(Q, T, h, w, gen)
    pred_id = torch.arange(Q)

    # --- VIS: up-scaling second resize with aux scores; identity second resize; down-scaling second resize
    cases = {}
    for name, (Ho, Wo), aux, max_num in (("up_aux", (67, 117), aux_cls, 10), ("identity", img, None, 10),
                                         ("down", (30, 52), None, 7), ("one", (45, 78), None, 1)):
        me = fake_self(K, Q, max_num=max_num)
        out = online.inference_video_vis(me, pred_cls.clone(), masks.clone(), img, Ho, Wo, first, pred_id,
                                         aux_pred_cls=None if aux is None else aux.clone())
        cases[name] = dict(output_size=(Ho, Wo), use_aux=aux is not None, max_num=max_num,
                           pred_scores=torch.tensor(out["pred_scores"]), pred_labels=torch.tensor(out["pred_labels"]),
                           pred_ids=torch.tensor(out["pred_ids"]), pred_masks=torch.stack(out["pred_masks"]))
    empty = online.inference_video_vis(fake_self(K, Q), pred_cls[:0], masks[:0], img, 45, 78, first, pred_id[:0])
    assert empty["pred_masks"] == [] and empty["pred_scores"] == []
    save("postprocess_vis.pt", dict(num_classes=K, pred_cls=pred_cls, aux_cls=aux_cls, pred_masks=masks, pred_id=pred_id,
                                    img_size=img, first_resize_size=first, cases=cases))

    # --- VSS
    cases = {}
    for name, (Ho, Wo), aux in (("up_aux", (67, 117), aux_cls), ("identity", img, None)):
        out = online.inference_video_vss(fake_self(K, Q), pred_cls.clone(), masks.clone(), img, Ho, Wo, first, pred_id,
                                         aux_pred_cls=None if aux is None else aux.clone())
        cases[name] = dict(output_size=(Ho, Wo), use_aux=aux is not None, pred_masks=out["pred_masks"])
    save("postprocess_vss.pt", dict(num_classes=K, pred_cls=pred_cls, aux_cls=aux_cls, pred_masks=masks,
                                    img_size=img, first_resize_size=first, cases=cases))

    # --- VPS: confident class logits so that several queries pass object_mask_threshold; stuff classes repeat
    gen = torch.Generator().manual_seed(12)
    vps_cls = torch.randn(Q, K + 1, generator=gen)
    lab = torch.tensor([0, 1, 3, 3, 4, 5, 2, 4, 0, 5, 3, 1])          # 5 = no-object; 3, 4 = stuff (thing = 3)
    vps_cls[torch.arange(Q), lab] += 4.0
    fields = F.interpolate(torch.randn(Q, T, 3, 5, generator=gen), size=(h, w), mode="bicubic", align_corners=False)
    vps_masks = 6.0 * (fields - fields.max(0, keepdim=True).values) + 2.0 + 0.2 * torch.randn(Q, T, h, w, generator=gen)
    cases = {}
    for name, (Ho, Wo), aux, ovl in (("up", (67, 117), None, 0.2), ("identity_aux", img, aux_cls, 0.5),
                                     ("none_kept", img, None, 0.8)):
        me = fake_self(K, Q, thing=3, object_mask_threshold=(2.0 if name == "none_kept" else 0.3), overlap_threshold=ovl)
        out = online.inference_video_vps(me, vps_cls.clone(), vps_masks.clone(), img, Ho, Wo, first, pred_id,
                                         aux_pred_cls=None if aux is None else aux.clone())
        cases[name] = dict(output_size=(Ho, Wo), use_aux=aux is not None, overlap_threshold=ovl,
                           object_mask_threshold=me.object_mask_threshold, pred_masks=out["pred_masks"],
                           segments_infos=out["segments_infos"], pred_ids=[int(i) for i in out["pred_ids"]])
        print(name, out["segments_infos"], cases[name]["pred_ids"])
    save("postprocess_vps.pt", dict(num_classes=K, num_thing_classes=3, pred_cls=vps_cls, aux_cls=aux_cls,
                                    pred_masks=vps_masks, pred_id=pred_id, img_size=img, first_resize_size=first, cases=cases))

    # --- post_processing: offline/online DVIS++ (mean logits + ids) and MinVIS (frame-by-frame Hungarian re-ordering)
    gen = torch.Generator().manual_seed(13)
    logits = torch.randn(1, T, Q, K + 1, generator=gen)
    aux = torch.randn(1, T, Q, K + 1, generator=gen)
    outs = dict(pred_logits=logits.clone(), pred_masks=masks[None].clone())
    o, a = online.post_processing(fake_self(K, Q), outs, aux_logits=aux.clone())
    embds = torch.randn(1, 16, T, Q, generator=gen)
    perm = [torch.randperm(Q, generator=gen) for _ in range(T)]         # the same objects, shuffled per frame + noise
    base = torch.randn(Q, 16, generator=gen)
    for t in range(T):
        embds[0, :, t, :] = (base[perm[t]] + 0.05 * torch.randn(Q, 16, generator=gen)).t()
    mv_in = dict(pred_logits=logits.clone(), pred_masks=masks[None].clone(), pred_embds=embds.clone())
    me = types.SimpleNamespace(match_from_embds=lambda a_, b_: M.MinVIS.match_from_embds(None, a_, b_))
    mv = M.MinVIS.post_processing(me, dict(mv_in))
    save("postprocess_logits.pt", dict(pred_logits=logits, aux_logits=aux, pred_masks=masks[None], pred_embds=embds,
                                       dvis_logits=o["pred_logits"], dvis_ids=o["ids"][0], dvis_aux=a,
                                       minvis_logits=mv["pred_logits"], minvis_masks=mv["pred_masks"]))


if __name__ == "__main__":
    main()
