"""DVIS-DAQ drop-ins against golden outputs of the unmodified reference modules
(tests/golden/daq_*.pt).  CPU tests check the host logic (per-instance bookkeeping, dynamic query count, resume);
the GPU test runs the same fixture through the kernels."""
import random

import pytest
import torch

from dvis_plus_b200 import modules as M


def close(a, b, tol=2e-4):
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a.double().cpu() - b.double()).abs().max().item()
    assert err <= tol * max(1.0, b.abs().max().item()), err


def run_cutter(g, device="cpu"):
    C, fQ = g["frame_embeds"].shape[1], g["frame_embeds"].shape[3]
    cut = M.VideoInstanceCutter(hidden_dim=C, feedforward_dim=128, num_head=8, decoder_layer_num=2, mask_dim=C, num_classes=5,
                                num_new_ins=fQ, inference_select_threshold=0.1, kick_out_frame_num=2, num_slots=3,
                                keep_threshold=0.01, ovis_infer=True).eval()
    assert not any(cut.load_state_dict(g["state_dict"]))
    cut = cut.to(device)
    emb = torch.nn.Embedding(fQ, C).to(device)
    emb.weight.data.copy_(g["seg_query_feat"])
    d = lambda t: t.to(device)
    info = lambda a, b: {"seg_query_feat": emb, "valid": [[d(v)] for v in g["valid"][a:b]], "pred_masks": [[d(p)] for p in g["pred_masks"][a:b]]}
    random.seed(g["seed"])
    fe, mf = d(g["frame_embeds"]), d(g["mask_features"])
    cut.inference(fe[:, :, :3], mf[:, :3], info(0, 3), 0, resume=False, to_store="cpu")
    cut.inference(fe[:, :, 3:], mf[:, 3:], info(3, 4), 3, resume=True, to_store="cpu")
    return cut


def check_cutter(cut, g, tol):
    assert len(cut.memory_seq_ids) == len(g["seqs"])
    for sid, ref in zip(cut.memory_seq_ids, g["seqs"]):
        s = cut.video_ins_hub[sid]
        assert s.sT == ref["sT"] and s.dead == ref["dead"] and list(s.appearance) == ref["appearance"]
        close(torch.stack(s.embeds), ref["embeds"], tol)
        close(torch.stack(s.pred_logits), ref["pred_logits"], tol)
        close(torch.stack(s.pred_masks), ref["pred_masks"], tol)
        close(s.similarity_guided_pos_embed, ref["pos"], tol)
    close(cut.track_queries, g["track_queries"], tol)
    close(cut.track_embeds, g["track_embeds"], tol)


@torch.no_grad()
def test_daq_tracker_cpu_matches_reference(golden):
    g = golden("daq_tracker_small.pt")
    check_cutter(run_cutter(g), g, 2e-4)


@torch.no_grad()
def test_daq_slot_layer_and_refiner_cpu(golden):
    g = golden("daq_slot_layer.pt")
    sl = M.SlotCrossAttentionLayer(d_model=64, nhead=8).eval()
    assert not any(sl.load_state_dict(g["state_dict"]))
    close(sl(g["tgt"], g["memory"], query_pos=g["query_pos"], slot_query=g["slot_query"]), g["out"])
    g = golden("daq_refiner_small.pt")
    rf = M.DAQTemporalRefiner(hidden_channel=64, feedforward_channel=128, num_head=8, decoder_layer_num=2, mask_dim=64,
                              class_num=5, windows=3, use_local_attn=False).eval()
    assert not any(rf.load_state_dict(g["state_dict"]))
    o = rf(g["instance_embeds"], None, g["frame_embeds"], g["mask_features"], None)
    close(o["pred_logits"], g["pred_logits"])
    close(o["pred_masks"], g["pred_masks"])
    close(o["pred_embds"], g["pred_embds"])


@pytest.mark.gpu
@torch.no_grad()
def test_daq_tracker_gpu_matches_reference(golden):
    """Same fixture on the GPU: mask einsum on the tcgen05 GEMM (run-time query count), mask-pooled embeddings as a GEMM.
    fp32 GEMM mode; masks go through bf16 operands -> 2e-2 of scale; the valid / invalid decisions must be identical."""
    from dvis_plus_b200 import _lib
    from dvis_plus_b200.modules.precision import precision
    g = golden("daq_tracker_small.pt")
    n0 = _lib.launch_count
    with precision("fp32"):
        cut = run_cutter(g, "cuda")
    assert _lib.launch_count > n0
    check_cutter(cut, g, 2e-2)


@pytest.mark.gpu
@torch.no_grad()
def test_daq_refiner_gpu(golden):
    from dvis_plus_b200.modules.precision import precision
    g = golden("daq_refiner_small.pt")
    rf = M.DAQTemporalRefiner(hidden_channel=64, feedforward_channel=128, num_head=8, decoder_layer_num=2, mask_dim=64,
                              class_num=5, windows=3, use_local_attn=False).eval().cuda()
    rf.load_state_dict(g["state_dict"])
    with precision("fp32"):
        o = rf(g["instance_embeds"].cuda(), None, g["frame_embeds"].cuda(), g["mask_features"].cuda(), None)
    close(o["pred_logits"], g["pred_logits"], 1e-3)
    close(o["pred_embds"], g["pred_embds"], 1e-3)
    close(o["pred_masks"], g["pred_masks"], 1e-2)


@torch.no_grad()
def test_daq_online_window_loop_matches_reference(golden):
    """pipeline.DAQOnlineRunner against the unmodified DVIS_DAQ_online.run_window_inference (fixture:
    tests/golden/make_golden_daq_runner.py): same surviving instances, ids, averaged logits, per-frame masks / logits and
    padding masks; dead sequences are dropped from the hub."""
    import types
    from dvis_plus_b200.pipeline import DAQOnlineRunner
    g = golden("daq_runner_small.pt")
    seg, K = g["seg"], g["num_classes"]
    C, fQ = seg["pred_embds"].shape[1], seg["pred_embds"].shape[3]
    cut = M.VideoInstanceCutter(hidden_dim=C, feedforward_dim=128, num_head=8, decoder_layer_num=2, mask_dim=C, num_classes=K,
                                num_new_ins=fQ, inference_select_threshold=0.1, kick_out_frame_num=2, num_slots=3,
                                keep_threshold=0.01, ovis_infer=True).eval()
    assert not any(cut.load_state_dict(g["state_dict"]))
    emb = torch.nn.Embedding(fQ, C)
    emb.weight.data.copy_(g["query_feat"])
    predictor = types.SimpleNamespace(query_feat=emb, query_embed=emb)

    def segment(window):                          # the fixture's precomputed segmenter outputs, sliced by frame index
        idx = window["frames"]
        return {"pred_embds": seg["pred_embds"][:, :, idx], "mask_features": seg["mask_features"][idx],
                "pred_logits": seg["pred_logits"][:, idx], "pred_masks": seg["pred_masks"][:, :, idx]}

    random.seed(g["seed"])
    runner = DAQOnlineRunner(None, predictor, cut, K, g["aux_inference_select_thr"], g["noise_frame_num"], segment=segment,
                             to_store="cpu")
    T = seg["pred_embds"].shape[2]
    out, ref = runner({"frames": torch.arange(T)}), g["out"]
    assert out["shape"] == ref["shape"]
    assert torch.equal(out["pred_ids"], ref["pred_ids"])
    assert torch.equal(out["padding_masks"], ref["padding_masks"])
    close(out["pred_logits"], ref["pred_logits"])
    close(out["full_logits"], ref["full_logits"])
    close(out["pred_masks"], ref["pred_masks"])
    assert all(not s.dead for s in cut.video_ins_hub.values())


@pytest.mark.parametrize("case", ["topk5_filled", "topk_all"])
@torch.no_grad()
def test_daq_offline_window_loop_matches_reference(golden, case):
    """pipeline.DAQOfflineRunner against the unmodified DVIS_DAQ_offline.run_window_inference (fixture:
    tests/golden/make_golden_daq_runner.py offline): the cutter's surviving sequences, top-k, the MinVIS-linked fill-in
    queries and the DAQ refiner on top -- same instance ids, logits and masks (instances compared in id order: the
    reference's topk(sorted=False) leaves their order open)."""
    import types
    from dvis_plus_b200.pipeline import DAQOfflineRunner
    g = golden("daq_offline_runner_small.pt")
    seg, K = g["seg"], g["num_classes"]
    C, fQ = seg["pred_embds"].shape[1], seg["pred_embds"].shape[3]
    cut = M.VideoInstanceCutter(hidden_dim=C, feedforward_dim=128, num_head=8, decoder_layer_num=2, mask_dim=C, num_classes=K,
                                num_new_ins=fQ, inference_select_threshold=0.1, kick_out_frame_num=2, num_slots=3,
                                keep_threshold=0.01, ovis_infer=True).eval()
    assert not any(cut.load_state_dict(g["cutter"]))
    rf = M.DAQTemporalRefiner(hidden_channel=C, feedforward_channel=128, num_head=8, decoder_layer_num=2, mask_dim=C, class_num=K,
                              windows=3, use_local_attn=False).eval()
    assert not any(rf.load_state_dict(g["refiner"]))
    emb = torch.nn.Embedding(fQ, C)
    emb.weight.data.copy_(g["query_feat"])
    predictor = types.SimpleNamespace(query_feat=emb, query_embed=emb)

    def segment(window):
        idx = window["frames"]
        return {"pred_embds": seg["pred_embds"][:, :, idx], "mask_features": seg["mask_features"][idx],
                "pred_logits": seg["pred_logits"][:, idx], "pred_masks": seg["pred_masks"][:, :, idx]}

    c = g["cases"][case]
    random.seed(g["seed"])
    torch.manual_seed(24)
    runner = DAQOfflineRunner(None, predictor, cut, rf, K, g["aux_inference_select_thr"], g["noise_frame_num"],
                              offline_topk_ins=c["offline_topk_ins"], window_size=g["window_size"], segment=segment, to_store="cpu")
    out, ref = runner({"frames": torch.arange(seg["pred_embds"].shape[2])}), c["out"]
    assert out["shape"] == ref["shape"]
    order, ref_order = out["pred_ids"][0].argsort(), ref["pred_ids"][0].argsort()
    assert torch.equal(out["pred_ids"][0][order], ref["pred_ids"][0][ref_order])
    # the refiner attends across instances, so logits / masks are compared after aligning the instance order
    close(out["pred_logits"][0][order], ref["pred_logits"][0][ref_order])
    close(out["pred_masks"][0][order], ref["pred_masks"][0][ref_order])
