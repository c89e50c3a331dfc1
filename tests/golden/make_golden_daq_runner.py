"""Golden vectors for the DVIS-DAQ online window loop: the UNMODIFIED DVIS_DAQ_online.run_window_inference
(D/dvis_daq/meta_architecture.py:488-597) driven with a stand-in `self` whose backbone / segmenter head return precomputed
(seeded) segmenter outputs, and the reference VideoInstanceCutter as the tracker.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_daq_runner.py
"""
import os
import random
import sys
import types
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import reference_loader as rl  # noqa: E402
from make_golden import cuda_to_cpu  # noqa: E402

warnings.filterwarnings("ignore")


class FakeHead:
    """sem_seg_head stand-in: slices precomputed segmenter outputs by the frame indices the fake backbone passes on."""

    def __init__(self, seg, query_feat, query_embed, num_classes):
        self.seg, self.num_classes = seg, num_classes
        self.predictor = types.SimpleNamespace(query_feat=query_feat, query_embed=query_embed)

    def __call__(self, features):
        idx = features["res2"]
        s = self.seg
        return {"aux_outputs": [], "pred_embds": s["pred_embds"][:, :, idx], "mask_features": s["mask_features"][idx],
                "pred_logits": s["pred_logits"][:, idx], "pred_masks": s["pred_masks"][:, :, idx]}


@torch.no_grad()
def main():
    rl.load()
    A = rl.load_daq_meta_architecture()
    R = rl.load()
    torch.manual_seed(17)
    random.seed(17)
    C, fQ, T, H, W, K = 64, 8, 7, 8, 12, 5
    cut = R.VideoInstanceCutter(hidden_dim=C, feedforward_dim=128, num_head=8, decoder_layer_num=2, mask_dim=C, num_classes=K,
                                num_new_ins=fQ, inference_select_threshold=0.1, kick_out_frame_num=2, num_slots=3,
                                keep_threshold=0.01, ovis_infer=True).eval()
    torch.nn.init.normal_(cut.class_embed.weight, std=0.5)
    query_feat, query_embed = torch.nn.Embedding(fQ, C), torch.nn.Embedding(fQ, C)
    seg = dict(pred_embds=torch.randn(1, C, T, fQ), mask_features=torch.randn(T, C, H, W),
               pred_logits=torch.randn(1, T, fQ, K + 1) * 2, pred_masks=torch.randn(1, fQ, T, H, W))
    me = types.SimpleNamespace(
        backbone=lambda idx: {"res2": idx, "res3": 0, "res4": 0, "res5": 0},
        sem_seg_head=FakeHead(seg, query_feat, query_embed, K), tracker=cut, keep=False, aux_inference_select_thr=0.3,
        noise_frame_num=2)
    with cuda_to_cpu():
        out = A.DVIS_DAQ_online.run_window_inference(me, torch.arange(T))        # window_size is hard-coded to 5 (py:491)
    assert len(out["pred_logits"]) > 0, "fixture without any surviving instance"
    print({k: (tuple(v.shape) if torch.is_tensor(v) else v) for k, v in out.items()})
    torch.save(dict(state_dict=cut.state_dict(), query_feat=query_feat.weight.detach(), seg=seg, num_classes=K, seed=17,
                    aux_inference_select_thr=0.3, noise_frame_num=2, out=out), os.path.join(HERE, "daq_runner_small.pt"))
    print("daq_runner_small.pt: %.1f KiB" % (os.path.getsize(os.path.join(HERE, "daq_runner_small.pt")) / 1024))




@torch.no_grad()
def main_offline():
    """DVIS_DAQ_offline.run_window_inference (py:1332-1365 -> common_inference py:1169-1330 -> minvis_post_processing
    py:1400-1437): segmenter over all windows, cutter over all windows, surviving instance sequences padded to the clip
    length, top-k by score, the remaining query slots filled with MinVIS-linked segmenter queries, DAQ refiner."""
    A = rl.load_daq_meta_architecture()
    R = rl.load()
    torch.manual_seed(23)
    random.seed(23)
    C, fQ, T, H, W, K = 64, 8, 7, 8, 12, 5
    cut = R.VideoInstanceCutter(hidden_dim=C, feedforward_dim=128, num_head=8, decoder_layer_num=2, mask_dim=C, num_classes=K,
                                num_new_ins=fQ, inference_select_threshold=0.1, kick_out_frame_num=2, num_slots=3,
                                keep_threshold=0.01, ovis_infer=True).eval()
    torch.nn.init.normal_(cut.class_embed.weight, std=0.5)
    rf = R.DAQTemporalRefiner(hidden_channel=C, feedforward_channel=128, num_head=8, decoder_layer_num=2, mask_dim=C, class_num=K,
                              windows=3, use_local_attn=False).eval()
    query_feat, query_embed = torch.nn.Embedding(fQ, C), torch.nn.Embedding(fQ, C)
    seg = dict(pred_embds=torch.randn(1, C, T, fQ), mask_features=torch.randn(T, C, H, W),
               pred_logits=torch.randn(1, T, fQ, K + 1) * 2, pred_masks=torch.randn(1, fQ, T, H, W))
    cases = {}
    for name, topk in (("topk5_filled", 5), ("topk_all", 100)):
        cut._clear_memory() if hasattr(cut, "_clear_memory") else None
        me = types.SimpleNamespace(
            backbone=lambda idx: {"res2": idx, "res3": 0, "res4": 0, "res5": 0},
            sem_seg_head=FakeHead(seg, query_feat, query_embed, K), tracker=cut, refiner=rf, keep=False, training=False,
            aux_inference_select_thr=0.3, noise_frame_num=2, offline_topk_ins=topk)
        for meth in ("segmenter_windows_inference", "common_inference", "minvis_post_processing"):
            setattr(me, meth, types.MethodType(getattr(A.DVIS_DAQ_offline, meth), me))
        me.match_from_embds = types.MethodType(A.MinVIS.match_from_embds, me)
        random.seed(23)
        torch.manual_seed(24)
        if name != "topk5_filled":                        # a fresh tracker memory per case
            cut2 = R.VideoInstanceCutter(hidden_dim=C, feedforward_dim=128, num_head=8, decoder_layer_num=2, mask_dim=C, num_classes=K,
                                         num_new_ins=fQ, inference_select_threshold=0.1, kick_out_frame_num=2, num_slots=3,
                                         keep_threshold=0.01, ovis_infer=True).eval()
            cut2.load_state_dict(cut.state_dict())
            me.tracker = cut2
        with cuda_to_cpu():
            out = A.DVIS_DAQ_offline.run_window_inference(me, torch.arange(T), window_size=3)
        print(name, {k: (tuple(v.shape) if torch.is_tensor(v) else v) for k, v in out.items()})
        cases[name] = dict(offline_topk_ins=topk, out=out)
    torch.save(dict(cutter=cut.state_dict(), refiner=rf.state_dict(), query_feat=query_feat.weight.detach(), seg=seg, num_classes=K,
                    seed=23, window_size=3, aux_inference_select_thr=0.3, noise_frame_num=2, cases=cases),
               os.path.join(HERE, "daq_offline_runner_small.pt"))
    print("daq_offline_runner_small.pt: %.1f KiB" % (os.path.getsize(os.path.join(HERE, "daq_offline_runner_small.pt")) / 1024))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "offline":
        main_offline()
    else:
        main()
        main_offline()
