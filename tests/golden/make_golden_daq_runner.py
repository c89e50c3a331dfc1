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


if __name__ == "__main__":
    main()
