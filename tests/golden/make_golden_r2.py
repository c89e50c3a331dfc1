"""Round-2 fixtures: the UNMODIFIED reference tracker / refiner at a width the fused temporal-stage kernels cover
(hidden 128 = 4 heads x 32; csrc/small_linear.cu needs C % 128 == 0, csrc/flash_attn.cu head dims 32 / 64).

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_r2.py
"""
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import reference_loader as rl  # noqa: E402
from make_golden import save  # noqa: E402

warnings.filterwarnings("ignore")


@torch.no_grad()
def main():
    R = rl.load()
    # tracker: hidden 128, 4 heads, 2 layers, Q=10, T=4 as windows [0:3] then [3:4] with resume
    torch.manual_seed(11)
    trk = R.ReferringTracker_noiser(hidden_channel=128, feedforward_channel=256, num_head=4, decoder_layer_num=2,
                                    mask_dim=64, class_num=5, noise_mode="none").eval()
    T, Q = 4, 10
    fe = torch.randn(1, 128, T, Q)
    fe_nn = fe + 0.1 * torch.randn(1, 128, T, Q)
    mfeat = torch.randn(1, T, 64, 8, 12)
    o1, i1 = trk(fe[:, :, :3], mfeat[:, :3], resume=False, return_indices=True, frame_embeds_no_norm=fe_nn[:, :, :3])
    o2, i2 = trk(fe[:, :, 3:], mfeat[:, 3:], resume=True, return_indices=True, frame_embeds_no_norm=fe_nn[:, :, 3:])
    save("tracker_w128.pt", dict(
        state_dict=trk.state_dict(), frame_embeds=fe, frame_embeds_no_norm=fe_nn, mask_features=mfeat,
        pred_logits=torch.cat([o1["pred_logits"], o2["pred_logits"]], 1),
        pred_masks=torch.cat([o1["pred_masks"], o2["pred_masks"]], 2),
        pred_embds=torch.cat([o1["pred_embds"], o2["pred_embds"]], 2),
        pred_references=torch.cat([o1["pred_references"], o2["pred_references"]], 2),
        indices=[torch.as_tensor(x) for x in (i1 + i2)]))

    # refiner: hidden 128, 4 heads, 2 layers, T=7 (exercises the k=5 replicate padding), Q=10
    torch.manual_seed(12)
    rf = R.TemporalRefiner(hidden_channel=128, feedforward_channel=256, num_head=4, decoder_layer_num=2,
                           mask_dim=64, class_num=5, windows=3).eval()
    T = 7
    inst = torch.randn(1, 128, T, Q)
    fr = torch.randn(1, 128, T, Q)
    mfeat = torch.randn(1, T, 64, 8, 12)
    o = rf(inst, fr, mfeat)
    save("refiner_w128.pt", dict(state_dict=rf.state_dict(), instance_embeds=inst, frame_embeds=fr, mask_features=mfeat,
                                 pred_logits=o["pred_logits"], pred_masks=o["pred_masks"], pred_embds=o["pred_embds"]))


if __name__ == "__main__":
    main()
