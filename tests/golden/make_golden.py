"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
Each fixture stores the seeded inputs (or the seed + recipe to regenerate them), the reference
module's state_dict where one exists, and the reference's outputs.  The oracle (oracle/) is pinned
against these in tests/test_oracle.py; the CUDA path is compared with them in the -m gpu tests.
"""
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import reference_loader as rl  # noqa: E402

warnings.filterwarnings("ignore")


def lsi_of(shapes):
    return torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))


def optest_inputs(dtype, N=1, M=2, D=2, Lq=2, L=2, P=2, shapes=((6, 4), (3, 2))):
    """Input recipe of OPS/test.py:24-39 (caller seeds)."""
    sh = torch.as_tensor(shapes, dtype=torch.long)
    S = int(sh.prod(1).sum())
    value = torch.rand(N, S, M, D) * 0.01
    loc = torch.rand(N, Lq, M, L, P, 2)
    attn = torch.rand(N, Lq, M, L, P) + 1e-5
    attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    return value.to(dtype), sh, lsi_of(sh), loc.to(dtype), attn.to(dtype)


def save(name, obj):
    path = os.path.join(HERE, name)
    torch.save(obj, path)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


@torch.no_grad()
def main():
    R = rl.load()

    # 1. OPS/test.py verbatim: shapes :24-28, seed :31, double then float forward checks in file order
    torch.manual_seed(3)
    v, sh, lsi, loc, aw = optest_inputs(torch.float64)
    out64 = R.ms_deform_attn_core_pytorch(v, sh, loc, aw)
    v32, _, _, loc32, aw32 = optest_inputs(torch.float32)
    out32 = R.ms_deform_attn_core_pytorch(v32, sh, loc32, aw32)
    save("msda_optest.pt", dict(shapes=sh, lsi=lsi, value64=v, loc64=loc, attn64=aw, out64=out64,
                                value32=v32, loc32=loc32, attn32=aw32, out32=out32))

    # 2. realistic head geometry (M=8, D=32, L=3, P=4), locations partly outside [0,1]; fwd + grads in fp64
    torch.manual_seed(0)
    sh = torch.as_tensor([(12, 20), (6, 10), (3, 5)], dtype=torch.long)
    S = int(sh.prod(1).sum())
    N, M, D, Lq, L, P = 1, 8, 32, 50, 3, 4
    v = torch.randn(N, S, M, D, dtype=torch.float64)
    loc = torch.rand(N, Lq, M, L, P, 2, dtype=torch.float64) * 1.4 - 0.2
    aw = torch.rand(N, Lq, M, L, P, dtype=torch.float64).flatten(-2).softmax(-1).view(N, Lq, M, L, P)
    with torch.enable_grad():
        v.requires_grad_(); loc.requires_grad_(); aw.requires_grad_()
        out = R.ms_deform_attn_core_pytorch(v, sh, loc, aw)
        g = torch.randn_like(out)
        out.backward(g)
    save("msda_small.pt", dict(shapes=sh, lsi=lsi_of(sh), value=v.detach(), loc=loc.detach(), attn=aw.detach(),
                               out=out.detach(), grad_out=g, grad_value=v.grad, grad_loc=loc.grad, grad_attn=aw.grad))

    # 3. BASELINE config 1 (256x256, 1 level, 8 heads, 4 points, Q=100): inputs are regenerated from the
    #    seed at test time (67 MB value tensor), only the reference output is stored.
    torch.manual_seed(3)
    v, sh, lsi, loc, aw = optest_inputs(torch.float32, N=1, M=8, D=32, Lq=100, L=1, P=4, shapes=((256, 256),))
    save("msda_cfg1_out.pt", dict(seed=3, out=R.ms_deform_attn_core_pytorch(v, sh, loc, aw)))

    # 4. MSDeformAttn module (d_model=64, 3 levels, 8 heads, 4 points), non-trivial offsets / weights
    torch.manual_seed(1)
    m = R.MSDeformAttn(d_model=64, n_levels=3, n_heads=8, n_points=4).eval()
    torch.nn.init.normal_(m.sampling_offsets.weight, std=0.05)
    torch.nn.init.normal_(m.attention_weights.weight, std=0.2)
    torch.nn.init.normal_(m.attention_weights.bias, std=0.2)
    sh = torch.as_tensor([(8, 12), (4, 6), (2, 3)], dtype=torch.long)
    S = int(sh.prod(1).sum())
    q = torch.randn(2, S, 64)
    src = torch.randn(2, S, 64)
    ref = torch.rand(2, S, 3, 2)
    out = m(q, ref, src, sh, lsi_of(sh), None)
    pad = torch.zeros(2, S, dtype=torch.bool)
    pad[:, ::7] = True
    out_pad = m(q, ref, src, sh, lsi_of(sh), pad)
    ref4 = torch.cat([ref, torch.rand(2, S, 3, 2) * 0.3], -1)
    out_box = m(q, ref4, src, sh, lsi_of(sh), None)
    save("msdeformattn_module.pt", dict(state_dict=m.state_dict(), shapes=sh, query=q, src=src, ref=ref, out=out,
                                        padding_mask=pad, out_pad=out_pad, ref4=ref4, out_box=out_box))

    # 5. pixel decoder (conv_dim 64, 2 encoder layers, 4 backbone maps of a 64x96 image)
    torch.manual_seed(2)
    chans = dict(res2=8, res3=16, res4=24, res5=32)
    strides = dict(res2=4, res3=8, res4=16, res5=32)
    shapes = {k: R.ShapeSpec(channels=chans[k], stride=strides[k]) for k in chans}
    pd = R.MSDeformAttnPixelDecoder(shapes, transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=128,
                                    transformer_enc_layers=2, conv_dim=64, mask_dim=64, norm="GN",
                                    transformer_in_features=["res3", "res4", "res5"], common_stride=4).eval()
    for layer in pd.transformer.encoder.layers:   # a trained model has non-zero offset/attention weights
        torch.nn.init.normal_(layer.self_attn.sampling_offsets.weight, std=0.02)
        torch.nn.init.normal_(layer.self_attn.attention_weights.weight, std=0.1)
    feats = {k: torch.randn(2, chans[k], 64 // strides[k], 96 // strides[k]) for k in chans}
    mf, o0, ms = pd.forward_features(feats)
    save("pixel_decoder_small.pt", dict(state_dict=pd.state_dict(), features=feats, mask_features=mf, out0=o0, multi_scale=ms))

    # 6. segmenter predictor (dvisPlus decoder): hidden 64, Q=12, 3 layers, reid head; eval, T=2 frames
    torch.manual_seed(4)
    dec = R.Decoder_dvisPlus(64, True, num_classes=5, hidden_dim=64, num_queries=12, nheads=8, dim_feedforward=128,
                             dec_layers=3, pre_norm=False, mask_dim=64, enforce_input_project=False, num_frames=2,
                             num_reid_head_layers=3, reid_hidden_dim=64).eval()
    out = dec(ms, mf)
    save("predictor_small.pt", dict(state_dict=dec.state_dict(), multi_scale=ms, mask_features=mf,
                                    pred_logits=out["pred_logits"], pred_masks=out["pred_masks"],
                                    pred_embds=out["pred_embds"], pred_embds_without_norm=out["pred_embds_without_norm"],
                                    aux_masks=[a["pred_masks"] for a in out["aux_outputs"]]))

    # 7. forward_prediction_heads alone (the mask head): Q=12 queries on the 16x24 map
    qfeat = torch.randn(12, 2, 64)
    cls, masks, am = dec.forward_prediction_heads(qfeat, mf, attn_mask_target_size=(4, 6))
    save("mask_head_small.pt", dict(output=qfeat, mask_features=mf, target_size=(4, 6), cls=cls, masks=masks, attn_mask=am))

    # 8. tracker: hidden 64, 2 layers, T=3 as windows [0:2] then [2:3] with resume
    torch.manual_seed(5)
    trk = R.ReferringTracker_noiser(hidden_channel=64, feedforward_channel=128, num_head=8, decoder_layer_num=2,
                                    mask_dim=64, class_num=5, noise_mode="none").eval()
    T, Q = 3, 12
    fe = torch.randn(1, 64, T, Q)
    fe_nn = fe + 0.1 * torch.randn(1, 64, T, Q)
    mfeat = torch.randn(1, T, 64, 16, 24)
    o1, i1 = trk(fe[:, :, :2], mfeat[:, :2], resume=False, return_indices=True, frame_embeds_no_norm=fe_nn[:, :, :2])
    o2, i2 = trk(fe[:, :, 2:], mfeat[:, 2:], resume=True, return_indices=True, frame_embeds_no_norm=fe_nn[:, :, 2:])
    save("tracker_small.pt", dict(
        state_dict=trk.state_dict(), frame_embeds=fe, frame_embeds_no_norm=fe_nn, mask_features=mfeat,
        pred_logits=torch.cat([o1["pred_logits"], o2["pred_logits"]], 1),
        pred_masks=torch.cat([o1["pred_masks"], o2["pred_masks"]], 2),
        pred_embds=torch.cat([o1["pred_embds"], o2["pred_embds"]], 2),
        pred_references=torch.cat([o1["pred_references"], o2["pred_references"]], 2),
        indices=[torch.as_tensor(x) for x in (i1 + i2)]))

    # 9. refiner: hidden 64, 2 layers, T=7 (exercises the k=5 replicate padding), windows=3
    torch.manual_seed(6)
    rf = R.TemporalRefiner(hidden_channel=64, feedforward_channel=128, num_head=8, decoder_layer_num=2,
                           mask_dim=64, class_num=5, windows=3).eval()
    T = 7
    inst = torch.randn(1, 64, T, Q)
    fr = torch.randn(1, 64, T, Q)
    mfeat = torch.randn(1, T, 64, 16, 24)
    o = rf(inst, fr, mfeat)
    save("refiner_small.pt", dict(state_dict=rf.state_dict(), instance_embeds=inst, frame_embeds=fr, mask_features=mfeat,
                                  pred_logits=o["pred_logits"], pred_masks=o["pred_masks"], pred_embds=o["pred_embds"]))


class cuda_to_cpu:
    """The DAQ reference hard-codes `.to("cuda")` (D/dvis_daq/track_module.py:757,779): while generating fixtures on this
    GPU-less box the harness maps that device string to "cpu"; the reference source itself stays untouched."""

    def __enter__(self):
        self._orig = torch.Tensor.to
        orig = self._orig

        def to(t, *args, **kwargs):
            args = tuple("cpu" if (isinstance(a, str) and a.startswith("cuda")) else a for a in args)
            if isinstance(kwargs.get("device"), str) and kwargs["device"].startswith("cuda"):
                kwargs["device"] = "cpu"
            return orig(t, *args, **kwargs)
        torch.Tensor.to = to

    def __exit__(self, *exc):
        torch.Tensor.to = self._orig


@torch.no_grad()
def main_daq():
    import random
    R = rl.load()
    # 10. DVIS-DAQ tracker inference: hidden 64, 2 layers, num_new_ins = fQ new-instance queries (track_module.py:640-641), 3 slots; T=4 frames as windows [0:3],[3:4]
    torch.manual_seed(7)
    random.seed(7)
    C, fQ, T, H, W = 64, 10, 4, 16, 24
    cut = R.VideoInstanceCutter(hidden_dim=C, feedforward_dim=128, num_head=8, decoder_layer_num=2, mask_dim=C, num_classes=5,
                                num_new_ins=fQ, inference_select_threshold=0.1, kick_out_frame_num=2, num_slots=3,
                                keep_threshold=0.01, ovis_infer=True).eval()
    torch.nn.init.normal_(cut.class_embed.weight, std=0.5)          # spread the scores so that some queries are (in)valid
    seg_query_feat = torch.nn.Embedding(fQ, C)
    fe = torch.randn(1, C, T, fQ)
    mf = torch.randn(1, T, C, H, W)
    valid = [[torch.rand(fQ) > 0.4] for _ in range(T)]
    pm = [[torch.randn(fQ, H, W)] for _ in range(T)]
    info = lambda a, b: {"seg_query_feat": seg_query_feat, "valid": valid[a:b], "pred_masks": pm[a:b]}
    with cuda_to_cpu():
        cut.inference(fe[:, :, :3], mf[:, :3], info(0, 3), 0, resume=False, to_store="cpu")
        cut.inference(fe[:, :, 3:], mf[:, 3:], info(3, 4), 3, resume=True, to_store="cpu")
    seqs = []
    for sid in cut.memory_seq_ids:
        s = cut.video_ins_hub[sid]
        seqs.append(dict(sT=s.sT, dead=s.dead, appearance=list(s.appearance), embeds=torch.stack(s.embeds),
                         pred_logits=torch.stack(s.pred_logits), pred_masks=torch.stack(s.pred_masks),
                         pos=s.similarity_guided_pos_embed))
    save("daq_tracker_small.pt", dict(state_dict=cut.state_dict(), seg_query_feat=seg_query_feat.weight.detach(), frame_embeds=fe,
                                      mask_features=mf, valid=[v[0] for v in valid], pred_masks=[p[0] for p in pm], seqs=seqs,
                                      track_queries=cut.track_queries, track_embeds=cut.track_embeds, seed=7))

    # 11. SlotCrossAttentionLayer alone
    torch.manual_seed(8)
    sl = R.SlotCrossAttentionLayer(d_model=C, nhead=8).eval()
    tgt, mem, qp, sq = torch.randn(7, 1, C), torch.randn(10, 1, C), torch.randn(7, 1, C), torch.randn(7, 1, C)
    save("daq_slot_layer.pt", dict(state_dict=sl.state_dict(), tgt=tgt, memory=mem, query_pos=qp, slot_query=sq,
                                   out=sl(tgt, mem, query_pos=qp, slot_query=sq)))

    # 12. DAQ TemporalRefiner (no local conv branch, like the released configs), eval
    torch.manual_seed(9)
    rf = R.DAQTemporalRefiner(hidden_channel=C, feedforward_channel=128, num_head=8, decoder_layer_num=2, mask_dim=C,
                              class_num=5, windows=3, use_local_attn=False).eval()
    Tn, Q = 5, 9
    inst, fr, mfeat = torch.randn(1, C, Tn, Q), torch.randn(1, C, Tn, Q), torch.randn(1, Tn, C, H, W)
    o = rf(inst, torch.zeros(1, Q, Tn, dtype=torch.bool), fr, mfeat, None)
    save("daq_refiner_small.pt", dict(state_dict=rf.state_dict(), instance_embeds=inst, frame_embeds=fr, mask_features=mfeat,
                                      pred_logits=o["pred_logits"], pred_masks=o["pred_masks"], pred_embds=o["pred_embds"]))


@torch.no_grad()
def main_variants():
    """Less common constructor variants the drop-ins also mirror: pre-norm blocks, DAQ refiner with the conv branch."""
    R = rl.load()
    g = torch.load(os.path.join(HERE, "predictor_small.pt"), weights_only=False)
    torch.manual_seed(10)
    dec = R.Decoder_dvisPlus(64, True, num_classes=5, hidden_dim=64, num_queries=12, nheads=8, dim_feedforward=128,
                             dec_layers=2, pre_norm=True, mask_dim=64, enforce_input_project=True, num_frames=2,
                             num_reid_head_layers=0, reid_hidden_dim=64).eval()
    out = dec(g["multi_scale"], g["mask_features"])
    save("predictor_prenorm_small.pt", dict(state_dict=dec.state_dict(), pred_logits=out["pred_logits"], pred_masks=out["pred_masks"],
                                            pred_embds=out["pred_embds"], pred_embds_without_norm=out["pred_embds_without_norm"]))
    torch.manual_seed(11)
    rf = R.DAQTemporalRefiner(hidden_channel=64, feedforward_channel=128, num_head=8, decoder_layer_num=2, mask_dim=64,
                              class_num=5, windows=4, use_local_attn=True).eval()
    inst, fr, mfeat = torch.randn(1, 64, 6, 7), torch.randn(1, 64, 6, 7), torch.randn(1, 6, 64, 8, 12)
    o = rf(inst, torch.zeros(1, 7, 6, dtype=torch.bool), fr, mfeat, None)
    save("daq_refiner_localattn_small.pt", dict(state_dict=rf.state_dict(), instance_embeds=inst, frame_embeds=fr, mask_features=mfeat,
                                                pred_logits=o["pred_logits"], pred_masks=o["pred_masks"], pred_embds=o["pred_embds"]))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "variants":
        main_variants()
    elif len(sys.argv) > 1 and sys.argv[1] == "daq":
        main_daq()
    else:
        main()
        main_daq()
        main_variants()
