"""Host-side logic of the frame-sharded clip pipeline on CPU with the gloo backend, world_size 2:
the packed query all-gather keeps temporal order, tracker + refiner are identical on all ranks, each rank produces the
masks of its own frames, and the sharded result equals the single-process result."""
import os
import socket
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dvis_plus_b200 import modules as M
from dvis_plus_b200.pipeline import OfflineClipRunner

T, Q, C, K, H, W = 4, 10, 64, 5, 8, 12


def _models():
    torch.manual_seed(0)
    trk = M.ReferringTracker_noiser(hidden_channel=C, feedforward_channel=128, num_head=8, decoder_layer_num=2,
                                    mask_dim=32, class_num=K, noise_mode="none").eval()
    rfn = M.TemporalRefiner(hidden_channel=C, feedforward_channel=128, num_head=8, decoder_layer_num=2, mask_dim=32,
                            class_num=K, windows=2).eval()
    return trk, rfn


def _seg_inputs():
    g = torch.Generator().manual_seed(1)
    return dict(pred_embds=torch.randn(1, C, T, Q, generator=g), pred_embds_without_norm=torch.randn(1, C, T, Q, generator=g),
                pred_logits=torch.randn(1, T, Q, K + 1, generator=g)), torch.randn(T, 32, H, W, generator=g)


def _slice(seg, mf, t0, t1):
    return dict(pred_embds=seg["pred_embds"][:, :, t0:t1], pred_embds_without_norm=seg["pred_embds_without_norm"][:, :, t0:t1],
                pred_logits=seg["pred_logits"][:, t0:t1]), mf[t0:t1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        trk, rfn = _models()
        runner = OfflineClipRunner(None, None, trk, rfn)
        assert runner.world == world and runner.rank == rank
        seg, mf = _seg_inputs()
        t = T // world
        seg_r, mf_r = _slice(seg, mf, rank * t, (rank + 1) * t)
        block = runner.gather_queries(runner.pack_queries(seg_r))
        assert torch.equal(block, runner.pack_queries(seg))          # contiguous blocks -> temporal order preserved
        out = runner.temporal_stage(seg_r, mf_r)
        torch.save({k: v for k, v in out.items()}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_frame_sharded_pipeline_world2_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, port, d), nprocs=2, join=True)
        r0 = torch.load(os.path.join(d, "rank0.pt"))
        r1 = torch.load(os.path.join(d, "rank1.pt"))
    trk, rfn = _models()
    single = OfflineClipRunner(None, None, trk, rfn)
    assert single.world == 1
    seg, mf = _seg_inputs()
    ref = single.temporal_stage(seg, mf)
    for k in ("pred_logits", "pred_embds", "online_pred_logits"):
        assert torch.equal(r0[k], r1[k]), k                           # replicated stage is bit-identical across ranks
        assert torch.allclose(r0[k], ref[k], atol=1e-6), k
    masks = torch.cat([r0["pred_masks"], r1["pred_masks"]], dim=2)    # (1, Q, T, H, W): rank order == frame order
    assert torch.allclose(masks, ref["pred_masks"], atol=1e-5)


def test_pack_unpack_roundtrip():
    seg, _ = _seg_inputs()
    block = OfflineClipRunner.pack_queries(seg)
    assert block.shape == (T, Q, 2 * C + K + 1)
    e, n, l = OfflineClipRunner.unpack_queries(block, C)
    assert torch.equal(e, seg["pred_embds"]) and torch.equal(n, seg["pred_embds_without_norm"]) and torch.equal(l, seg["pred_logits"])


# ---- the B200 fast path of the frame-sharded pipeline, world size 2, on the emulated device -------------------------------
def _vis_inputs():
    g = torch.Generator().manual_seed(3)
    seg = dict(pred_embds=torch.randn(1, C, T, Q, generator=g), pred_embds_without_norm=torch.randn(1, C, T, Q, generator=g),
               pred_logits=torch.randn(1, T, Q, K + 1, generator=g))
    mf = torch.randn(T, 64, H, W, generator=g).to(torch.bfloat16, memory_format=torch.channels_last)
    return seg, mf


def _vis_models():
    from dvis_plus_b200.modules.postprocess import VideoPostProcessor
    torch.manual_seed(0)
    trk = M.ReferringTracker_noiser(hidden_channel=C, feedforward_channel=128, num_head=8, decoder_layer_num=2, mask_dim=64,
                                    class_num=K, noise_mode="none").eval()
    trk.use_cuda_graph = False
    rfn = M.TemporalRefiner(hidden_channel=C, feedforward_channel=128, num_head=8, decoder_layer_num=2, mask_dim=64, class_num=K,
                            windows=2).eval()
    return trk, rfn, VideoPostProcessor(K, num_queries=Q, max_num=4)


def _run_vis(runner, post, seg, mf):
    """Kernels run from their original sources on the SIMT emulator (tests/simt/emulated_device.py)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt"))
    from emulated_device import emulated_b200
    from dvis_plus_b200.modules.precision import precision
    with emulated_b200(), torch.no_grad(), precision("bf16"):
        block = runner.gather_queries(runner.pack_queries(seg))          # the one collective of the path (gloo here, NCCL on the box)
        return runner.vis_from_block(block, mf, C, post, (4 * H - 3, 4 * W - 2), (5 * H, 5 * W + 1))


def _vis_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        trk, rfn, post = _vis_models()
        runner = OfflineClipRunner(None, None, trk, rfn)
        seg, mf = _vis_inputs()
        t = T // world
        seg_r, mf_r = _slice(seg, mf, rank * t, (rank + 1) * t)
        out = _run_vis(runner, post, seg_r, mf_r.contiguous(memory_format=torch.channels_last))
        torch.save(dict(out), os.path.join(out_dir, f"vis_rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_frame_sharded_vis_postprocessing_world2_on_the_emulated_device():
    """vis_from_block (instance selection before the final mask GEMM + fused resize / threshold) with the frames split over
    two ranks: scores / labels / ids identical on both ranks and equal to the single-process result, masks = the two ranks'
    frame blocks side by side."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_vis_worker, args=(2, port, d), nprocs=2, join=True)
        r0 = torch.load(os.path.join(d, "vis_rank0.pt"))
        r1 = torch.load(os.path.join(d, "vis_rank1.pt"))
    trk, rfn, post = _vis_models()
    seg, mf = _vis_inputs()
    ref = _run_vis(OfflineClipRunner(None, None, trk, rfn), post, seg, mf)
    for k in ("pred_scores", "pred_labels", "pred_ids"):
        assert torch.equal(r0[k], r1[k]) and torch.equal(r0[k], ref[k]), k
    masks = torch.cat([r0["pred_masks"], r1["pred_masks"]], dim=1)      # (n, T, H_out, W_out): rank order == frame order
    assert masks.shape == ref["pred_masks"].shape == (4, T, 5 * H, 5 * W + 1)
    assert torch.equal(masks, ref["pred_masks"])


# ---- round-robin ownership of the temporal stage for streams of clips -----------------------------------------------------
def _clip(i):
    g = torch.Generator().manual_seed(100 + i)
    seg = dict(pred_embds=torch.randn(1, C, T, Q, generator=g), pred_embds_without_norm=torch.randn(1, C, T, Q, generator=g),
               pred_logits=torch.randn(1, T, Q, K + 1, generator=g))
    return seg, torch.randn(T, 32, H, W, generator=g)


class _SegmentStub:
    """Stands in for pixel decoder + predictor: the 'features' of a clip ARE its (sliced) segmenter outputs."""

    def forward_features(self, feats):
        return feats["mask_features"], None, feats


def _rr_worker(rank, world, port, out_dir):
    from dvis_plus_b200.pipeline import RoundRobinClipRunner
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        trk, rfn = _models()
        runner = OfflineClipRunner(_SegmentStub(), lambda ms, mf: {k: v for k, v in ms.items() if k != "mask_features"}, trk, rfn)
        t = T // world

        def local(i):
            seg, mf = _clip(i)
            seg_r, mf_r = _slice(seg, mf, rank * t, (rank + 1) * t)
            return dict(seg_r, mask_features=mf_r)
        rr = RoundRobinClipRunner(runner, local(0))
        assert rr.depth == world + 2 and not rr.cuda
        outs = []
        for i in range(5):                                            # owners: 0, 1, 0, 1, 0
            outs.append({k: v.clone() for k, v in rr.submit(local(i))["out"].items()})
        rr.wait_all()
        torch.save(outs, os.path.join(out_dir, f"rr_rank{rank}.pt"))
        # the same with the fused VIS post-processing as the last stage: the owner selects the instances and broadcasts only
        # their mask embeddings (kernels on the SIMT emulator)
        import sys
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt"))
        from emulated_device import emulated_b200
        from dvis_plus_b200.modules.postprocess import VideoPostProcessor
        from dvis_plus_b200.modules.precision import precision
        trk.use_cuda_graph = False
        vis = dict(post=VideoPostProcessor(K, num_queries=Q, max_num=3), img_size=(30, 45), output_size=(40, 57))
        rv = RoundRobinClipRunner(runner, local(0), graphs=False, vis=vis)    # built on CPU tensors: the synchronous gloo path
        assert not rv.cuda
        with emulated_b200(), precision("fp32"):
            vouts = [{k: v.clone() for k, v in rv.submit(local(i))["out"].items()} for i in range(3)]
        torch.save(vouts, os.path.join(out_dir, f"rrvis_rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_round_robin_temporal_ownership_world2_matches_replicated():
    """RoundRobinClipRunner: clip n's tracker + refiner run on rank n mod 2 only and are broadcast; every clip's results
    equal the single-process (replicated) pipeline's."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_rr_worker, args=(2, port, d), nprocs=2, join=True)
        r0 = torch.load(os.path.join(d, "rr_rank0.pt"))
        r1 = torch.load(os.path.join(d, "rr_rank1.pt"))
        v0 = torch.load(os.path.join(d, "rrvis_rank0.pt"))
        v1 = torch.load(os.path.join(d, "rrvis_rank1.pt"))
    trk, rfn = _models()
    single = OfflineClipRunner(None, None, trk, rfn)
    for i in range(5):
        seg, mf = _clip(i)
        ref = single.temporal_stage(seg, mf)
        for k in ("pred_logits", "pred_embds", "online_pred_logits"):
            assert torch.equal(r0[i][k], r1[i][k]), (i, k)            # both ranks hold the owner's result
            assert torch.allclose(r0[i][k], ref[k], atol=1e-6), (i, k)
        masks = torch.cat([r0[i]["pred_masks"], r1[i]["pred_masks"]], dim=2)
        assert torch.allclose(masks, ref["pred_masks"], atol=1e-5), i
    # fused VIS variant: both ranks report the owner's selection; masks = the two frame blocks side by side == single process
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt"))
    from emulated_device import emulated_b200
    from dvis_plus_b200.modules.postprocess import VideoPostProcessor
    from dvis_plus_b200.modules.precision import precision
    trk.use_cuda_graph = False
    post = VideoPostProcessor(K, num_queries=Q, max_num=3)
    for i in range(3):
        seg, mf = _clip(i)
        with emulated_b200(), precision("fp32"), torch.no_grad():
            ref = single.vis_from_block(single.pack_queries(seg), mf, C, post, (30, 45), (40, 57))
        for k in ("pred_scores", "pred_labels", "pred_ids"):
            assert torch.equal(v0[i][k], v1[i][k]) and torch.allclose(v0[i][k].float(), ref[k].float(), atol=1e-6), (i, k)
        masks = torch.cat([v0[i]["pred_masks"], v1[i]["pred_masks"]], dim=1)
        assert (masks != ref["pred_masks"]).float().mean().item() < 1e-3, i
