"""The stream / event / CUDA-graph orchestration of pipeline.GraphedClipRunner and pipeline.RoundRobinClipRunner, executed
on CPU with a FAKE CUDA runtime (streams and events are no-ops, graph capture runs the body once, replay re-runs it) on top
of the emulated device.  This cannot say anything about overlap or timing; it checks that the device branches of the
runners -- which the gloo tests never enter -- run without errors and produce the eager pipeline's results."""
import contextlib
import os
import sys

import pytest
import torch

from dvis_plus_b200 import modules as M
from dvis_plus_b200.modules.precision import precision
from dvis_plus_b200.pipeline import GraphedClipRunner, OfflineClipRunner, RoundRobinClipRunner

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt"))

T, Q, C, K, H, W = 4, 10, 64, 5, 8, 12


class _Stream:
    def __init__(self, priority=0):
        pass

    def wait_stream(self, other):
        pass

    def wait_event(self, event):
        pass


class _Event:
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass


class _Graph:
    """Capture = remember a closure that re-runs the captured region; replay = run it again."""
    body = None

    def replay(self):
        self.body()


@contextlib.contextmanager
def fake_cuda_runtime(monkeypatch):
    monkeypatch.setattr(torch.cuda, "Stream", _Stream)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "CUDAGraph", _Graph)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _Stream())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
    yield


class _SegmentStub:
    """pixel decoder + predictor stand-in: the clip's 'features' are its segmenter outputs."""

    def forward_features(self, feats):
        return feats["mask_features"], None, feats


def _setup():
    torch.manual_seed(0)
    trk = M.ReferringTracker_noiser(hidden_channel=C, feedforward_channel=128, num_head=8, decoder_layer_num=2, mask_dim=64,
                                    class_num=K, noise_mode="none").eval()
    trk.use_cuda_graph = False
    rfn = M.TemporalRefiner(hidden_channel=C, feedforward_channel=128, num_head=8, decoder_layer_num=2, mask_dim=64, class_num=K,
                            windows=2).eval()
    runner = OfflineClipRunner(_SegmentStub(), lambda ms, mf: {k: v for k, v in ms.items() if k != "mask_features"}, trk, rfn)

    def clip(i):
        g = torch.Generator().manual_seed(50 + i)
        return dict(pred_embds=torch.randn(1, C, T, Q, generator=g), pred_embds_without_norm=torch.randn(1, C, T, Q, generator=g),
                    pred_logits=torch.randn(1, T, Q, K + 1, generator=g),
                    mask_features=torch.randn(T, 64, H, W, generator=g).to(torch.bfloat16, memory_format=torch.channels_last))
    return runner, clip


@pytest.mark.timeout(1200)
@pytest.mark.parametrize("cls", [GraphedClipRunner, RoundRobinClipRunner])
@torch.no_grad()
def test_device_branches_of_the_clip_runners(cls, monkeypatch):
    from emulated_device import emulated_b200
    runner, clip = _setup()
    with emulated_b200(), precision("bf16"), fake_cuda_runtime(monkeypatch):
        # graph capture under the fake runtime: torch.cuda.graph(g, stream=...) records a closure re-running the region.
        # The runners capture regions that read slot buffers and write slot outputs, so re-running the region == replay.
        captured = []

        @contextlib.contextmanager
        def fake_graph(g, stream=None):
            captured.append(g)
            yield
        monkeypatch.setattr(torch.cuda, "graph", fake_graph)
        eager = [runner(clip(i)) for i in range(3)]
        kwargs = dict(depth=2, d2h_stream=True) if cls is GraphedClipRunner else dict(graphs=False)
        pipe = cls(runner, clip(0), **kwargs)
        assert getattr(pipe, "cuda", True)
        host = {k: torch.empty_like(v) for k, v in eager[0].items() if k in ("pred_masks", "pred_logits")}
        for i in range(3):
            if cls is GraphedClipRunner:
                # no real capture on CPU: give every graph of the slot a body that recomputes the slot from its inputs
                slot = pipe.slots[i % pipe.depth]

                def stage_a(slot=slot):
                    blk, mf = runner.segment_stage(slot["in"])
                    slot["block"].copy_(blk)
                    slot["mf"].copy_(mf)

                def stage_b(slot=slot):
                    out = runner.temporal_from_block(slot["gathered"], slot["mf"], pipe._C(slot["block"]))
                    for k in out:
                        slot["out"][k].copy_(out[k])
                slot["ga"].body, slot["gb"].body = stage_a, stage_b
            slot = pipe.submit(clip(i), d2h=host)
            pipe.wait_all()
            for k in ("pred_logits", "pred_masks"):
                assert torch.allclose(slot["out"][k].float(), eager[i][k].float(), atol=1e-5), (cls.__name__, i, k)
                assert torch.allclose(host[k].float(), eager[i][k].float(), atol=1e-5)
        # the graph-capturing constructor path of the round-robin runner (3 graphs per slot) at least runs
        if cls is RoundRobinClipRunner:
            from dvis_plus_b200 import _lib
            before = _lib.launch_count
            rr = RoundRobinClipRunner(runner, clip(0), depth=2, graphs=True)
            assert rr.graphs and len(rr.slots) == 2 and all(k in rr.slots[0] for k in ("ga", "gt", "gm", "payload", "out"))
            assert rr.captured_launches > 0 and _lib.launch_count > before
            for i in range(3):                        # ... and its submit() path with graph replays (bodies = the captured regions)
                slot = rr.slots[i % rr.depth]

                def seg(slot=slot):
                    blk, mf = runner.segment_stage(slot["in"])
                    slot["block"].copy_(blk)
                    slot["mf"].copy_(mf)

                def temporal(slot=slot):
                    slot["payload"].copy_(runner.temporal_payload(slot["gathered"], rr._C(slot["block"])))

                def masks(slot=slot):
                    out = runner.outputs_from_payload(slot["payload"], slot["mf"], rr._C(slot["block"]))
                    slot["out"]["pred_masks"].copy_(out["pred_masks"])          # the other entries are views of the payload
                slot["ga"].body, slot["gt"].body, slot["gm"].body = seg, temporal, masks
                slot = rr.submit(clip(i), d2h=host)
                rr.wait_all()
                for k in ("pred_logits", "pred_masks"):
                    assert torch.allclose(slot["out"][k].float(), eager[i][k].float(), atol=1e-5), ("graphs", i, k)


@pytest.mark.timeout(1200)
@pytest.mark.parametrize("cls", [GraphedClipRunner, RoundRobinClipRunner])
@torch.no_grad()
def test_device_branches_with_fused_vis_postprocessing(cls, monkeypatch):
    """`vis=`: stage B of the runners is vis_from_block (instances selected before the final mask GEMM, fused resize /
    threshold, bit-packed masks): equal to the eager call on every clip."""
    from emulated_device import emulated_b200
    from dvis_plus_b200 import ops
    from dvis_plus_b200.modules.postprocess import VideoPostProcessor
    runner, clip = _setup()
    post = VideoPostProcessor(K, num_queries=Q, max_num=4)
    vis = dict(post=post, img_size=(30, 45), output_size=(41, 61), packed=True)
    with emulated_b200(), precision("bf16"), fake_cuda_runtime(monkeypatch):
        @contextlib.contextmanager
        def fake_graph(g, stream=None):
            yield
        monkeypatch.setattr(torch.cuda, "graph", fake_graph)

        def eager(i):
            blk, mf = runner.segment_stage(clip(i))
            return runner.vis_from_block(blk, mf, C, **vis)
        pipe = cls(runner, clip(0), depth=2, vis=vis) if cls is GraphedClipRunner else cls(runner, clip(0), graphs=False, vis=vis)
        for i in range(3):
            if cls is GraphedClipRunner:
                slot = pipe.slots[i % pipe.depth]

                def stage_a(slot=slot):
                    blk, mf = runner.segment_stage(slot["in"])
                    slot["block"].copy_(blk)
                    slot["mf"].copy_(mf)

                def stage_b(slot=slot):
                    out = pipe._stage_b(slot["gathered"], slot["mf"], pipe._C(slot["block"]))
                    for k in out:
                        slot["out"][k].copy_(out[k])
                slot["ga"].body, slot["gb"].body = stage_a, stage_b
            out, ref = pipe.submit(clip(i))["out"], eager(i)
            pipe.wait_all()
            assert out["pred_masks"].dtype == torch.uint8 and out["pred_masks"].shape == (4, T, 41, 8)
            for k in ref:
                assert torch.equal(out[k], ref[k]), (cls.__name__, i, k)
            assert ops.unpack_masks(out["pred_masks"], 61).shape == (4, T, 41, 61)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("mode", ["one_rank", "two_ranks_gloo"])
def test_bench_main_rehearsal(mode):
    """bench.py's main() end to end on the emulated device + fake runtime (tests/simt/rehearse_bench.py, own processes
    because it patches torch globally): the line is assembled and the timed path's results -- 3 clips in flight on one rank;
    on two ranks (gloo: real all-gather / broadcast / barriers) the round-robin temporal stage -- equal the eager runner's."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    flag = [] if mode == "one_rank" else ["--world2"]
    p = subprocess.run([sys.executable, os.path.join(root, "tests", "simt", "rehearse_bench.py")] + flag,
                       capture_output=True, text=True, timeout=850, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout[p.stdout.index("{"):])
    assert line["gpu_launches"] > 0 and line["roofline"]["kernel"].startswith("msda")
    assert line["parity_check"]["bit_identical"], line["parity_check"]
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    if mode == "one_rank":
        assert line["n_gpus"] == 1 and "3 clips in flight" in line["config"]["execution"]
    else:
        assert line["n_gpus"] == 2 and "round-robin" in line["config"]["parallelism"]
