"""On-device parity of the runners bench.py times: the CUDA-graphed, software-pipelined GraphedClipRunner
(2 clips in flight; 3 clips in flight with the result copies on their own stream = the bench's N = 1 path) and the
RoundRobinClipRunner (world 1: the same three graphs per clip the N > 1 path replays) against the eager OfflineClipRunner on the
bench workload (Swin-L channel widths, 720p, Q = 200; T = 8 here to keep the pinned buffers small) -- BIT-IDENTICAL, for two
different clips interleaved through the slots, with host inputs and host outputs like the bench's end-to-end leg.
World 2 over NCCL is bench.py's own `parity_check` (0.0 at N = 2, 4, 8: profiles/r2_scale_n*.json); gloo world 2 on the CPU:
tests/test_pipeline_dist.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu
KEYS = ("pred_masks", "pred_logits")


@pytest.fixture(scope="module")
def workload():
    import bench
    from dvis_plus_b200.modules.precision import set_precision
    set_precision("bf16")
    runner = bench.build_models("cuda", queries=200)
    clips = []
    for seed in (0, 1):
        host = {k: v.contiguous(memory_format=torch.channels_last).pin_memory() for k, v in bench.synthetic_features(8, "swinl", seed=seed).items()}
        with torch.no_grad():
            out = runner({k: v.cuda() for k, v in host.items()})
        clips.append((host, {k: out[k].cpu() for k in KEYS}))
    torch.cuda.synchronize()
    return runner, clips


def _drive(g, clips, n=5):
    outs = []
    for i in range(n):
        host, ref = clips[i % 2]
        buf = {k: torch.empty(ref[k].shape, dtype=ref[k].dtype).pin_memory() for k in KEYS}
        g.submit(host, buf)
        outs.append((buf, ref))
    g.wait_all()
    torch.cuda.synchronize()
    for i, (buf, ref) in enumerate(outs):
        for k in KEYS:
            assert torch.equal(buf[k], ref[k]), f"clip {i}: {k} differs from the eager runner"


@pytest.mark.parametrize("depth,d2h_stream", [(2, False), (3, True)])
def test_graphed_pipelined_runner_is_bit_identical_to_the_eager_runner(workload, depth, d2h_stream):
    from dvis_plus_b200.pipeline import GraphedClipRunner
    runner, clips = workload
    g = GraphedClipRunner(runner, {k: v.cuda() for k, v in clips[0][0].items()}, depth=depth, d2h_stream=d2h_stream)
    assert g.captured_launches > 100, "libdvis_b200 kernels were not captured"
    _drive(g, clips)


def test_round_robin_runner_world1_is_bit_identical_to_the_eager_runner(workload):
    from dvis_plus_b200.pipeline import RoundRobinClipRunner
    runner, clips = workload
    g = RoundRobinClipRunner(runner, {k: v.cuda() for k, v in clips[0][0].items()})
    _drive(g, clips)


def test_eager_runner_is_deterministic(workload):
    runner, clips = workload
    host, ref = clips[1]
    with torch.no_grad():
        out = runner({k: v.cuda() for k, v in host.items()})
    for k in KEYS:
        assert torch.equal(out[k].cpu(), ref[k])
