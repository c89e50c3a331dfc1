"""TEST INFRASTRUCTURE ONLY.  Rehearse bench.py's `main()` WITHOUT a GPU: the real control flow of the measured legs (resident,
end to end, parity check, instrumented eager pass; --world2: two ranks over gloo with the round-robin temporal stage), the
real runners and module fast paths, every CUDA-core kernel on the SIMT emulator, a fake CUDA runtime (streams / events are
no-ops, a graph replay re-runs the region the runner captured).  Sizes are cut down (32x64 frames, 2 frames, 10 queries, 1-2
layers per stack).  Times printed by this run mean nothing; what it checks is that the line is assembled and that the timed
path's results equal the eager runner's.

    python tests/simt/rehearse_bench.py [--world2]
"""
import contextlib
import functools
import io
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "simt")]
from emulated_device import emulated_b200  # noqa: E402
from rehearse_gpu_tests import install_host_shims  # noqa: E402

import bench  # noqa: E402


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

    def elapsed_time(self, other):
        return 1.0


class _Graph:
    """No capture on CPU: the runner constructors are wrapped (give_graphs_bodies) so that every graph of a slot gets a body
    recomputing what the captured region computes, from the slot's buffers into the slot's buffers."""
    body = None

    def replay(self):
        self.body()


def give_graphs_bodies():
    import dvis_plus_b200.pipeline as P
    g_init, r_init = P.GraphedClipRunner.__init__, P.RoundRobinClipRunner.__init__

    def graphed_init(self, runner, *a, **k):
        g_init(self, runner, *a, **k)
        for slot in self.slots:
            def stage_a(slot=slot):
                blk, mf = runner.segment_stage(slot["in"])
                slot["block"].copy_(blk)
                slot["mf"].copy_(mf)

            def stage_b(slot=slot):
                out = self._stage_b(slot["gathered"], slot["mf"], self._C(slot["block"]))
                for key in out:
                    slot["out"][key].copy_(out[key])
            slot["ga"].body, slot["gb"].body = stage_a, stage_b

    def rr_init(self, runner, *a, **k):
        r_init(self, runner, *a, **k)
        for slot in self.slots:
            def stage_a(slot=slot):
                blk, mf = runner.segment_stage(slot["in"])
                slot["block"].copy_(blk)
                slot["mf"].copy_(mf)

            def temporal(slot=slot):
                slot["payload"].copy_(self._payload(slot["gathered"], self._C(slot["block"])))

            def masks(slot=slot):
                out = self._finish(slot["payload"], slot["mf"], self._C(slot["block"]))
                for key in out:
                    slot["out"][key].copy_(out[key])
            slot["ga"].body, slot["gt"].body, slot["gm"].body = stage_a, temporal, masks
    P.GraphedClipRunner.__init__, P.RoundRobinClipRunner.__init__ = graphed_init, rr_init


def run(argv):
    install_host_shims()
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.Stream, torch.cuda.Event, torch.cuda.CUDAGraph = _Stream, _Event, _Graph
    torch.cuda.current_stream = lambda *a, **k: _Stream()
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.graph = lambda g, stream=None: contextlib.nullcontext()
    torch.cuda.is_current_stream_capturing = lambda: False
    bench.synthetic_features = functools.partial(bench.synthetic_features, hw=(32, 64))
    bench.build_models = functools.partial(bench.build_models, enc_layers=1, dec_layers=1, trk_layers=1, ref_layers=1)
    give_graphs_bodies()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:                    # --world 2: the ranks talk over gloo
        import torch.distributed as dist
        orig_init = dist.init_process_group
        dist.init_process_group = lambda backend=None, device_id=None, **k: orig_init("gloo", **k)
    exits = []
    os._exit = lambda code: (_ for _ in ()).throw(SystemExit(code)) if not exits.append(code) else None
    sys.argv = ["bench.py"] + argv
    out = io.StringIO()
    with emulated_b200(), contextlib.redirect_stdout(out):
        try:
            bench.main()
        except SystemExit:
            pass
    lines = [ln for ln in out.getvalue().splitlines() if ln.startswith("{")]
    if int(os.environ.get("RANK", "0")) != 0:
        assert not lines, out.getvalue()                              # only rank 0 prints
        return None
    assert len(lines) == 1, out.getvalue()
    return json.loads(lines[0])


def run_world(world, port=29631):
    """bench.py on `world` ranks (one process each, gloo): frames sharded, the real all-gather / broadcast / barriers, and
    the non-zero ranks' side of the guarded extra legs.  -> rank 0's line."""
    import subprocess
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="4")
        procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--as-rank"], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=1500) for p in procs]
    for rank, (p, (o, e)) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d: %s" % (rank, e[-3000:])
    return json.loads(outs[0][0][outs[0][0].index("{"):])


if __name__ == "__main__":
    flags = set(sys.argv[1:])
    if "--world2" in flags:
        print(json.dumps(run_world(2), indent=1))
        sys.exit(0)
    if "--as-rank" in flags:
        w = os.environ["WORLD_SIZE"]
        line = run(["--gpus", w, "--frames", w, "--queries", "10", "--steps", "1", "--warmup", "1", "--no-cpu-baseline"])
        if line is not None:
            print(json.dumps(line))
        sys.exit(0)
    line = run(["--frames", "2", "--queries", "10", "--steps", "1", "--warmup", "1", "--no-cpu-baseline"])
    print(json.dumps(line, indent=1))
