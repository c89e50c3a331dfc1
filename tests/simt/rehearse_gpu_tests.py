"""TEST INFRASTRUCTURE ONLY.  Rehearse the -m gpu test files WITHOUT a GPU: the same test functions, the real
dvis_plus_b200.ops front ends and module fast paths, every CUDA-core kernel on the SIMT emulator (plain-loop doubles for the
tcgen05 mask GEMM) -- emulated_device.emulated_b200 plus a few shims so the tests' own `.cuda()` / `device="cuda"` calls
stay on the host and the tracker runs its frame body without CUDA-graph capture.

    python tests/simt/rehearse_gpu_tests.py [-k substring] [test_module ...]     (default: the four files below; ~15 min)

Expected artefacts of the rehearsal, not failures of the code: tests asserting that CPU tensors are REJECTED (every tensor
claims to be on the device here), tests that need the reference's own CUDA kernel, and the full-size cases skipped below.
Used in round 1 to check that the refactored Hungarian kernel and the post-processing kernels written without GPU time
leave the existing GPU suites green (test_modules_gpu, test_msda_gpu, test_msda_backward_gpu, test_postprocess_gpu: all
cases passed except those artefacts)."""
import functools
import inspect
import itertools
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "simt")]
from emulated_device import emulated_b200  # noqa: E402

import dvis_plus_b200.modules as M  # noqa: E402

DEFAULT = ["test_postprocess_gpu", "test_modules_gpu", "test_msda_gpu", "test_msda_backward_gpu"]
SKIP = {"test_vis_masks_full_size": "minutes per case in emulation",
        "test_pipeline_vis_from_block_equals_postprocessing_all_masks": "twin in tests/test_simt_modules.py",
        "test_backward_matches_reference_cuda_kernel": "needs the reference's CUDA kernel",
        "test_argument_errors_like_reference": "asserts that CPU tensors are rejected"}


def _is_cuda_dev(a):
    return (isinstance(a, str) and a.startswith("cuda")) or (isinstance(a, torch.device) and a.type == "cuda")


def install_host_shims():
    orig_init = M.ReferringTracker_noiser.__init__

    def init(self, *a, **k):
        orig_init(self, *a, **k)
        self.use_cuda_graph = False                      # graph capture is a CUDA-runtime feature
    M.ReferringTracker_noiser.__init__ = init
    orig_to, orig_mod_to = torch.Tensor.to, torch.nn.Module.to

    def tensor_to(t, *a, **k):
        a = tuple("cpu" if _is_cuda_dev(x) else x for x in a)
        if _is_cuda_dev(k.get("device")):
            k["device"] = "cpu"
        return orig_to(t, *a, **k)
    torch.Tensor.to = tensor_to
    torch.nn.Module.to = lambda self, *a, **k: orig_mod_to(self, *tuple("cpu" if _is_cuda_dev(x) else x for x in a),
                                                           **{kk: ("cpu" if _is_cuda_dev(v) else v) for kk, v in k.items()})
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.synchronize = lambda *a, **k: None
    for fname in ("randn", "rand", "zeros", "ones", "empty", "full", "arange", "tensor", "as_tensor", "randint", "randperm",
                  "linspace", "eye"):
        def wrap(f):
            @functools.wraps(f)
            def g(*a, **k):
                if _is_cuda_dev(k.get("device")):
                    k.pop("device")
                return f(*a, **k)
            return g
        setattr(torch, fname, wrap(getattr(torch, fname)))


class EnvPatch:
    """The slice of pytest's monkeypatch the GPU tests use."""

    def __init__(self):
        self.saved = {}

    def setenv(self, k, v):
        self.saved.setdefault(k, os.environ.get(k))
        os.environ[k] = v

    def undo(self):
        for k, v in self.saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def cases_of(fn):
    names, values = [], []
    for m in [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]:
        n, v = m.args[0], list(m.args[1])
        if "," in n:
            names.append(tuple(x.strip() for x in n.split(",")))
            values.append(v)
        else:
            names.append((n,))
            values.append([(x,) for x in v])
    for combo in itertools.product(*values) if values else [()]:
        kwargs = {}
        for ns, vs in zip(names, combo):
            kwargs.update(dict(zip(ns, vs)))
        yield kwargs


def main(mods, only=None):
    install_host_shims()
    golden = lambda name: torch.load(os.path.join(ROOT, "tests", "golden", name), map_location="cpu", weights_only=False)  # noqa: E731
    ran = failed = 0
    with emulated_b200():
        for modname in mods:
            T = __import__(modname)
            for name, fn in inspect.getmembers(T, inspect.isfunction):
                if not name.startswith("test_") or fn.__module__ != modname:
                    continue
                if only is not None and only not in name:
                    continue
                if name in SKIP:
                    print(f"{modname}::{name} skipped ({SKIP[name]})", flush=True)
                    continue
                for kwargs in cases_of(fn):
                    params = inspect.signature(fn).parameters
                    env = EnvPatch()
                    if "golden" in params:
                        kwargs["golden"] = golden
                    if "monkeypatch" in params:
                        kwargs["monkeypatch"] = env
                    t = time.time()
                    try:
                        fn(**kwargs)
                        status = "ok"
                    except BaseException as e:           # pytest's Failed derives from BaseException
                        status = "FAIL %s: %s" % (type(e).__name__, str(e)[:160])
                        failed += 1
                    finally:
                        env.undo()
                    ran += 1
                    shown = {k: v for k, v in kwargs.items() if k not in ("golden", "monkeypatch")}
                    print(f"{modname}::{name} {shown} {status} {time.time() - t:.1f}s", flush=True)
    print(f"rehearsed {ran} cases, {failed} failed")
    return failed


if __name__ == "__main__":
    argv, only = sys.argv[1:], None
    if "-k" in argv:
        i = argv.index("-k")
        only = argv[i + 1]
        del argv[i:i + 2]
    sys.exit(1 if main(argv or DEFAULT, only) else 0)
