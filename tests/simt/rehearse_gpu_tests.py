"""TEST INFRASTRUCTURE ONLY.  Rehearse tests/test_postprocess_gpu.py (the -m gpu parity tests of the post-processing kernels)
WITHOUT a GPU: the same test functions, the real dvis_plus_b200.ops front ends, kernels on the SIMT emulator
(emulated_device.emulated_b200).  Skipped: the full-size case (minutes in emulation) and the pipeline case that needs CUDA
graphs (its twin lives in tests/test_simt_modules.py).  Usage: python tests/simt/rehearse_gpu_tests.py"""
import inspect
import itertools
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "simt")]
from emulated_device import emulated_b200
import test_postprocess_gpu as T
T.DEV = "cpu"
golden = lambda name: torch.load(os.path.join(ROOT, 'tests', 'golden', name), map_location='cpu', weights_only=False)  # noqa: E731
class MP:
    def __init__(self): self.saved = {}
    def setenv(self, k, v): self.saved.setdefault(k, os.environ.get(k)); os.environ[k] = v
    def undo(self):
        for k, v in self.saved.items():
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = v
skip = {"test_vis_masks_full_size", "test_pipeline_vis_from_block_equals_postprocessing_all_masks"}
ran = 0
with emulated_b200():
    for name, fn in inspect.getmembers(T, inspect.isfunction):
        if not name.startswith("test_") or name in skip: continue
        marks = [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]
        names, values = [], []
        for m in marks:
            names.append(m.args[0]); values.append(m.args[1])
        for combo in itertools.product(*values) if values else [()]:
            kwargs = dict(zip(names, combo))
            params = inspect.signature(fn).parameters
            mp = MP()
            if "golden" in params: kwargs["golden"] = golden
            if "monkeypatch" in params: kwargs["monkeypatch"] = mp
            t = time.time()
            try:
                fn(**kwargs)
            finally:
                mp.undo()
            ran += 1
        print(name, "ok", flush=True)
print("rehearsed", ran, "test cases")
