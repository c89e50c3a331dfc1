"""TEST INFRASTRUCTURE ONLY.  `with emulated_b200():` makes the PRODUCT's host code (dvis_plus_b200.ops and the module
fast paths that run on the B200) execute on CPU tensors: the ctypes binding is pointed at tests/simt/_build/libdvis_simt.so
(the kernels' original sources on the SIMT emulator; plain-loop test doubles for the tcgen05 mask GEMM), tensors report
`is_cuda`, and the two CUDA-runtime touch points of ops.py (current stream, device guard) become no-ops.  Library calls the
modules make (cuBLAS / cuDNN through torch) run as torch CPU ops.  Everything is restored on exit."""
import contextlib
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import simt_binding  # noqa: E402

from dvis_plus_b200 import _lib, ops  # noqa: E402


@contextlib.contextmanager
def emulated_b200():
    l = simt_binding.lib()
    missing = [n for n in _lib.SIGNATURES if not hasattr(l, n)]
    assert not missing, f"entry points without an emulated kernel or test double: {missing}"
    for name, argtypes in _lib.SIGNATURES.items():
        fn = getattr(l, name)
        fn.argtypes, fn.restype = argtypes, ctypes.c_int
    l.dvis_abi_version.restype = ctypes.c_int
    saved = (_lib._lib, ops._stream, torch.cuda.device, torch.Tensor.is_cuda)
    _lib._lib = l
    ops._stream = lambda: None
    torch.cuda.device = lambda device: contextlib.nullcontext()
    torch.Tensor.is_cuda = property(lambda self: True)
    try:
        yield
    finally:
        _lib._lib, ops._stream, torch.cuda.device, torch.Tensor.is_cuda = saved
