"""ctypes access to oracle/_ref/libref_msda.so -- the reference's own CUDA kernels built for sm_100a
(oracle/ref_cuda/build.py).  ORACLE / BASELINE INFRASTRUCTURE ONLY: used by tests/ as a second checker and as
"the reference CUDA kernel on the same B200" in the micro-benchmarks; never by the product."""
import ctypes
import os

import torch

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_msda.so")
_lib = None


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_PATH)
    return _lib


def _args(*ts):
    return [ctypes.c_void_p(t.data_ptr()) for t in ts]


def forward(value, shapes, lsi, loc, attn, im2col_step=128):
    N, S, M, D = value.shape
    L, Lq, P = shapes.shape[0], loc.shape[1], loc.shape[4]
    out = torch.zeros(N, Lq, M * D, device=value.device, dtype=torch.float32)   # the reference memsets (cu:59)
    rc = lib().ref_msda_forward_f32(*_args(value, shapes, lsi, loc, attn), N, S, M, D, L, Lq, P, im2col_step,
                                    ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, rc
    return out


def backward(value, shapes, lsi, loc, attn, grad_out, im2col_step=128):
    N, S, M, D = value.shape
    L, Lq, P = shapes.shape[0], loc.shape[1], loc.shape[4]
    gv, gl, ga = torch.zeros_like(value), torch.zeros_like(loc), torch.zeros_like(attn)
    rc = lib().ref_msda_backward_f32(*_args(value, shapes, lsi, loc, attn, grad_out), N, S, M, D, L, Lq, P, im2col_step,
                                     *_args(gv, gl, ga), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, rc
    return gv, gl, ga
