"""ctypes wrappers over the plain-C oracle (oracle/msda_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _prep(value, shapes, lsi, loc, attn):
    dt = np.float64 if value.dtype == np.float64 else np.float32
    value = np.ascontiguousarray(value, dtype=dt)
    loc = np.ascontiguousarray(loc, dtype=dt)
    attn = np.ascontiguousarray(attn, dtype=dt)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    lsi = np.ascontiguousarray(lsi, dtype=np.int64)
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    assert shapes.shape == (L, 2) and lsi.shape == (L,) and attn.shape == (N, Lq, M, L, P)
    return dt, value, shapes, lsi, loc, attn, (N, S, M, D, L, Lq, P)


def msda_forward(value, shapes, lsi, loc, attn):
    """value (N,S,M,D), shapes (L,2) int64, lsi (L,), loc (N,Lq,M,L,P,2), attn (N,Lq,M,L,P) -> (N,Lq,M*D)."""
    dt, value, shapes, lsi, loc, attn, dims = _prep(value, shapes, lsi, loc, attn)
    N, S, M, D, L, Lq, P = dims
    out = np.zeros((N, Lq, M * D), dtype=dt)
    fn = lib().msda_oracle_f64_forward if dt == np.float64 else lib().msda_oracle_f32_forward
    fn(_p(value), _p(shapes), _p(lsi), _p(loc), _p(attn), *[ctypes.c_int(x) for x in dims], _p(out))
    return out


def msda_backward(value, shapes, lsi, loc, attn, grad_out):
    dt, value, shapes, lsi, loc, attn, dims = _prep(value, shapes, lsi, loc, attn)
    grad_out = np.ascontiguousarray(grad_out, dtype=dt)
    gv, gl, ga = np.zeros_like(value), np.zeros_like(loc), np.zeros_like(attn)
    fn = lib().msda_oracle_f64_backward if dt == np.float64 else lib().msda_oracle_f32_backward
    fn(_p(value), _p(shapes), _p(lsi), _p(loc), _p(attn), _p(grad_out), *[ctypes.c_int(x) for x in dims],
       _p(gv), _p(gl), _p(ga))
    return gv, gl, ga


def mask_logits(emb, feat):
    """emb (B,Q,C), feat (B,C,H,W) -> (B,Q,H,W) with double accumulation."""
    emb = np.ascontiguousarray(emb, dtype=np.float32)
    feat = np.ascontiguousarray(feat, dtype=np.float32)
    B, Q, C = emb.shape
    _, _, H, W = feat.shape
    out = np.zeros((B, Q, H, W), dtype=np.float32)
    lib().mask_oracle_f32(_p(emb), _p(feat), ctypes.c_int(B), ctypes.c_int(Q), ctypes.c_int(C),
                          ctypes.c_long(H * W), _p(out))
    return out
