"""TEST INFRASTRUCTURE ONLY.  ctypes front end of tests/simt/_build/libdvis_simt.so: the C-ABI entry points of
csrc/postproc.cu and csrc/lap.cu executed by the SIMT emulator on CPU tensors.  Argument types are the product binding's
(dvis_plus_b200._lib.SIGNATURES), so the calls below read exactly like the ones in dvis_plus_b200/ops.py."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import simt_build as _build  # noqa: E402

from dvis_plus_b200 import _lib as product_binding  # noqa: E402  (signatures only; the product library is not loaded)

_DT = {torch.float32: 0, torch.bfloat16: 2}
_lib = None
ENTRY_POINTS = ("dvis_class_scores", "dvis_vis_topk", "dvis_vis_masks", "dvis_vps_argmax", "dvis_vps_paint",
                "dvis_vss_argmax", "dvis_lap_chain")


def lib():
    global _lib
    if _lib is None:
        l = ctypes.CDLL(_build.build())
        l.dvis_last_error.restype = ctypes.c_char_p
        for name in ENTRY_POINTS:
            fn = getattr(l, name)
            fn.argtypes = product_binding.SIGNATURES[name]
            fn.restype = ctypes.c_int
        _lib = l
    return _lib


def call(name, *args):
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib().dvis_last_error().decode()}")


def _p(t):
    return t.data_ptr() if t is not None else None


def _geom(m, first, img, out):
    return [int(v) for v in (m.shape[-2], m.shape[-1], *first, *img, *out)]


def class_scores(cls, aux=None):
    cls = cls.float().contiguous()
    aux = aux.float().contiguous() if aux is not None else None
    out = torch.empty_like(cls)
    call("dvis_class_scores", _p(cls), _p(aux), cls.shape[0], cls.shape[1], _p(out), None)
    return out


def vis_topk(cls, max_num, aux=None):
    cls = cls.float().contiguous()
    aux = aux.float().contiguous() if aux is not None else None
    Q, K1 = cls.shape
    ws = torch.empty(Q, K1)
    s, l, q = torch.empty(max_num), torch.empty(max_num, dtype=torch.int64), torch.empty(max_num, dtype=torch.int64)
    call("dvis_vis_topk", _p(cls), _p(aux), Q, K1, max_num, _p(ws), _p(s), _p(l), _p(q), None)
    return s, l, q


def vis_masks(m, sel, first, img, out_size):
    assert m.stride(3) == 1 and m.stride(2) == m.shape[3]
    n = m.shape[0] if sel is None else sel.numel()
    T = m.shape[1]
    out = torch.full((n, T, *out_size), 7, dtype=torch.uint8)          # 7: a pixel the kernel failed to write shows up
    call("dvis_vis_masks", _p(m), _DT[m.dtype], m.stride(0), m.stride(1), _p(sel), n, T, *_geom(m, first, img, out_size), _p(out), None)
    assert int(out.max()) <= 1, "unwritten output pixels"
    return out.bool()


def vps_argmax(m, keep_idx, keep_score, first, img, out_size):
    T, n = m.shape[1], keep_idx.numel()
    win = torch.full((T, *out_size), 1 << 20, dtype=torch.int32)
    areas = torch.full((3, n), -1, dtype=torch.int64)
    call("dvis_vps_argmax", _p(m), _DT[m.dtype], m.stride(0), m.stride(1), _p(keep_idx), _p(keep_score), n, T,
         *_geom(m, first, img, out_size), _p(win), _p(areas), None)
    return win, areas


def vps_paint(win, seg):
    out = torch.empty_like(win)
    call("dvis_vps_paint", _p(win), _p(seg), win.numel(), _p(out), None)
    return out


def vss_argmax(m, mask_cls, first, img, out_size):
    Q, T = m.shape[:2]
    out = torch.full((T, *out_size), -1, dtype=torch.int64)
    call("dvis_vss_argmax", _p(m), _DT[m.dtype], m.stride(0), m.stride(1), _p(mask_cls), mask_cls.stride(0), Q, mask_cls.shape[1], T,
         *_geom(m, first, img, out_size), _p(out), None)
    return out


def lap_chain(cost, idx_init=None):
    cost = cost.float().contiguous()
    T, n, _ = cost.shape
    sigma = torch.empty((T, n), dtype=torch.int64)
    idx = torch.empty((T, n), dtype=torch.int64)
    call("dvis_lap_chain", _p(cost), T, n, _p(idx_init), _p(sigma), _p(idx), None)
    return sigma, idx


def set_jitter(one_in):
    """Race shaker: delay roughly one in `one_in` threads after every __syncthreads() (0 = off)."""
    lib().simt_set_jitter(int(one_in))
