"""TEST INFRASTRUCTURE ONLY.  ctypes front end of tests/hostcore/_build/libpostproc_host.so: the host compilation of the
post-processing kernels' per-pixel arithmetic, called with CPU torch tensors."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import hostcore_build as _build  # noqa: E402

_lib = None
_DT = {torch.float32: 0, torch.bfloat16: 2}


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _geom(masks, first_resize_size, img_size, out_size):
    return [int(v) for v in (masks.shape[-2], masks.shape[-1], *first_resize_size, *img_size, *out_size)]


def vis_masks(masks, sel, first_resize_size, img_size, out_size):
    """masks (Q, T, h, w) f32|bf16 with contiguous (h, w) planes; sel (n,) int64 or None -> (n, T, Ho, Wo) bool."""
    assert masks.stride(-1) == 1 and masks.stride(-2) == masks.shape[-1]
    n = masks.shape[0] if sel is None else sel.numel()
    T = masks.shape[1]
    out = torch.zeros((n, T, *out_size), dtype=torch.uint8)
    i64 = ctypes.c_int64
    rc = lib().hostcore_vis_masks(_p(masks), _DT[masks.dtype], i64(masks.stride(0)), i64(masks.stride(1)), _p(sel), n, T,
                                  *_geom(masks, first_resize_size, img_size, out_size), _p(out))
    assert rc == 0
    return out.bool()


def vps_argmax(masks, keep_idx, keep_score, first_resize_size, img_size, out_size):
    T = masks.shape[1]
    n = keep_idx.numel()
    win = torch.zeros((T, *out_size), dtype=torch.int32)
    areas = torch.zeros(3 * n, dtype=torch.int64)
    i64 = ctypes.c_int64
    rc = lib().hostcore_vps_argmax(_p(masks), _DT[masks.dtype], i64(masks.stride(0)), i64(masks.stride(1)), _p(keep_idx),
                                   _p(keep_score), n, T, *_geom(masks, first_resize_size, img_size, out_size), _p(win), _p(areas))
    assert rc == 0
    return win, areas.view(3, n)


def vss_argmax(masks, mask_cls, first_resize_size, img_size, out_size):
    Q, T = masks.shape[:2]
    out = torch.zeros((T, *out_size), dtype=torch.int64)
    i64 = ctypes.c_int64
    rc = lib().hostcore_vss_argmax(_p(masks), _DT[masks.dtype], i64(masks.stride(0)), i64(masks.stride(1)), _p(mask_cls),
                                   i64(mask_cls.stride(0)), Q, mask_cls.shape[1], T,
                                   *_geom(masks, first_resize_size, img_size, out_size), _p(out))
    assert rc == 0
    return out
