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
ENTRY_POINTS = ("dvis_class_scores", "dvis_vis_topk", "dvis_vis_masks", "dvis_vis_masks_packed", "dvis_vps_argmax", "dvis_vps_paint",
                "dvis_vss_argmax", "dvis_lap_chain", "dvis_flash_attn")


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


def vis_masks_packed(m, sel, first, img, out_size):
    assert m.stride(3) == 1 and m.stride(2) == m.shape[3]
    n = m.shape[0] if sel is None else sel.numel()
    T = m.shape[1]
    out = torch.full((n, T, out_size[0], (out_size[1] + 7) // 8), 0xAA, dtype=torch.uint8)
    call("dvis_vis_masks_packed", _p(m), _DT[m.dtype], m.stride(0), m.stride(1), _p(sel), n, T, *_geom(m, first, img, out_size),
         _p(out), None)
    return out


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


# ---- the remaining CUDA-core entry points (same argument order as dvis_plus_b200/ops.py) ------------------------------------
ENTRY_POINTS += ("dvis_msda_forward", "dvis_msda_backward", "dvis_msda_fused_forward", "dvis_msda_fused_forward_hm", "dvis_add_layernorm", "dvis_groupnorm_nhwc",
                 "dvis_resize_bilinear_nhwc", "dvis_attn_bias_from_logits", "dvis_msda_pack_pairs",
                 "dvis_msda_pair_forward", "dvis_lap_rect")
_DT[torch.float64] = 1


def msda_forward(value, shapes, lsi, loc, attn, item_order=None):
    N, S, M, D = value.shape
    L, Lq, P = shapes.shape[0], loc.shape[1], loc.shape[4]
    out = torch.full((N, Lq, M * D), float("nan"), dtype=value.dtype)
    call("dvis_msda_forward", _p(value), _p(shapes), _p(lsi), _p(loc), _p(attn), N, S, M, D, L, Lq, P, _DT[value.dtype],
         _p(item_order), _p(out), None)
    return out


def msda_backward(value, shapes, lsi, loc, attn, grad_out):
    N, S, M, D = value.shape
    L, Lq, P = shapes.shape[0], loc.shape[1], loc.shape[4]
    gv, gl, ga = torch.zeros_like(value), torch.zeros_like(loc), torch.zeros_like(attn)
    call("dvis_msda_backward", _p(value), _p(shapes), _p(lsi), _p(loc), _p(attn), _p(grad_out), N, S, M, D, L, Lq, P,
         _DT[value.dtype], _p(gv), _p(gl), _p(ga), None)
    return gv, gl, ga


def msda_fused_forward(value, shapes, lsi, offsets, logits, ref, L, P, item_order=None, out_dtype=None, pair=False, head_major=False):
    """offsets (N, Lq, M*L*P*2), logits (N, Lq, M*L*P) -- possibly column slices of one tensor; ref (N, Lq, L, 2|4) f32."""
    N, S, M, D = value.shape
    Lq = offsets.shape[1]
    out_dtype = out_dtype or value.dtype
    out = torch.full((N, Lq, M * D), float("nan"), dtype=out_dtype)
    if head_major:
        hm_al = value.permute(0, 2, 1, 3).contiguous()
        call("dvis_msda_fused_forward_hm", _p(hm_al), _p(shapes), _p(lsi), _p(offsets), offsets.stride(1), _p(logits), logits.stride(1),
             _DT[offsets.dtype], _p(ref), ref.shape[-1], N, S, M, D, L, Lq, P, _p(item_order), _p(out), None)
        return out
    if pair:
        pairs = torch.zeros((N, S + 1, M, 2, D), dtype=torch.bfloat16)
        call("dvis_msda_pack_pairs", _p(value), N, S, M, D, _p(pairs), None)
        call("dvis_msda_pair_forward", _p(pairs), _p(shapes), _p(lsi), _p(offsets), offsets.stride(1), _p(logits), logits.stride(1),
             _DT[offsets.dtype], _p(ref), ref.shape[-1], N, S, M, D, L, Lq, P, _p(item_order), _p(out), None)
        return out
    call("dvis_msda_fused_forward", _p(value), _DT[value.dtype], _p(shapes), _p(lsi), _p(offsets), offsets.stride(1), _p(logits),
         logits.stride(1), _DT[offsets.dtype], _p(ref), ref.shape[-1], N, S, M, D, L, Lq, P, _p(item_order), _p(out),
         _DT[out_dtype], None)
    return out


def add_layernorm(x, residual, weight, bias, eps=1e-5, lp_dtype=None, pos=None):
    C = x.shape[-1]
    rows = x.numel() // C
    out32 = torch.full(x.shape, float("nan"), dtype=torch.float32)
    lp = torch.zeros(x.shape, dtype=lp_dtype) if lp_dtype is not None else None
    lpp = torch.zeros(x.shape, dtype=lp_dtype) if pos is not None else None
    call("dvis_add_layernorm", _p(x), _DT[x.dtype], _p(residual), _DT[residual.dtype] if residual is not None else 0, _p(weight),
         _p(bias), _p(pos), pos.numel() // C if pos is not None else 0, rows, C, float(eps), _p(out32), _p(lp), _p(lpp),
         _DT[lp_dtype] if lp_dtype is not None else 0, None)
    return out32, lp, lpp


def groupnorm_nhwc(x, G, weight, bias, eps=1e-5, relu=False, up=None, up_hw=None, hw=None, pos=None, lp_dtype=torch.bfloat16):
    N, HW, C = x.shape
    ws = torch.empty(2 * N * G * (1 + (HW + 255) // 256), dtype=torch.float64)
    out32 = torch.full((N, HW, C), float("nan"), dtype=torch.float32)
    lp = torch.zeros((N, HW, C), dtype=lp_dtype)
    lpp = torch.zeros((N, HW, C), dtype=lp_dtype) if pos is not None else None
    uh, uw = up_hw if up is not None else (0, 0)
    H, W = hw if up is not None else (0, 0)
    call("dvis_groupnorm_nhwc", _p(x), _DT[x.dtype], x.stride(0), N, HW, C, G, _p(weight), _p(bias), float(eps), int(relu), _p(ws),
         _p(up), up.stride(0) if up is not None else 0, uh, uw, H, W, _p(pos), _p(out32), _p(lp), _p(lpp), _DT[lp_dtype],
         out32.stride(0), None)
    return out32, lp, lpp


def resize_bilinear_nhwc(x_nhwc, size):
    N, h, w, C = x_nhwc.shape
    out = torch.zeros((N, size[0], size[1], C), dtype=torch.bfloat16)
    call("dvis_resize_bilinear_nhwc", _p(x_nhwc), N, h, w, C, _p(out), size[0], size[1], None)
    return out


def attn_bias_from_logits(logits, dtype=torch.float32):
    hw = logits.shape[-1]
    bias = torch.full(logits.shape, 7.0, dtype=dtype)
    call("dvis_attn_bias_from_logits", _p(logits), logits.numel() // hw, hw, _p(bias), _DT[dtype], None)
    return bias


def flash_attn(q, k, v, scale, mask_bits=None):
    B, Lq, H, Dh = q.shape
    Lk = k.shape[1]
    out = torch.full((B, Lq, H * Dh), float("nan"), dtype=torch.bfloat16)
    call("dvis_flash_attn", _p(q), q.stride(1), q.stride(0), q.stride(2), _p(k), k.stride(1), k.stride(0), k.stride(2), _p(v),
         v.stride(1), v.stride(0), v.stride(2), _p(out), H * Dh, Lq * H * Dh, _p(mask_bits),
         mask_bits.stride(1) if mask_bits is not None else 0, mask_bits.stride(0) if mask_bits is not None else 0,
         B, Lq, Lk, H, Dh, float(scale), None)
    return out


def linear_small(w, bias=None, **kw):
    """the product's own front end (dvis_plus_b200.ops.linear_small) on the emulated library"""
    import emulated_device
    from dvis_plus_b200 import ops
    with emulated_device.emulated_b200():
        return ops.linear_small(w, bias, **kw)


def linear_small_ln(w, bias, x, residual, ln, **kw):
    import emulated_device
    from dvis_plus_b200 import ops
    with emulated_device.emulated_b200():
        return ops.linear_small_ln(w, bias, x, residual, ln, **kw)


def lap_rect(cost):
    c = cost.float().contiguous()
    B = 1 if c.dim() == 2 else c.shape[0]
    out = torch.full(c.shape[:-1], -7, dtype=torch.int64)
    call("dvis_lap_rect", _p(c), B, c.shape[-2], c.shape[-1], _p(out), None)
    return out
