"""ctypes binding of libdvis_b200.so (interface: include/dvis_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, a RuntimeError is raised
so a caller can never silently end up on a slow / different code path.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdvis_b200.so")

DVIS_F32, DVIS_F64, DVIS_BF16 = 0, 1, 2
ABI_VERSION = 1

_vp, _i, _i64, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

# name -> argtypes; mirrors include/dvis_b200.h one to one (tests/test_abi.py cross-checks against the header)
SIGNATURES = {
    "dvis_msda_forward": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp],
    "dvis_msda_backward": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "dvis_msda_fused_forward": [_vp, _i, _vp, _vp, _vp, _i64, _vp, _i64, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp,
                                _vp, _i, _vp],
    "dvis_msda_fused_forward_hm": [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp],
    "dvis_msda_pack_pairs": [_vp, _i, _i, _i, _i, _vp, _vp],
    "dvis_msda_pair_forward": [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp],
    "dvis_mask_logits_clip": [_vp, _vp, _i, _i, _i, _i64, _vp, _i, _vp],
    "dvis_mask_logits_tf32": [_vp, _vp, _i, _i, _i, _i64, _vp, _vp],
    "dvis_mask_attn_bias_tf32": [_vp, _vp, _i, _i, _i, _i64, _vp, _vp, _vp],
    "dvis_linear_tc": [_vp, _i64, _vp, _vp, _i, _i, _i, _i, _vp, _i64, _vp],
    "dvis_linear_tc_heads": [_vp, _i64, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp],
    "dvis_linear_tc_add_ln": [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _f, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp],
    "dvis_mask_logits": [_vp, _vp, _i, _i, _i, _i64, _vp, _i, _vp],
    "dvis_groupnorm_nhwc": [_vp, _i, _i64, _i, _i, _i, _i, _vp, _vp, _f, _i, _vp, _vp, _i64, _i, _i, _i, _i, _vp, _vp, _vp,
                            _vp, _i, _i64, _vp],
    "dvis_resize_bilinear_nhwc": [_vp, _i, _i, _i, _i, _vp, _i, _i, _vp],
    "dvis_level_tokens": [_vp, _i, _i64, _vp, _vp, _i, _i, _i, _vp, _vp, _vp],
    "dvis_attn_bias_from_logits": [_vp, _i64, _i, _vp, _i, _vp],
    "dvis_lap_chain": [_vp, _i, _i, _vp, _vp, _vp, _vp],
    "dvis_lap_rect": [_vp, _i, _i, _i, _vp, _vp],
    "dvis_mask_logits_strided": [_vp, _i64, _vp, _i, _i, _i, _i64, _vp, _i64, _i, _vp],
    "dvis_mask_attn_bias": [_vp, _vp, _i, _i, _i, _i64, _vp, _i, _vp, _vp],
    "dvis_mask_attn_bits": [_vp, _vp, _i, _i, _i, _i64, _vp, _i64, _vp, _vp],
    "dvis_flash_attn": [_vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64,
                        _i, _i, _i, _i, _i, _f, _vp],
    "dvis_linear_small": [_vp, _i64, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _f, _vp, _vp, _vp, _i64, _vp, _i64,
                          _vp, _i64, _i, _vp, _vp, _i64, _i64, _i, _i, _i, _i, _vp, _vp, _vp],
    "dvis_linear_small_ln": [_vp, _i64, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _f, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp,
                             _vp, _vp, _vp, _vp],
    "dvis_set_pdl": [_i],
    "dvis_debug_linear_small_stamps": [_vp],
    "dvis_class_scores": [_vp, _vp, _i, _i, _vp, _vp],
    "dvis_vis_topk": [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "dvis_vis_masks": [_vp, _i, _i64, _i64, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "dvis_vis_masks_packed": [_vp, _i, _i64, _i64, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "dvis_vps_argmax": [_vp, _i, _i64, _i64, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp],
    "dvis_vps_paint": [_vp, _vp, _i64, _vp, _vp],
    "dvis_vss_argmax": [_vp, _i, _i64, _i64, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "dvis_add_layernorm": [_vp, _i, _vp, _i, _vp, _vp, _vp, _i64, _i64, _i, _f, _vp, _vp, _vp, _i, _vp],
}

_lib = None
launch_count = 0  # number of C-ABI calls that launched device work (bench.py reports it as gpu_launches)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m dvis_plus_b200.csrc.build` "
                "(there is no CPU or PyTorch fallback for the dvis_plus_b200 kernels)")
        l = ctypes.CDLL(LIB_PATH)
        l.dvis_abi_version.restype = _i
        l.dvis_last_error.restype = ctypes.c_char_p
        if l.dvis_abi_version() != ABI_VERSION:
            raise RuntimeError(f"libdvis_b200 ABI {l.dvis_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
        for name, argtypes in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = _i
        _lib = l
    return _lib


_timing = None  # list of (name, start_event, end_event) while per-kernel timing is on (bench.py roofline)


def start_timing():
    global _timing
    _timing = []


def stop_timing():
    """-> {entry point: (launches, total milliseconds)} measured with CUDA events on the launching stream."""
    global _timing
    import torch
    torch.cuda.synchronize()
    out = {}
    for name, s, e in _timing or []:
        n, t = out.get(name, (0, 0.0))
        out[name] = (n + 1, t + s.elapsed_time(e))
    _timing = None
    return out


def call(name, *args):
    """Invoke a C-ABI entry point; raise RuntimeError (like the reference's c10::Error) on a non-zero status."""
    global launch_count
    l = lib()
    if _timing is not None:
        import torch
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
    rc = getattr(l, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {l.dvis_last_error().decode()}")
    if _timing is not None:
        e.record()
        _timing.append((name, s, e))
    launch_count += 1
