"""Arithmetic policy of the module-level drop-ins.

"fp32": every GEMM in fp32 on CUDA cores (tight parity with the reference's forced-fp32 pixel decoder,
        P/mask2former/modeling/pixel_decoder/msdeformattn.py:314,320); mask logits still run on the bf16
        tensor path (the reference evaluates that einsum in fp16 under autocast, P/train_net_video.py:259).
"bf16": GEMM / conv inputs rounded to bf16, fp32 accumulation, LayerNorm / softmax / residual stream in fp32.
"""
import contextlib

import torch

_state = {"mode": "bf16"}


def set_precision(mode: str):
    assert mode in ("fp32", "bf16"), mode
    _state["mode"] = mode


def get_precision() -> str:
    return _state["mode"]


@contextlib.contextmanager
def precision(mode: str):
    old = _state["mode"]
    set_precision(mode)
    try:
        yield
    finally:
        _state["mode"] = old


def gemm_dtype():
    return torch.bfloat16 if _state["mode"] == "bf16" else torch.float32
