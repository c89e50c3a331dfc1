"""Arithmetic policy of the module-level drop-ins.

"fp32": every GEMM in fp32 on CUDA cores (tight parity with the reference's forced-fp32 pixel decoder,
        P/mask2former/modeling/pixel_decoder/msdeformattn.py:314,320); the mask-head GEMMs take fp32 operands and multiply
        them as TF32 on the tensor cores (csrc/mask_gemm.cu kTf32: 1e-3 of the output scale).
"bf16": GEMM / conv inputs rounded to bf16, fp32 accumulation, LayerNorm / softmax / residual stream in fp32.
"""
import contextlib

import torch

_state = {"mode": "bf16"}


def set_precision(mode: str):
    assert mode in ("fp32", "bf16"), mode
    _state["mode"] = mode
    from .. import ops
    ops.set_mask_operand_dtype(torch.float32 if mode == "fp32" else torch.bfloat16)


def get_precision() -> str:
    return _state["mode"]


@contextlib.contextmanager
def precision(mode: str):
    old = _state["mode"]
    set_precision(mode)
    try:
        yield
    finally:
        set_precision(old)


def gemm_dtype():
    return torch.bfloat16 if _state["mode"] == "bf16" else torch.float32
