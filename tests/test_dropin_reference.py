"""The drop-in boundary against the UNMODIFIED reference sources (only where /root/reference exists, i.e. the build
container; skipped on the GPU box): the reference's own `OPS/functions/ms_deform_attn_func.py` imports our shim as
`MultiScaleDeformableAttention` (py:22) and its autograd Function reaches our entry points."""
import importlib
import os
import sys

import pytest
import torch

REF = "/root/reference/DVIS_Plus"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


@pytest.fixture()
def reference_func_module():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import reference_loader as rl
    from dvis_plus_b200 import ops
    saved = sys.modules.pop("MultiScaleDeformableAttention", None)
    shim = ops.install_as_reference_extension()
    rl.install()
    name = "mask2former.modeling.pixel_decoder.ops.functions.ms_deform_attn_func"
    sys.modules.pop(name, None)
    mod = importlib.import_module(name)
    yield mod, shim
    sys.modules.pop(name, None)
    if saved is not None:
        sys.modules["MultiScaleDeformableAttention"] = saved


def test_reference_function_binds_to_our_extension(reference_func_module):
    mod, shim = reference_func_module
    assert mod.MSDA is shim                                           # `import MultiScaleDeformableAttention as MSDA`
    assert callable(mod.MSDA.ms_deform_attn_forward) and callable(mod.MSDA.ms_deform_attn_backward)
    calls = []
    orig = shim.ms_deform_attn_forward
    shim.ms_deform_attn_forward = lambda *a: (calls.append(len(a)), orig(*a))[1]
    value = torch.rand(1, 30, 2, 4)
    shapes = torch.as_tensor([(6, 4), (3, 2)])
    lsi = torch.as_tensor([0, 24])
    loc, attn = torch.rand(1, 2, 2, 2, 2, 2), torch.rand(1, 2, 2, 2, 2)
    # the reference's autograd Function calls OUR forward with the reference's 6 positional arguments; on CPU tensors it
    # raises exactly what the reference extension raises (OPS/src/ms_deform_attn.h:43)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        mod.MSDeformAttnFunction.apply(value, shapes, lsi, loc, attn, 128)
    assert calls == [6]


def test_reference_cpu_fallback_still_works_next_to_the_shim(reference_func_module):
    mod, _ = reference_func_module
    value = torch.rand(1, 30, 2, 4)
    out = mod.ms_deform_attn_core_pytorch(value, torch.as_tensor([(6, 4), (3, 2)]), torch.rand(1, 2, 2, 2, 2, 2),
                                          torch.rand(1, 2, 2, 2, 2))
    assert out.shape == (1, 2, 8)
