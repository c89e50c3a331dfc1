"""Oracle self-consistency on random shapes (CPU): the plain-C restatement of the CUDA kernel arithmetic and the
torch restatement of the reference's grid_sample path must agree with each other (they follow different reference files:
ms_deform_im2col_cuda.cuh vs ms_deform_attn_func.py:52-72), forward and backward; plus repository rules as tests."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import c_oracle, torch_port as tp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def lsi_of(shapes):
    return torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))


@pytest.mark.parametrize("seed", range(8))
def test_c_oracle_equals_torch_port_on_random_shapes(seed):
    g = torch.Generator().manual_seed(seed)
    L = int(torch.randint(1, 5, (1,), generator=g))
    shapes = torch.randint(1, 9, (L, 2), generator=g)
    N, M, D = (int(torch.randint(1, 4, (1,), generator=g)) for _ in range(3))
    D = [1, 3, 8, 32][seed % 4]
    Lq, P = int(torch.randint(1, 20, (1,), generator=g)), int(torch.randint(1, 5, (1,), generator=g))
    S = int(shapes.prod(1).sum())
    value = torch.randn(N, S, M, D, dtype=torch.float64, generator=g)
    loc = torch.rand(N, Lq, M, L, P, 2, dtype=torch.float64, generator=g) * 1.6 - 0.3
    attn = torch.rand(N, Lq, M, L, P, dtype=torch.float64, generator=g)
    ref = tp.msda_core(value, shapes.tolist(), loc, attn)
    out = c_oracle.msda_forward(value.numpy(), shapes.numpy(), lsi_of(shapes).numpy(), loc.numpy(), attn.numpy())
    assert np.allclose(out, ref.numpy(), rtol=1e-10, atol=1e-12)
    # backward: C restatement of col2im vs autograd through the grid_sample formulation
    value.requires_grad_(); loc.requires_grad_(); attn.requires_grad_()
    o = tp.msda_core(value, shapes.tolist(), loc, attn)
    go = torch.randn(o.shape, dtype=torch.float64, generator=g)
    o.backward(go)
    gv, gl, ga = c_oracle.msda_backward(value.detach().numpy(), shapes.numpy(), lsi_of(shapes).numpy(), loc.detach().numpy(),
                                        attn.detach().numpy(), go.numpy())
    assert np.allclose(gv, value.grad.numpy(), rtol=1e-8, atol=1e-10)
    assert np.allclose(ga, attn.grad.numpy(), rtol=1e-8, atol=1e-10)
    assert np.allclose(gl, loc.grad.numpy(), rtol=1e-7, atol=1e-9)


def test_mask_oracle_equals_einsum():
    torch.manual_seed(0)
    emb, feat = torch.randn(2, 7, 16), torch.randn(2, 16, 5, 6)
    ref = torch.einsum("bqc,bchw->bqhw", emb.double(), feat.double()).float()
    assert np.allclose(c_oracle.mask_logits(emb.numpy(), feat.numpy()), ref.numpy(), rtol=1e-6, atol=1e-6)


def _py_files(*dirs):
    for d in dirs:
        for base, _, files in os.walk(os.path.join(ROOT, d)):
            for f in files:
                if f.endswith(".py"):
                    yield os.path.join(base, f)


def test_product_never_imports_the_oracle_or_the_reference():
    """The oracle is test infrastructure: nothing under dvis_plus_b200/ may import it (or read /root/reference)."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|/root/reference", re.M)
    offenders = [p for p in _py_files("dvis_plus_b200") if pat.search(open(p).read())]
    assert not offenders, offenders
    # and bench.py touches the oracle only inside its CPU-baseline / reference-arm functions
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count("from oracle") == 1 and "def cpu_reference_step" in src.split("from oracle")[0].splitlines()[-2]


def test_product_has_no_host_implementation_of_the_postprocessing_kernels():
    """tests/hostcore compiles csrc/resize_core.cuh for the host as a CHECKER; the product must not: no python file of the
    package mentions it and libdvis_b200.so exports none of its symbols."""
    import ctypes
    from dvis_plus_b200 import _lib
    assert not [p for p in _py_files("dvis_plus_b200") if "hostcore" in open(p).read()]
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in ("hostcore_vis_masks", "hostcore_vps_argmax", "hostcore_vss_argmax"):
        assert not hasattr(lib, name)


def test_gpu_tests_do_not_read_the_reference_tree():
    """/root/reference does not exist on the GPU box: only the fixture generators and the explicitly skipped drop-in test
    may mention it."""
    allowed = {"reference_loader.py", "make_golden.py", "make_golden_r2.py", "make_golden_postprocess.py", "make_golden_daq_runner.py", "test_dropin_reference.py", "test_oracle_properties.py"}
    bad = [p for p in _py_files("tests") if os.path.basename(p) not in allowed and "/root/reference" in open(p).read()]
    assert not bad, bad


def test_ops_reject_cpu_tensors():
    from dvis_plus_b200 import ops
    shapes = torch.as_tensor([(4, 4)])
    with pytest.raises(RuntimeError, match="CPU"):
        ops.ms_deform_attn_forward(torch.randn(1, 16, 2, 8), shapes, torch.zeros(1, dtype=torch.long), torch.rand(1, 3, 2, 1, 2, 2),
                                   torch.rand(1, 3, 2, 1, 2), 128)
    with pytest.raises(RuntimeError, match="CPU"):
        ops.ms_deform_attn_backward(torch.randn(1, 16, 2, 8), shapes, torch.zeros(1, dtype=torch.long), torch.rand(1, 3, 2, 1, 2, 2),
                                    torch.rand(1, 3, 2, 1, 2), torch.rand(1, 3, 16), 128)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.mask_logits(torch.randn(1, 4, 64), torch.randn(1, 64, 4, 4))
