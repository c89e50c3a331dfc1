"""GPU parity of the MSDA backward kernels: golden fp64 gradients of the reference's own PyTorch path, the C oracle,
torch.autograd.gradcheck on the OPS/test.py:66-89 channel list, and the reference's CUDA kernel built for sm_100a."""
import numpy as np
import pytest
import torch

from oracle import c_oracle

pytestmark = pytest.mark.gpu


def lsi_of(shapes):
    return torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))


def _bwd(value, shapes, loc, attn, gout):
    from dvis_plus_b200 import ops
    c = lambda t: t.cuda()
    gv, gl, ga = ops.ms_deform_attn_backward(c(value), c(shapes), c(lsi_of(shapes)), c(loc), c(attn), c(gout), 128)
    return gv.cpu(), gl.cpu(), ga.cpu()


def test_backward_golden_fp64_and_fp32(golden):
    g = golden("msda_small.pt")
    gv, gl, ga = _bwd(g["value"], g["shapes"], g["loc"], g["attn"], g["grad_out"])
    assert torch.allclose(gv, g["grad_value"]) and torch.allclose(gl, g["grad_loc"]) and torch.allclose(ga, g["grad_attn"])
    gv, gl, ga = _bwd(g["value"].float(), g["shapes"], g["loc"].float(), g["attn"].float(), g["grad_out"].float())
    for a, b in ((gv, g["grad_value"]), (gl, g["grad_loc"]), (ga, g["grad_attn"])):
        assert (a.double() - b).abs().max() < 2e-4 * max(1.0, b.abs().max().item())


@pytest.mark.parametrize("D", [8, 16, 32, 64, 128, 30, 71, 1025, 2048, 3096])
def test_backward_vs_oracle(D):
    torch.manual_seed(D)
    shapes = torch.as_tensor([(6, 4), (3, 2)])
    S = int(shapes.prod(1).sum())
    N, M, Lq, L, P = 2, 2, 5, 2, 2
    dt = torch.float32 if D <= 128 and D % 8 == 0 else torch.float64
    value = (torch.rand(N, S, M, D) * 0.01).to(dt)
    loc = (torch.rand(N, Lq, M, L, P, 2) * 1.3 - 0.15).to(dt)
    attn = torch.rand(N, Lq, M, L, P) + 1e-5
    attn = (attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)).to(dt)
    gout = torch.randn(N, Lq, M * D).to(dt)
    gv, gl, ga = _bwd(value, shapes, loc, attn, gout)
    rv, rl, ra = c_oracle.msda_backward(value.numpy(), shapes.numpy(), lsi_of(shapes).numpy(), loc.numpy(), attn.numpy(), gout.numpy())
    tol = 1e-4 if dt == torch.float32 else 1e-10
    for a, b in ((gv, rv), (gl, rl), (ga, ra)):
        assert np.abs(a.numpy() - b).max() <= tol * max(1.0, np.abs(b).max())


@pytest.mark.parametrize("channels", [30, 32, 64, 71])
def test_gradcheck_like_reference(channels):
    """OPS/test.py:66-89 (check_gradient_numerical): shapes :24-28, im2col_step=2, fp64 gradcheck through the autograd
    Function that wraps our forward / backward."""
    from dvis_plus_b200.modules import MSDeformAttnFunction
    torch.manual_seed(3)
    N, M, Lq, L, P = 1, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long).cuda()
    lsi = lsi_of(shapes)
    S = int(shapes.prod(1).sum())
    value = (torch.rand(N, S, M, channels).cuda() * 0.01).double().requires_grad_()
    loc = torch.rand(N, Lq, M, L, P, 2).cuda().double().requires_grad_()
    attn = torch.rand(N, Lq, M, L, P).cuda() + 1e-5
    attn = (attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)).double().requires_grad_()
    assert torch.autograd.gradcheck(MSDeformAttnFunction.apply, (value, shapes, lsi, loc, attn, 2))


def test_backward_matches_reference_cuda_kernel():
    """Same inputs through the reference's own col2im kernels compiled for sm_100a (oracle/_ref); skipped if absent."""
    from oracle import ref_cuda_binding as refcuda
    if not refcuda.available():
        pytest.skip("oracle/_ref/libref_msda.so not built")
    torch.manual_seed(0)
    shapes = torch.as_tensor([(23, 40), (12, 20), (6, 10)]).cuda()
    lsi = lsi_of(shapes)
    S = int(shapes.prod(1).sum())
    N, M, D, L, P, Lq = 2, 8, 32, 3, 4, 500
    value = torch.randn(N, S, M, D, device="cuda")
    loc = torch.rand(N, Lq, M, L, P, 2, device="cuda") * 1.2 - 0.1
    attn = torch.rand(N, Lq, M, L, P, device="cuda").flatten(-2).softmax(-1).view(N, Lq, M, L, P)
    gout = torch.randn(N, Lq, M * D, device="cuda")
    from dvis_plus_b200 import ops
    ours_f = ops.ms_deform_attn_forward(value, shapes, lsi, loc, attn, 128)
    ref_f = refcuda.forward(value, shapes, lsi, loc, attn)
    assert (ours_f - ref_f).abs().max() < 1e-5
    ours = ops.ms_deform_attn_backward(value, shapes, lsi, loc, attn, gout, 128)
    ref = refcuda.backward(value, shapes, lsi, loc, attn, gout)
    for a, b in zip(ours, ref):
        assert (a - b).abs().max() < 1e-4 * max(1.0, b.abs().max().item())
