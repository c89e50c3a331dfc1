"""Non-finite inputs (NaN, +-inf, 1e30) through every CUDA-core entry point on the SIMT emulator: nothing may index out of
range or hang, whatever the values.  Most useful under tests/simt/run_sanitized.sh address (every load / store is
bounds-checked there); in the plain suite it catches crashes and dead-locks.  This fuzz found nothing in the kernels after
the top-k fix (a NaN score used to produce an out-of-range index there)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt"))
import simt_binding as simt  # noqa: E402

pytestmark = pytest.mark.timeout(900)
BAD = [float("nan"), float("inf"), -float("inf"), 1e30, -1e30, 3e38]


def test_nonfinite_inputs_never_leave_the_buffers():
    g = torch.Generator().manual_seed(0)

    def poison(t, frac=0.05):
        t = t.clone()
        m = torch.rand(t.shape, generator=g) < frac
        vals = torch.tensor(BAD)[torch.randint(0, len(BAD), t.shape, generator=g)]
        t[m] = vals[m].to(t.dtype)
        return t

    sh, lsi, S = torch.tensor([(6, 9), (3, 5)]), torch.tensor([0, 54]), 69
    for trial in range(2):
        value = torch.randn(2, S, 4, 32, generator=g)
        loc = poison(torch.rand(2, 11, 4, 2, 4, 2, generator=g) * 1.4 - 0.2, 0.2)
        attn = poison(torch.rand(2, 11, 4, 2, 4, generator=g), 0.1)
        simt.msda_forward(poison(value, 0.01), sh, lsi, loc, attn)
        simt.msda_backward(value, sh, lsi, loc, attn, poison(torch.randn(2, 11, 128, generator=g), 0.05))
        fused = poison(torch.randn(2, S, 4 * 2 * 4 * 3, generator=g) * 3, 0.1)
        ref = poison(torch.rand(2, S, 2, 2, generator=g), 0.1)
        simt.msda_fused_forward(value, sh, lsi, fused[..., :64], fused[..., 64:], ref, 2, 4)
        simt.msda_fused_forward(value.bfloat16(), sh, lsi, fused[..., :64].bfloat16(), fused[..., 64:].bfloat16(), ref, 2, 4, pair=True)
        m = poison(torch.randn(5, 2, 7, 9, generator=g) * 3, 0.1)
        for out in ((25, 33), (40, 51), (9, 11)):
            simt.vis_masks(m, torch.tensor([4, 0]), (28, 36), (25, 33), out)
            simt.vis_masks_packed(m.bfloat16(), None, (28, 36), (25, 33), out)
            win, _ = simt.vps_argmax(m, torch.tensor([1, 3]), poison(torch.rand(2, generator=g), 0.3), (28, 36), (25, 33), out)
            assert int(torch.where(win >= 0, win, ~win).max()) < 2
            lab = simt.vss_argmax(m, poison(torch.rand(5, 4, generator=g), 0.2), (28, 36), (25, 33), out)
            assert 0 <= int(lab.min()) and int(lab.max()) < 4
        cls = poison(torch.randn(9, 5, generator=g), 0.2)
        simt.class_scores(cls, poison(torch.randn(9, 5, generator=g), 0.2))
        _, l, q = simt.vis_topk(cls, 12)
        assert 0 <= int(q.min()) and int(q.max()) < 9 and 0 <= int(l.min()) and int(l.max()) < 4
        simt.add_layernorm(poison(torch.randn(7, 128, generator=g)), poison(torch.randn(7, 128, generator=g)), torch.randn(128),
                           torch.randn(128), lp_dtype=torch.bfloat16)
        simt.groupnorm_nhwc(poison(torch.randn(2, 24, 128, generator=g)), 32, torch.randn(128), torch.randn(128))
        simt.attn_bias_from_logits(poison(torch.randn(3, 4, 30, generator=g), 0.2))
        qkv = poison(torch.randn(1, 13, 3, 2, 32, generator=g)).bfloat16()
        simt.flash_attn(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], 0.2)
    # the Hungarian kernels: NaN counts as 0 (reference semantics); inf / huge costs terminate with in-range assignments
    for bad in (float("nan"), 1e30, float("inf")):
        c = torch.rand(2, 9, 9, generator=g)
        c[0, 2, :] = bad
        c[1, :, 4] = bad
        s, _ = simt.lap_chain(c)
        assert int(s.min()) >= 0 and int(s.max()) < 9 and all(len(set(row.tolist())) == 9 for row in s)
        r = simt.lap_rect(torch.where(torch.rand(5, 8, generator=g) < 0.3, torch.tensor(bad), torch.rand(5, 8, generator=g)))
        assert int(r.max()) < 8 and len(set(r.tolist())) == 5


def test_msda_plain_op_propagates_nonfinite_values_like_the_reference():
    """The plain op skips corners outside the map (ms_deform_im2col_cuda.cuh:61-83) instead of multiplying a redirected row by 0:
    a NaN / Inf in `value` -- also at pixel 0 of a head, where round 1's kernel leaked it into every fully-outside point --
    reaches exactly the outputs it reaches in the reference (oracle/msda_oracle.c restates the reference's conditionals)."""
    import numpy as np
    from oracle import c_oracle
    g = torch.Generator().manual_seed(4)
    sh, lsi, S = torch.tensor([(6, 9), (3, 5)]), torch.tensor([0, 54]), 69
    for D in (32, 8):
        value = torch.randn(2, S, 4, D, generator=g)
        value[:, 0] = float("nan")                    # pixel 0 of every head (the redirect target of round 1's kernel)
        value[0, 17, 1, 3] = float("inf")
        value[1, 60, 2] = float("nan")
        loc = torch.rand(2, 23, 4, 2, 4, 2, generator=g) * 1.6 - 0.3        # many points partly or fully outside the maps
        attn = torch.rand(2, 23, 4, 2, 4, generator=g)
        out = simt.msda_forward(value, sh, lsi, loc, attn).numpy()
        ref = c_oracle.msda_forward(value.numpy(), sh.numpy(), lsi.numpy(), loc.numpy(), attn.numpy())
        assert np.array_equal(np.isnan(out), np.isnan(ref)) and np.array_equal(np.isinf(out), np.isinf(ref))
        assert np.isnan(ref).any() and (~np.isnan(ref)).any()
        fin = np.isfinite(ref)
        assert np.abs(out[fin] - ref[fin]).max() <= 2e-5 * max(1.0, np.abs(ref[fin]).max())
