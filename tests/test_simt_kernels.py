"""The CUDA-core kernels of csrc/postproc.cu and csrc/lap.cu, executed from their ORIGINAL sources by the SIMT emulator of
tests/simt/ through the same C-ABI entry points libdvis_b200.so exports, against the oracle / SciPy -- on CPU.

This covers what the host-compiled per-pixel core (test_postproc_hostcore.py) cannot: launch geometry, thread -> pixel
mapping, guards, vector / tail stores, shared-memory counters, warp collectives and the entry points' argument handling.
Test infrastructure only; the -m gpu tests (test_postprocess_gpu.py) run the same cases on the device."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import postprocess_port as pp
from postproc_util import assert_labels_match, assert_masks_match

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt"))
import simt_binding as simt  # noqa: E402

pytestmark = pytest.mark.timeout(900)          # an emulated kernel that dead-locks must fail, not hang the suite

GEOMS = [  # (h, w), first resize, image size, output size
    ((12, 20), (48, 80), (45, 78), (45, 78)),      # identity second resize: strip kernel, byte-store tail
    ((12, 20), (48, 80), (48, 80), (48, 80)),      # strip kernel, 8-byte vector stores
    ((7, 9), (28, 36), (25, 33), (25, 33)),        # odd sizes
    ((46, 80), (184, 320), (180, 320), (180, 320)),  # several row bands, strips and CTAs per plane
    ((12, 20), (48, 80), (45, 78), (67, 117)),     # two-stage kernel, byte-store tail
    ((12, 20), (48, 80), (45, 78), (30, 52)),      # two-stage, 4-byte vector stores
    ((23, 40), (92, 160), (90, 160), (180, 320)),  # exact 2x second resize, several CTAs
    ((12, 20), (48, 80), (45, 78), (200, 301)),    # strong up-scale
    ((30, 30), (120, 120), (118, 119), (17, 13)),  # strong down-scale
    ((5, 6), (20, 24), (20, 24), (3, 2)),          # output smaller than the logits
]


def test_class_scores_and_topk():
    g = torch.Generator().manual_seed(0)
    for Q, K, max_num, use_aux in ((50, 25, 10, True), (7, 3, 21, False), (40, 124, 30, True)):
        cls = torch.randn(Q, K + 1, generator=g) * 3
        aux = torch.randn(Q, K + 1, generator=g) * 3 if use_aux else None
        ref = pp.vis_scores(cls, aux)
        sc = simt.class_scores(cls, aux)
        torch.testing.assert_close(sc[:, :-1], ref, rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(sc[:, -1], cls.softmax(-1)[:, -1], rtol=1e-5, atol=1e-7)
        s, l, q = simt.vis_topk(cls, max_num, aux)
        rs, ri = ref.flatten().topk(max_num, sorted=True)
        torch.testing.assert_close(s, rs, rtol=1e-5, atol=1e-7)
        assert torch.equal(q * K + l, ri)
    # ties: lower flat index first
    cls = torch.zeros(4, 3)
    s, l, q = simt.vis_topk(cls, 5)
    assert (q * 2 + l).tolist() == [0, 1, 2, 3, 4]
    with pytest.raises(RuntimeError, match="out of range"):
        simt.vis_topk(torch.randn(3, 4), 10)
    # non-finite logits: NaN scores rank above everything (torch.topk's convention), indices always stay in range
    cls = torch.randn(6, 4, generator=g)
    cls[2, 1] = float("nan")                                            # the whole softmax row 2 becomes NaN
    s, l, q = simt.vis_topk(cls, 7)
    assert (q[:3] == 2).all() and l[:3].tolist() == [0, 1, 2] and torch.isnan(s[:3]).all()
    rest = pp.vis_scores(cls).flatten()
    rest[torch.isnan(rest)] = -1
    rs, ri = rest.topk(4, sorted=True)
    assert torch.equal(q[3:] * 3 + l[3:], ri) and torch.allclose(s[3:], rs)
    s, l, q = simt.vis_topk(torch.full((5, 3), float("nan")), 10)        # everything NaN: first 10 flat indices, in order
    assert (q * 2 + l).tolist() == list(range(10))


@pytest.mark.parametrize("geom", GEOMS)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_vis_masks_kernels(geom, dtype):
    (h, w), first, img, out = geom
    g = torch.Generator().manual_seed(h * 100 + w)
    masks = (torch.randn(5, 2, h, w, generator=g) * 3).to(dtype)
    sel = torch.tensor([4, 0, 4], dtype=torch.int64)
    ours = simt.vis_masks(masks, sel, first, img, out)
    ref = pp.resize_chain(masks[sel].float(), img, out[0], out[1], first)
    assert_masks_match(ours, ref > 0, ref, tol=2e-5, max_boundary_frac=1e-3)
    fm = masks.transpose(0, 1).contiguous().transpose(0, 1)             # frame-major storage, no selection
    ours2 = simt.vis_masks(fm, None, first, img, out)
    ref2 = pp.resize_chain(masks.float(), img, out[0], out[1], first)
    assert_masks_match(ours2, ref2 > 0, ref2, tol=2e-5, max_boundary_frac=1e-3)


@pytest.mark.parametrize("geom", GEOMS)
def test_vis_masks_packed_kernels(geom):
    """One bit per pixel: identical to packing the byte kernel's result (little bit order, zero padding bits)."""
    (h, w), first, img, out = geom
    g = torch.Generator().manual_seed(h * 10 + w)
    masks = torch.randn(4, 2, h, w, generator=g) * 3
    sel = torch.tensor([3, 1], dtype=torch.int64)
    ref = simt.vis_masks(masks, sel, first, img, out)
    packed = simt.vis_masks_packed(masks, sel, first, img, out)
    assert packed.shape == (2, 2, out[0], (out[1] + 7) // 8)
    assert np.array_equal(packed.numpy(), np.packbits(ref.numpy(), axis=-1, bitorder="little"))
    from dvis_plus_b200 import ops
    assert torch.equal(ops.unpack_masks(packed, out[1]), ref)


@pytest.mark.parametrize("geom", GEOMS)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_vis_masks_tiled_variant_is_bit_identical(geom, dtype, monkeypatch):
    """DVIS_VIS_MASKS_TILED=1: the CTA stages its source window in shared memory and runs the same walkers on it --
    byte and packed results equal the default strip kernels' exactly (strongly down-scaling chains fall back to them)."""
    (h, w), first, img, out = geom
    g = torch.Generator().manual_seed(h * 7 + w)
    masks = (torch.randn(4, 2, h, w, generator=g) * 3).to(dtype)
    sel = torch.tensor([3, 0, 3], dtype=torch.int64)
    fm = masks.transpose(0, 1).contiguous().transpose(0, 1)
    ref, ref_p, ref_fm = simt.vis_masks(masks, sel, first, img, out), simt.vis_masks_packed(masks, sel, first, img, out), \
        simt.vis_masks(fm, None, first, img, out)
    monkeypatch.setenv("DVIS_VIS_MASKS_TILED", "1")
    assert torch.equal(simt.vis_masks(masks, sel, first, img, out), ref)
    assert torch.equal(simt.vis_masks_packed(masks, sel, first, img, out), ref_p)
    assert torch.equal(simt.vis_masks(fm, None, first, img, out), ref_fm)


def test_vis_masks_tiled_variant_falls_back_when_the_window_is_too_large(monkeypatch):
    """A strongly down-scaling chain over a large source: one tile's source window (120 x 216 floats) exceeds 48 KB of
    shared memory, the launcher keeps the strip kernel."""
    h, w, first, img, out = 120, 216, (480, 864), (480, 854), (60, 107)
    masks = torch.randn(1, 1, h, w, generator=torch.Generator().manual_seed(2)) * 3
    ref = simt.vis_masks(masks, None, first, img, out)
    monkeypatch.setenv("DVIS_VIS_MASKS_TILED", "1")
    assert torch.equal(simt.vis_masks(masks, None, first, img, out), ref)
    chain = pp.resize_chain(masks, img, out[0], out[1], first)
    assert_masks_match(ref, chain > 0, chain, tol=2e-5, max_boundary_frac=1e-3)


def test_vis_masks_argument_checks():
    m = torch.randn(2, 2, 4, 4)
    with pytest.raises(RuntimeError, match="crop"):
        simt.vis_masks(m, None, (16, 16), (17, 16), (17, 16))
    with pytest.raises(RuntimeError, match="65535"):
        simt.call("dvis_vis_masks", m.data_ptr(), 0, 32, 16, None, 40000, 2, 4, 4, 16, 16, 16, 16, 16, 16, m.data_ptr(), None)
    with pytest.raises(RuntimeError, match="f32 or bf16"):
        simt.call("dvis_vis_masks", m.data_ptr(), 1, 32, 16, None, 2, 2, 4, 4, 16, 16, 16, 16, 16, 16, m.data_ptr(), None)


@pytest.mark.parametrize("geom", [GEOMS[0], GEOMS[2], GEOMS[4], GEOMS[5]])
def test_vps_kernels(geom):
    (h, w), first, img, out = geom
    g = torch.Generator().manual_seed(7 + h)
    masks = torch.randn(9, 2, h, w, generator=g) * 3
    keep_idx = torch.tensor([1, 3, 4, 8], dtype=torch.int64)
    keep_score = torch.tensor([0.9, 0.5, 0.7, 0.95])
    win, areas = simt.vps_argmax(masks, keep_idx, keep_score, first, img, out)
    cur = pp.resize_chain(masks[keep_idx], img, out[0], out[1], first, sigmoid=True)
    prob = keep_score.view(-1, 1, 1, 1) * cur
    ref_ids = prob.argmax(0)
    ids = torch.where(win >= 0, win, ~win).long()
    assert_labels_match(ids, ref_ids, prob, tol=1e-5)
    n = keep_idx.numel()
    ref_areas = torch.stack([torch.stack([(ref_ids == k).sum() for k in range(n)]),
                             torch.stack([(cur[k] >= 0.5).sum() for k in range(n)]),
                             torch.stack([((ref_ids == k) & (cur[k] >= 0.5)).sum() for k in range(n)])])
    assert (areas - ref_areas).abs().max().item() <= 3, (areas, ref_areas)
    assert areas[0].sum().item() == ref_ids.numel()                     # every pixel counted exactly once
    seg = torch.tensor([5, 0, 7, 7], dtype=torch.int32)
    pan = simt.vps_paint(win, seg)
    assert torch.equal(pan, torch.where(win >= 0, seg[win.clamp(min=0).long()], torch.zeros_like(win)))


@pytest.mark.parametrize("geom", [GEOMS[0], GEOMS[4]])
@pytest.mark.parametrize("K", [5, 19])
def test_vss_kernel(geom, K):
    (h, w), first, img, out = geom
    g = torch.Generator().manual_seed(3 + K)
    masks = torch.randn(10, 2, h, w, generator=g) * 3
    cls = torch.randn(10, K + 1, generator=g) * 2
    scores = simt.class_scores(cls)
    ours = simt.vss_argmax(masks, scores[:, :-1], first, img, out)
    ref = pp.inference_video_vss(cls, masks, img, out[0], out[1], first, return_scores=True)
    assert_labels_match(ours, ref["pred_masks"], ref["semseg"], tol=1e-5)


def test_lap_chain_kernel_vs_scipy():
    from scipy.optimize import linear_sum_assignment
    g = torch.Generator().manual_seed(4)
    for T, n in ((3, 5), (2, 33), (1, 45)):
        cost = torch.rand(T, n, n, generator=g)
        cost[0, 1, 2] = float("nan")                                    # NaN counts as 0 (noiser.py:52)
        init = torch.randperm(n, generator=g)
        sigma, idx = simt.lap_chain(cost, init)
        c = torch.where(torch.isnan(cost), torch.zeros_like(cost), cost).numpy()
        prev = init.numpy()
        for t in range(T):
            s = linear_sum_assignment(c[t])[1]
            assert np.array_equal(sigma[t].numpy(), s)
            prev = s[prev]
            assert np.array_equal(idx[t].numpy(), prev)


@pytest.mark.parametrize("jitter", [0, 5])
def test_lap_rect_kernel_vs_scipy(jitter):
    """Rectangular assignment (DVIS-DAQ track x query matching): rows < cols, rows > cols (solved through the transpose,
    unmatched rows = -1), square, single row / column; batched."""
    from scipy.optimize import linear_sum_assignment
    g = torch.Generator().manual_seed(12)
    simt.set_jitter(jitter)
    try:
        for B, rows, cols in ((2, 7, 19), (2, 19, 7), (1, 12, 12), (1, 1, 9), (1, 9, 1), (1, 40, 70)):
            cost = torch.rand(B, rows, cols, generator=g)
            got = simt.lap_rect(cost)
            for b in range(B):
                r, c = linear_sum_assignment(cost[b].numpy())
                ref = np.full(rows, -1, dtype=np.int64)
                ref[r] = c
                assert np.array_equal(got[b].numpy(), ref), (rows, cols)
        assert simt.lap_rect(torch.rand(5, 8, generator=g)).shape == (5,)
    finally:
        simt.set_jitter(0)
    with pytest.raises(RuntimeError, match="1024"):
        simt.lap_rect(torch.zeros(1, 2, 2000))


def test_barrier_protocols_under_jitter():
    """Race shaker: the kernels that communicate through shared memory (block arg-max of the top-k, vps area counters,
    the Hungarian kernel) give the same results when random threads are delayed after every __syncthreads()."""
    from scipy.optimize import linear_sum_assignment
    simt.set_jitter(5)
    try:
        g = torch.Generator().manual_seed(8)
        cost = torch.rand(2, 40, 40, generator=g)
        sigma, _ = simt.lap_chain(cost)
        for t in range(2):
            assert np.array_equal(sigma[t].numpy(), linear_sum_assignment(cost[t].numpy())[1])
        cls = torch.randn(30, 11, generator=g) * 3
        s, l, q = simt.vis_topk(cls, 12)
        rs, ri = pp.vis_scores(cls).flatten().topk(12, sorted=True)
        assert torch.equal(q * 10 + l, ri)
        masks = torch.randn(6, 1, 7, 9, generator=g) * 3
        keep_idx, keep_score = torch.tensor([0, 2, 5]), torch.tensor([0.9, 0.6, 0.8])
        win, areas = simt.vps_argmax(masks, keep_idx, keep_score, (28, 36), (25, 33), (25, 33))
        assert areas[0].sum().item() == 25 * 33 and (areas[2] <= areas[0]).all() and (areas[2] <= areas[1]).all()
    finally:
        simt.set_jitter(0)
