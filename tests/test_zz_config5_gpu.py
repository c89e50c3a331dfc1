"""BASELINE.json config 5 (DVIS-DAQ ViT-L: 1080p, Q=300) at FULL per-frame size on the GPU against the oracle port on the
host, on identical seeded inputs and weights: the pixel decoder with ViT-Adapter-L channels
(1024 at all four strides) on a 1088x1920 padded frame (S = 42 840 encoder tokens, HW = 130 560 mask pixels), the mask
GEMM at Q = 300 (two in-place query slices), and the fused post-processing at 1080p.  The clip-level part of config 5
(T = 32 over 8 GPUs, 4 frames each) only multiplies these per-frame launches; the DAQ tracker blocks are pinned to the
reference's golden vectors in tests/test_daq.py.  Sorted last on purpose: the largest cases run after everything else.
Tolerances as in tests/test_configs_gpu.py."""
import pytest
import torch

from dvis_plus_b200 import _lib, ops
from dvis_plus_b200.modules.precision import precision
from oracle import postprocess_port as pp
from oracle import torch_port as tp
from postproc_util import assert_masks_match

pytestmark = pytest.mark.gpu
H1080, W1080 = 1088, 1920          # 1080p padded to a multiple of 32


def rel_err(a, b):
    assert a.shape == b.shape, (a.shape, b.shape)
    return (a.double().cpu() - b.double()).abs().max().item() / max(1e-6, b.abs().max().item())


@torch.no_grad()
def test_config5_vitl_pixel_decoder_1080p():
    import bench
    runner = bench.build_models("cpu", queries=100, backbone="vitl")
    feats = {k: v.float().contiguous() for k, v in
             bench.synthetic_features(1, "vitl", seed=5, dtype=torch.float32, hw=(H1080, W1080)).items()}
    sd = {k: v.detach().float().cpu() for k, v in runner.pixel_decoder.state_dict().items()}
    ref_mf, ref_o0, ref_ms = tp.pixel_decoder_forward_features(sd, feats, num_layers=6)
    pd = runner.pixel_decoder.cuda()
    n0 = _lib.launch_count
    with precision("bf16"):
        mf, o0, ms = pd.forward_features({k: v.cuda() for k, v in feats.items()})
    assert _lib.launch_count - n0 > 20, "libdvis_b200 kernels did not run"
    assert mf.shape == (1, 256, 272, 480) and [tuple(m.shape[-2:]) for m in ms] == [(34, 60), (68, 120), (136, 240)]
    assert rel_err(mf.float(), ref_mf) < 2e-2
    assert rel_err(o0.float(), ref_o0) < 2e-2
    for a, b in zip(ms, ref_ms):
        assert rel_err(a.float(), b) < 2e-2


@torch.no_grad()
def test_config5_mask_logits_q300_1080p():
    g = torch.Generator().manual_seed(9)
    emb = torch.randn(1, 300, 256, generator=g)
    feat = torch.randn(1, 256, 272, 480, generator=g)
    eb, fb = emb.bfloat16().float(), feat.bfloat16().float()               # the GEMM's operands are bf16
    ref = torch.einsum("bqc,bchw->bqhw", eb, fb)
    out = ops.mask_logits(emb.cuda(), feat.cuda().to(torch.bfloat16, memory_format=torch.channels_last), torch.float32)
    assert out.shape == (1, 300, 272, 480)
    assert rel_err(out, ref) < 1e-3
    out16 = ops.mask_logits(emb.cuda(), feat.cuda().to(torch.bfloat16, memory_format=torch.channels_last), torch.bfloat16)
    assert rel_err(out16.float(), ref) < 1e-2


@torch.no_grad()
def test_config5_postprocessing_1080p():
    gen = torch.Generator().manual_seed(2)
    Q, T, (h, w), first, img, out = 30, 4, (272, 480), (H1080, W1080), (1080, 1920), (1080, 1920)
    coarse = torch.randn(Q, T, h // 8, w // 8, generator=gen) * 6.0
    masks = torch.nn.functional.interpolate(coarse, size=(h, w), mode="bicubic", align_corners=False) - 1.0
    sel = torch.randperm(Q, generator=gen)[:20]
    d, dsel = masks.cuda(), sel.cuda()
    ours = ops.vis_masks(d, dsel, first, img, out)
    assert ours.shape == (20, T, 1080, 1920) and 0.02 < ours.float().mean().item() < 0.98
    ref = pp.resize_chain(masks[sel[:3]][:, [0, T - 1]], img, out[0], out[1], first)
    assert_masks_match(ours[:3][:, [0, T - 1]], ref > 0, ref, max_boundary_frac=1e-4)
    neg = ops.vis_masks(-d, dsel, first, img, out)
    assert (neg & ours).sum().item() == 0 and (~(neg | ours)).float().mean().item() < 1e-4
    assert torch.equal(ops.vis_masks(d[:, 1:3], dsel[5:9], first, img, out), ours[5:9, 1:3])


@torch.no_grad()
def test_daq_track_query_matching_on_device():
    """D/dvis_daq/track_module.py:749-759 on the GPU Hungarian kernel (ops.lap_rect) == SciPy on the host, at DAQ sizes
    (up to ~100 tracks x 100-200 segmenter queries, both orientations), and through VideoInstanceCutter.match_with_embeds."""
    import numpy as np
    from scipy.optimize import linear_sum_assignment
    from dvis_plus_b200 import modules as M
    g = torch.Generator().manual_seed(21)
    for rows, cols in ((37, 100), (100, 100), (150, 100), (1, 200), (300, 300)):
        cost = torch.rand(rows, cols, generator=g)
        got = ops.lap_rect(cost.cuda()).cpu().numpy()
        r, c = linear_sum_assignment(cost.numpy())
        ref = np.full(rows, -1, dtype=np.int64)
        ref[r] = c
        assert np.array_equal(got, ref), (rows, cols)
    batched = torch.rand(4, 20, 33, generator=g)
    got = ops.lap_rect(batched.cuda()).cpu()
    for b in range(4):
        assert np.array_equal(got[b].numpy(), linear_sum_assignment(batched[b].numpy())[1])
    cutter = M.VideoInstanceCutter(hidden_dim=64, feedforward_dim=128, num_head=8, decoder_layer_num=1, mask_dim=64, num_classes=5).cuda().eval()
    trc, seg = torch.randn(30, 1, 64, generator=g).cuda(), torch.randn(100, 1, 64, generator=g).cuda()
    host = cutter.match_with_embeds(trc, seg)
    cutter.match_on_host = False
    assert torch.equal(cutter.match_with_embeds(trc, seg), host)
