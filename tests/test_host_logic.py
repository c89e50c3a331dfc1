"""Host-side helpers that need no GPU: locality schedule, precision policy, bench helpers."""
import numpy as np
import torch

from dvis_plus_b200 import locality
from dvis_plus_b200.modules import precision as P


def test_tiled_item_order_is_a_permutation_and_tile_major():
    shapes, M = ((23, 40), (12, 20), (6, 10)), 8
    order = locality._tiled_order_np(shapes, M, 16, 16)
    S = sum(h * w for h, w in shapes)
    assert order.dtype == np.int32 and order.size == S * M
    assert np.array_equal(np.sort(order), np.arange(S * M))
    # first chunk: head 0 of the first 16x16 tile of level 0, row-major inside the tile
    q, m = order[:256] // M, order[:256] % M
    assert (m == 0).all()
    ys, xs = q // 40, q % 40
    assert ys.max() == 15 and xs.max() == 15 and (np.diff(q[:16]) == 1).all()
    # every tile's heads are contiguous blocks
    assert (order[256:512] % M == 1).all()


def test_precision_policy_context():
    assert P.get_precision() in ("fp32", "bf16")
    before = P.get_precision()
    with P.precision("fp32"):
        assert P.gemm_dtype() == torch.float32
        with P.precision("bf16"):
            assert P.gemm_dtype() == torch.bfloat16
        assert P.gemm_dtype() == torch.float32
    assert P.get_precision() == before


def test_bench_helpers():
    import bench
    f = bench.synthetic_features(2, "swinl", hw=(64, 96))
    assert set(f) == {"res2", "res3", "res4", "res5"}
    assert f["res2"].shape == (2, 192, 16, 24) and f["res5"].shape == (2, 1536, 2, 3)
    assert f["res3"].dtype == torch.bfloat16 and f["res3"].is_contiguous(memory_format=torch.channels_last)
    g = bench.synthetic_features(2, "swinl", hw=(64, 96))
    assert torch.equal(f["res4"], g["res4"])                      # seeded
    s = bench.ClockSampler(0)
    s.proc, s.lines = object.__new__(type("P", (), {"terminate": lambda self: None, "wait": lambda self, timeout=None: 0,
                                                     "kill": lambda self: None})), [
        "0, 1965, 1965, 412.1, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active",
        "0, 1800, 1965, 800.0, 0x0000000000000004, Not Active, Not Active, Not Active, Active",
        "0, 1950, 1965, 500.0, 0x0, Not Active, Not Active, Not Active, Not Active"]
    out = s.stop()
    assert out["sm_mhz"] == 1950.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"] and out["samples"] == 3


def test_bench_rel_max_diff():
    """bench.rel_max_diff (the line's parity_check): max over tensors of max|a-b| / max|b|; NaN counts as a mismatch."""
    import bench
    a = {"x": torch.tensor([1.0, 2.0, 4.0]), "y": torch.tensor([10.0])}
    assert bench.rel_max_diff(a, {k: v.clone() for k, v in a.items()}) == 0.0
    b = {"x": torch.tensor([1.0, 2.0, 5.0]), "y": torch.tensor([10.0])}
    assert abs(bench.rel_max_diff(a, b) - 0.2) < 1e-6
    c = {"x": torch.tensor([1.0, float("nan"), 4.0]), "y": torch.tensor([10.0])}
    assert bench.rel_max_diff(c, a) == float("inf")
