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


def test_bench_e2e_leg_merge():
    """bench.merge_e2e_legs: both legs reported; the headline switches only to a faster leg with agreeing results."""
    import bench
    base = {"value": 590.04, "unit": "frames/s", "ms_per_step": 27.117, "pipeline": "first"}
    e = bench.merge_e2e_legs(dict(base), 16, 27.117, 20.5, 0.0)
    assert e["ms_per_step"] == 20.5 and abs(e["value"] - 16 / 20.5e-3) < 0.01 and "3 clips" in e["pipeline"]
    assert list(e["legs_ms_per_step"].values()) == [27.117, 20.5]
    slower = bench.merge_e2e_legs(dict(base), 16, 27.117, 30.0, 0.0)
    assert slower["value"] == 590.04 and slower["pipeline"] == "first" and slower["legs_ms_per_step"]
    wrong = bench.merge_e2e_legs(dict(base), 16, 27.117, 20.5, 0.5)
    assert wrong["value"] == 590.04 and wrong["overlapped_leg_rel_max_diff_vs_first_leg"] == 0.5


def test_bench_round_robin_leg_merge():
    """bench.merge_round_robin_leg: times always reported; value / e2e switch only where faster and results agree."""
    import bench

    def base():
        return {"value": 1833.0, "ms_per_step": 8.73, "config": {"parallelism": "8 GPUs", "execution": "2 graphs"},
                "e2e": {"value": 1500.0, "ms_per_step": 10.67, "pipeline": "first", "legs_ms_per_step": {"a": 10.67}}}
    ln = bench.merge_round_robin_leg(base(), 16, 3.2, 4.0, 1e-4)
    assert ln["ms_per_step"] == 3.2 and ln["value"] == 5000.0 and "round-robin" in ln["config"]["parallelism"]
    assert ln["e2e"]["ms_per_step"] == 4.0 and ln["e2e"]["value"] == 4000.0 and len(ln["e2e"]["legs_ms_per_step"]) == 2
    assert ln["temporal_stage_legs_ms_per_step"]["replicated on every rank"] == 8.73
    mixed = bench.merge_round_robin_leg(base(), 16, 3.2, 12.0, 0.0)
    assert mixed["value"] == 5000.0 and mixed["e2e"]["value"] == 1500.0 and mixed["e2e"]["pipeline"] == "first"
    for bad in (0.3, float("nan")):
        ln = bench.merge_round_robin_leg(base(), 16, 3.2, 4.0, bad)
        assert ln["value"] == 1833.0 and ln["e2e"]["value"] == 1500.0 and ln["config"]["parallelism"] == "8 GPUs"
        assert ln["temporal_stage_legs_ms_per_step"]["owned round-robin per clip + 1 broadcast"] == 3.2


def test_bench_sm_carveout_leg_merge():
    import bench

    def base():
        return {"value": 792.0, "ms_per_step": 20.2, "config": {"execution": "2 graphs"}}
    ln = bench.merge_sm_carveout_leg(base(), 16, 8, 18.0, 1e-3)
    assert ln["ms_per_step"] == 18.0 and abs(ln["value"] - 888.89) < 0.01 and "8 SMs" in ln["config"]["execution"]
    for ms, diff in ((21.0, 0.0), (18.0, 0.5), (18.0, float("inf"))):
        ln = bench.merge_sm_carveout_leg(base(), 16, 8, ms, diff)
        assert ln["value"] == 792.0 and ln["config"]["execution"] == "2 graphs" and ln["sm_carveout_legs"][0]["ms_per_step"] == ms
    ln = bench.merge_sm_carveout_leg(bench.merge_sm_carveout_leg(base(), 16, 8, 18.0, 0.0), 16, 16, 17.0, 0.0)
    assert ln["ms_per_step"] == 17.0 and ln["config"]["execution"].count("cuBLASLt") == 1 and "16 SMs" in ln["config"]["execution"]
    assert [l["sms_left_free_by_cublaslt"] for l in ln["sm_carveout_legs"]] == [8, 16]
