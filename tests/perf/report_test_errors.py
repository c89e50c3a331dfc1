"""Runs the module- / config-level GPU parity tests with rel_err() instrumented and prints the errors they actually measure
(the asserted tolerances are upper bounds).  Writes gpurun_out/test_errors.json."""
import inspect
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402


def golden(name):
    return torch.load(os.path.join(ROOT, "tests", "golden", name), map_location="cpu", weights_only=False)


def run(mod, names):
    out = {}
    rec = []
    orig = mod.rel_err

    def rel_err(a, b):
        v = orig(a, b)
        rec.append(round(float(v), 6))
        return v
    mod.rel_err = rel_err
    for name, kwargs in names:
        fn = getattr(mod, name)
        rec.clear()
        params = inspect.signature(fn).parameters
        kw = dict(kwargs)
        if "golden" in params:
            kw["golden"] = golden
        try:
            fn(**kw)
            out[name + (str(kwargs) if kwargs else "")] = {"max": max(rec) if rec else None, "errors": list(rec)}
        except Exception as e:   # noqa: BLE001
            out[name + (str(kwargs) if kwargs else "")] = {"failed": repr(e)[:200], "errors": list(rec)}
        print(name, kwargs, out[name + (str(kwargs) if kwargs else "")], flush=True)
    mod.rel_err = orig
    return out


def main():
    import test_configs_gpu as tc
    import test_modules_gpu as tm
    res = {}
    res.update(run(tc, [("test_config2_r50_pixel_decoder_and_predictor_720p", {}), ("test_config3_online_tracker_T5_Q200", {}),
                        ("test_config4_offline_temporal_stage_T16_Q200", {})]))
    mods = [(n, {}) for n in dir(tm) if n.startswith("test_") and set(inspect.signature(getattr(tm, n)).parameters) <= {"golden"}]
    res.update(run(tm, mods))
    res.update(run(tm, [("test_predictor_golden", {"materialize": False}), ("test_predictor_golden", {"materialize": True}),
                        ("test_refiner_golden", {"mode": "bf16", "tol": 1.0}), ("test_refiner_golden", {"mode": "fp32", "tol": 1.0}),
                        ("test_tracker_golden", {"mode": "bf16", "tol": 1.0}), ("test_tracker_golden", {"mode": "fp32", "tol": 1.0})]))
    import test_zz_config5_gpu as t5
    res.update(run(t5, [(n, {}) for n in dir(t5) if n.startswith("test_") and not inspect.signature(getattr(t5, n)).parameters]))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "test_errors.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
