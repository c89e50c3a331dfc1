"""Prints the fp32-tier errors of the predictor against the reference fixture (tests/golden/predictor_small.pt) with the mask
head on TF32 operands vs bf16 operands -- the numbers behind the tolerances of tests/test_modules_gpu.py::test_predictor_golden."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from dvis_plus_b200 import ops
from dvis_plus_b200.modules.precision import precision
from test_modules_cpu import build_predictor


def rel(a, b):
    return ((a.float().cpu() - b.float()).abs().max() / b.float().abs().max()).item()


g = torch.load(os.path.join(ROOT, "tests", "golden", "predictor_small.pt"))
d = build_predictor(g).cuda()
out = {}
with torch.no_grad():
    for name, operand in (("tf32", torch.float32), ("bf16", torch.bfloat16)):
        for materialize in (True, False):
            d.materialize_aux_masks = materialize
            with precision("fp32"):
                ops.set_mask_operand_dtype(operand)
                o = d([x.cuda() for x in g["multi_scale"]], g["mask_features"].cuda())
            out[f"{name}_materialize_{materialize}"] = {k: rel(o[k], g[k]) for k in ("pred_logits", "pred_masks", "pred_embds",
                                                                                    "pred_embds_without_norm")}
print(json.dumps(out, indent=1))
