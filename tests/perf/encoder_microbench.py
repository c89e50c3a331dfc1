"""Encoder-layer pieces at the bench shape (T frames x 19 320 tokens x 256, bf16): library GEMM + separate passes vs
csrc/linear_tc.cu, token-major vs head-major gather.  CUDA events, inputs far larger than L2."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import torch
import torch.nn.functional as F

from dvis_plus_b200 import ops
from dvis_plus_b200.locality import tiled_item_order


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main(T=16):
    shapes = ((92, 160), (46, 80), (23, 40))
    M, D, L, P, C = 8, 32, 3, 4, 256
    S = sum(h * w for h, w in shapes)
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(T, S, C, generator=g, device=dev).bfloat16()
    w = (torch.randn(C, C, generator=g, device=dev) / 16).bfloat16()
    b32 = torch.randn(C, generator=g, device=dev)
    bb = b32.bfloat16()
    res = torch.randn(T, S, C, generator=g, device=dev)
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    sh = torch.as_tensor(shapes, dtype=torch.long, device=dev)
    lsi = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    fused = (torch.randn(T, S, M * L * P * 3, generator=g, device=dev) * 1.5).bfloat16()
    n_off = M * L * P * 2
    ys, xs = [], []
    for h, w_ in shapes:
        yy, xx = torch.meshgrid((torch.arange(h, device=dev) + 0.5) / h, (torch.arange(w_, device=dev) + 0.5) / w_, indexing="ij")
        ys.append(yy.reshape(-1)); xs.append(xx.reshape(-1))
    ref = torch.stack([torch.cat(xs), torch.cat(ys)], -1)[None, :, None, :].expand(T, S, L, 2).contiguous()
    order = tiled_item_order(shapes, M, torch.device(dev))
    value = F.linear(x, w, bb).view(T, S, M, D)
    value_hm = ops.linear_tc_heads(x, w, b32)
    out = {"T": T, "tokens": T * S}
    out["value_proj_cublas_us"] = timeit(lambda: F.linear(x, w, bb))
    out["value_proj_tc_heads_us"] = timeit(lambda: ops.linear_tc_heads(x, w, b32))
    out["linear_tc_plain_us"] = timeit(lambda: ops.linear_tc(x, w, b32))
    out["msda_token_major_us"] = timeit(lambda: ops.msda_fused_forward(value, sh, lsi, fused[..., :n_off], fused[..., n_off:], ref, M, L, P,
                                                                       item_order=order))
    out["msda_head_major_us"] = timeit(lambda: ops.msda_fused_forward_hm(value_hm, sh, lsi, fused[..., :n_off], fused[..., n_off:], ref, L, P,
                                                                         item_order=order))
    out["out_proj_cublas_plus_add_layernorm_us"] = timeit(
        lambda: ops.add_layernorm(F.linear(x, w, bb), res, gamma, beta, 1e-5, lp_dtype=torch.bfloat16))
    out["add_layernorm_alone_us"] = timeit(lambda: ops.add_layernorm(x, res, gamma, beta, 1e-5, lp_dtype=torch.bfloat16))
    out["out_proj_tc_add_ln_us"] = timeit(lambda: ops.linear_tc_add_layernorm(x, w, b32, res, gamma, beta, 1e-5))
    out["out_proj_tc_add_ln_bf16_only_us"] = timeit(lambda: ops.linear_tc_add_layernorm(x, w, b32, res, gamma, beta, 1e-5, want_f32=False))
    out["out_proj_tc_add_ln_f32_only_us"] = timeit(lambda: ops.linear_tc_add_layernorm(x, w, b32, res, gamma, beta, 1e-5, want_lp=False))
    # algorithmic bytes of the fused output projection: x bf16 + residual f32 in, f32 + bf16 out
    byts = T * S * C * (2 + 4 + 4 + 2)
    out["out_proj_tc_add_ln_GBps"] = byts / out["out_proj_tc_add_ln_us"] / 1e3
    out["value_proj_tc_heads_GBps"] = T * S * C * 4 / out["value_proj_tc_heads_us"] / 1e3
    a = ops.msda_fused_forward(value, sh, lsi, fused[..., :n_off], fused[..., n_off:], ref, M, L, P, item_order=order)
    h = ops.msda_fused_forward_hm(value_hm, sh, lsi, fused[..., :n_off], fused[..., n_off:], ref, L, P, item_order=order)
    out["msda_hm_vs_token_major_max_abs"] = (a.float() - h.float()).abs().max().item()
    print(json.dumps(out))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 16)
