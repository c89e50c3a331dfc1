"""What does a high-priority chain of tiny dependent kernels (the temporal stage) cost a stream of big kernels (the per-frame
stage) running next to it?  For each (big kernel type A) x (tiny kernel type B): time A's graph alone, B's graph alone, and both
together (B on a priority stream, long enough to cover A).  Writes gpurun_out/interference_probe.json."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dvis_plus_b200 import ops  # noqa: E402


def capture(fn, n, stream):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream), torch.no_grad():
        for _ in range(n):
            fn()
    return g


def main():
    dev = "cuda"
    T, S, C = 16, 19320, 256
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(T, S, C, generator=gen, device=dev).bfloat16()
    res = torch.randn(T, S, C, generator=gen, device=dev)
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    w1 = (torch.randn(1024, C, generator=gen, device=dev) / 16).bfloat16()
    b1 = torch.zeros(1024, device=dev).bfloat16()
    xs = torch.randn(200, 512, generator=gen, device=dev).bfloat16()
    ws = (torch.randn(512, 512, generator=gen, device=dev) / 22).bfloat16()
    bs = torch.zeros(512, device=dev).bfloat16()
    r32 = torch.randn(200, 512, generator=gen, device=dev)
    g5, b5 = torch.ones(512, device=dev), torch.zeros(512, device=dev)
    q = torch.randn(1, 200, 8, 64, generator=gen, device=dev).bfloat16()
    big = {
        "add_layernorm_309k_rows": (lambda: ops.add_layernorm(x, res, gamma, beta, 1e-5, lp_dtype=torch.bfloat16), 60),
        "cublas_ffn1_309k_rows": (lambda: torch._addmm_activation(b1, x.view(-1, C), w1.t()), 40),
        "bf16_copy_158MB": (lambda: x.clone(), 100),
    }
    tiny = {
        "cublas_gemm_200x512x512": (lambda: F.linear(xs, ws, bs), 1500),
        "aten_add_200x512": (lambda: xs + 1, 1500),
        "add_layernorm_200x512": (lambda: ops.add_layernorm(xs, r32, g5, b5, 1e-5, lp_dtype=torch.bfloat16), 1500),
        "flash_attn_200q": (lambda: ops.flash_attn(q, q, q, 0.125), 1500),
    }
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)
    out = {}

    def run(ga, gb):
        torch.cuda.synchronize()
        ea0, ea1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eb0, eb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if gb is not None:
            with torch.cuda.stream(sb):
                eb0.record(); gb.replay(); eb1.record()
        if ga is not None:
            with torch.cuda.stream(sa):
                ea0.record(); ga.replay(); ea1.record()
        torch.cuda.synchronize()
        return (ea0.elapsed_time(ea1) if ga is not None else None, eb0.elapsed_time(eb1) if gb is not None else None)

    graphs_b = {k: capture(fn, n, sb) for k, (fn, n) in tiny.items()}
    for ka, (fa, na) in big.items():
        ga = capture(fa, na, sa)
        run(ga, None)
        a_alone = min(run(ga, None)[0] for _ in range(3))
        out[ka] = {"alone_ms": round(a_alone, 3)}
        for kb, gb in graphs_b.items():
            run(None, gb)
            b_alone = min(run(None, gb)[1] for _ in range(2))
            both = [run(ga, gb) for _ in range(3)]
            a_with = min(t[0] for t in both)
            b_with = min(t[1] for t in both)
            n_b = tiny[kb][1]
            # B covers A entirely when b_with >= a_with; cost per tiny kernel = A's slowdown / number of tiny kernels that ran meanwhile
            overl = min(1.0, a_with / b_with) * n_b
            out[ka][kb] = {"tiny_chain_alone_ms": round(b_alone, 3), "big_with_chain_ms": round(a_with, 3), "chain_with_big_ms": round(b_with, 3),
                           "big_slowdown_ms": round(a_with - a_alone, 3),
                           "us_of_big_stream_lost_per_tiny_kernel": round((a_with - a_alone) * 1e3 / overl, 2)}
            print(ka, kb, out[ka][kb], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "interference_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
