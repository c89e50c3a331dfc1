"""What a DEPENDENT chain of temporal-stage kernels costs on a B200 (the tracker's critical path is 36 such steps per frame):
a CUDA graph of `n` launches, each consuming the previous one's output, replayed; per-step time = total / n.
Also prints the in-kernel clock64 stamps of dvis_linear_small (DVIS_LS_PROF=1).  Writes gpurun_out/r2_chain_probe.json."""
import ctypes
import json
import math
import os
import sys

os.environ["DVIS_LS_PROF"] = "1"
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dvis_plus_b200 import _lib, ops  # noqa: E402


def graph_time(fn, iters=30):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def stamps():
    buf = (ctypes.c_longlong * 8)()
    torch.cuda.synchronize()
    _lib.lib().dvis_debug_linear_small_stamps(buf)
    t = list(buf)
    return {"ring_issued": t[1] - t[0], "first_chunk_landed": t[2] - t[1], "k_loop": t[3] - t[2], "epilogue": t[4] - t[3], "total_cycles": t[4] - t[0]}


@torch.no_grad()
def main():
    dev = "cuda"
    out = {"stamps_cycles": {}, "chain_us_per_step": {}}
    M, C = 200, 512
    w = [(torch.randn(C, C, device=dev) / C ** 0.5).to(torch.bfloat16) for _ in range(4)]
    b = torch.randn(C, device=dev)
    b16 = b.to(torch.bfloat16)
    x = torch.randn(M, C, device=dev).to(torch.bfloat16)
    res = torch.randn(M, C, device=dev)
    g = torch.ones(C, device=dev)
    w2048 = (torch.randn(2048, C, device=dev) / C ** 0.5).to(torch.bfloat16)
    wdown = (torch.randn(C, 2048, device=dev) / 2048 ** 0.5).to(torch.bfloat16)
    b2048 = torch.randn(2048, device=dev)
    for name, fn in {"plain (200,512,512)": lambda: ops.linear_small(w[0], b, x=x),
                     "plain (200,512,2048)": lambda: ops.linear_small(wdown, b, x=torch.zeros(M, 2048, device=dev, dtype=torch.bfloat16)),
                     "ln prologue (200,1536,512)": lambda: ops.linear_small(torch.cat(w[:3]), None, src0=res, ln1=(g, g), want_side1=True)}.items():
        for _ in range(3):
            fn()
        out["stamps_cycles"][name] = stamps()
    for n in (1, 8, 32):
        def chain_ours():
            y = x
            for i in range(n):
                y = ops.linear_small(w[i % 4], b, x=y)[1]
            return y

        def chain_lib():
            y = x
            for i in range(n):
                y = torch.addmm(b16, y, w[i % 4].t())
            return y

        def chain_block_ours():          # out-proj(+res) -> FFN1 (LN prologue, ReLU) -> FFN2(+res): the tracker's layer tail
            pre = res
            for i in range(n):
                _, h, _, x2 = ops.linear_small(w2048, b2048, src0=pre, ln1=(g, g), want_side1=True, relu=True)
                pre = ops.linear_small(wdown, b, x=h, residual=x2, out_f32=True, out_bf16=False)[0]
            return pre

        def chain_block_lib():
            x32 = res
            for i in range(n):
                y32, y16, _ = ops.add_layernorm(x32, None, g, g, 1e-5, lp_dtype=torch.bfloat16)
                h = torch._addmm_activation(b2048.to(torch.bfloat16), y16, w2048.t())
                x32 = torch.addmm(b16, h, wdown.t()).float() + y32
            return x32
        out["chain_us_per_step"]["linear 200x512x512 x%d" % n] = {"dvis_linear_small": round(graph_time(chain_ours) / n, 2),
                                                                  "torch_addmm": round(graph_time(chain_lib) / n, 2)}
        out["chain_us_per_step"]["ffn block (LN,FFN1,FFN2,res) x%d" % n] = {"ours (2 launches)": round(graph_time(chain_block_ours) / n, 2),
                                                                             "library (LN + 2 GEMM + add, 5 launches)": round(graph_time(chain_block_lib) / n, 2)}
    for pdl in (False, True):
        ops.set_pdl(pdl)

        def chain32():
            y = x
            for i in range(32):
                y = ops.linear_small(w[i % 4], b, x=y)[1]
            return y
        out["chain_us_per_step"]["linear x32, pdl=%s" % pdl] = round(graph_time(chain32) / 32, 2)
    ops.set_pdl(False)
    q = torch.randn(1, 200, 3, 8, 64, device=dev).to(torch.bfloat16)

    def attn_chain_ours():
        o = None
        for _ in range(16):
            o = ops.flash_attn(q[:, :, 0], q[:, :, 1], q[:, :, 2], 0.125)
        return o

    def attn_chain_lib():
        o = None
        for _ in range(16):
            o = F.scaled_dot_product_attention(q[:, :, 0].transpose(1, 2), q[:, :, 1].transpose(1, 2), q[:, :, 2].transpose(1, 2), scale=0.125)
        return o
    out["chain_us_per_step"]["attention (1,200,200,8x64) x16"] = {"dvis_flash_attn": round(graph_time(attn_chain_ours) / 16, 2),
                                                                  "torch_sdpa": round(graph_time(attn_chain_lib) / 16, 2)}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_chain_probe.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
